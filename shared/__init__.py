"""Drop-in `shared` package: the reference's Python surface for the observation path
(disturbances_gpu, clip_ppo_utils, disturbance_types), backed by the sm_100a kernels.

The reference's own `shared/` directory is a namespace package (no __init__.py); once this repository is on
PYTHONPATH this regular package wins the name.  The modules it does not replace - `shared.disturbances` (the cv2
CPU twin), `shared.checkpoint_utils`, the benchmark / test scripts - must stay importable, so every other `shared`
directory found on sys.path is appended to this package's search path (after this one: same-named modules resolve
here).  The training scripts put their directories on sys.path right before importing, so the scan also reruns
when a submodule is missing.
"""
import importlib as _importlib
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def _extend_search_path() -> None:
    for entry in list(_sys.path):
        cand = _os.path.abspath(_os.path.join(entry or _os.curdir, "shared"))
        if cand != _HERE and _os.path.isdir(cand) and cand not in __path__:
            __path__.append(cand)


_extend_search_path()


def __getattr__(name: str):
    """`from shared import disturbances` after a late sys.path change: rescan, then import the submodule."""
    if name.startswith("_"):
        raise AttributeError(name)
    _extend_search_path()
    try:
        return _importlib.import_module(f"{__name__}.{name}")
    except ModuleNotFoundError as e:
        if e.name != f"{__name__}.{name}":
            raise
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
