"""Top-level module aliases for the reference's `sys.path` hack.

`clip_ppo_minigrid.py:27-29` / `clip_ppo_atari.py:26` put the reference's `shared/` DIRECTORY at the front of
sys.path and `import clip_ppo_utils` as a top-level module - which would find the reference's file whatever
PYTHONPATH says.  `install()` adds a meta-path finder that resolves the three top-level names this repository
replaces to its own `shared.*` modules (the same module objects, so enums compare equal under both names).
`sitecustomize.py` at the repository root calls it, which makes `PYTHONPATH=<this repo>` a complete drop-in."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

ALIASES = {
    "clip_ppo_utils": "shared.clip_ppo_utils",
    "disturbances_gpu": "shared.disturbances_gpu",
    "disturbance_types": "shared.disturbance_types",
}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if path is None and fullname in ALIASES:
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        real = importlib.import_module(ALIASES[spec.name])
        self._keep = (real.__spec__, real.__loader__, real.__name__)
        return real

    def exec_module(self, module):
        # the import machinery stamped the alias' spec onto the shared module object: put its own identity back
        module.__spec__, module.__loader__, module.__name__ = self._keep


def install() -> None:
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())
