"""OPT-IN drop-in hook.  Python imports a module named `sitecustomize` at start-up when it finds one on sys.path; with this
repository on PYTHONPATH that is this file.  It does NOTHING unless CLIPPPO_DROPIN=1 is set: a repository on the path should
not change how `import clip_ppo_utils` resolves for unrelated programs.

    CLIPPPO_DROPIN=1 PYTHONPATH=/path/to/this/repo python clip_ppo_minigrid.py ...      # unmodified reference script

With the variable set it installs the top-level aliases (`import clip_ppo_utils` -> this repository's shared.clip_ppo_utils)
that the reference scripts' sys.path hack needs - see clip-ppo_b200/dropin.py.  The explicit, preferred route is two lines at
the top of the script:   `from clip_ppo_b200 import dropin; dropin.install()`   (INTEGRATION.md).
A `sitecustomize` further down sys.path (e.g. the distribution's) is shadowed by this file, so it is always chained to."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _install():
    if os.environ.get("CLIPPPO_DROPIN", "") in ("", "0"):
        return
    spec = importlib.util.spec_from_file_location("_clipppo_dropin", os.path.join(_HERE, "clip-ppo_b200", "dropin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.install()


def _chain():
    for entry in sys.path:
        cand = os.path.join(entry or os.curdir, "sitecustomize.py")
        if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != _HERE:
            spec = importlib.util.spec_from_file_location("_shadowed_sitecustomize", cand)
            spec.loader.exec_module(importlib.util.module_from_spec(spec))
            return


for _step in (_install, _chain):
    try:
        _step()
    except Exception as e:      # never break interpreter start-up
        sys.stderr.write(f"clip-ppo-b200 sitecustomize: {_step.__name__} failed ({e})\n")
