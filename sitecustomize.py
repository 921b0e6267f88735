"""Imported automatically by Python when this repository is on PYTHONPATH: installs the top-level aliases
(`import clip_ppo_utils` -> this repository's shared.clip_ppo_utils) the reference scripts' sys.path hack needs.
See clip-ppo_b200/dropin.py.  Nothing else happens here - in particular torch is not imported.  A `sitecustomize`
further down sys.path (e.g. the distribution's) is shadowed by this file, so it is run from here afterwards."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _install():
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
