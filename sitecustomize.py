"""Imported automatically by Python when this repository is on PYTHONPATH: installs the top-level aliases
(`import clip_ppo_utils` -> this repository's shared.clip_ppo_utils) the reference scripts' sys.path hack needs.
See clip-ppo_b200/dropin.py.  Nothing else happens here - in particular torch is not imported."""
import importlib.abc
import importlib.util
import os
import sys


def _install():
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_clipppo_dropin", os.path.join(here, "clip-ppo_b200", "dropin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.install()


try:
    _install()
except Exception as e:      # never break interpreter start-up
    sys.stderr.write(f"clip-ppo-b200 sitecustomize: aliases not installed ({e})\n")
