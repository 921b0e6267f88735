"""Stage the REFERENCE's own hot-path modules under oracle/_ref/ and load them.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it means copying three importable files, unmodified, from where they lie
under /root/reference into ``oracle/_ref/shared/``:

    shared/disturbances_gpu.py     DisturbanceWrapperGPU                        (reference shared/disturbances_gpu.py:14-214)
    shared/disturbance_types.py    DisturbanceSeverity, SEVERITY_CONFIGS        (shared/disturbance_types.py:8-43)
    shared/clip_ppo_utils.py       generate_clip_embeddings, cosine loss, ...   (shared/clip_ppo_utils.py:26-240)

``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored: staged here by
``__graft_entry__.build()``, it travels to the GPU box with the snapshot exactly like the built ``.so``.  There
``bench.py --impl reference`` / ``cpu_baseline`` time these functions - the reference's stock code path - on the host cores
(``kind: "reference"``), and the ``eager_gpu_baseline`` leg runs the same functions with ``device="cuda"``.

    python oracle/build_ref.py            # stage (needs /root/reference) and print the manifest

openai/CLIP (``import clip`` at shared/clip_ppo_utils.py:6) is absent on every box: ``load()`` installs a stub ``clip``
module for the duration of the import whose ``clip.load`` returns the tower the caller supplies (the fp32 oracle tower on
the CPU, an fp16 eager torch tower on the GPU) - the same arrangement ``oracle/make_goldens.py`` generates the fixtures with.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
STAGE = os.path.join(HERE, "_ref")
FILES = ("shared/disturbances_gpu.py", "shared/disturbance_types.py", "shared/clip_ppo_utils.py")


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(ref_root: str = REF_ROOT) -> bool:
    """Copy the three files (byte for byte) and write a manifest with their hashes.  Returns False, touching nothing, when
    the reference tree is not present (the GPU box: the prebuilt copy is used there)."""
    if not all(os.path.isfile(os.path.join(ref_root, f)) for f in FILES):
        return False
    manifest = {}
    for f in FILES:
        dst = os.path.join(STAGE, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, f), dst)
        manifest[f] = _sha(dst)
    with open(os.path.join(STAGE, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref_root, "sha256": manifest}, fh, indent=1)
    return True


def available() -> bool:
    """True when a complete, unmodified staged copy exists (hashes match the manifest)."""
    try:
        manifest = json.load(open(os.path.join(STAGE, "MANIFEST.json")))["sha256"]
        return all(_sha(os.path.join(STAGE, f)) == manifest[f] for f in FILES)
    except (OSError, KeyError, ValueError):
        return False


def make_stub_clip(tower):
    """The surface of the openai package that shared/clip_ppo_utils.py touches (:6, :90, :136, :187, :212), with
    ``clip.load`` handing back ``tower`` (anything with ``encode_image``, ``eval()`` and ``parameters()``)."""
    import torch
    clip = types.ModuleType("clip")
    model = types.ModuleType("clip.model")

    class VisionTransformer(torch.nn.Module):
        pass

    class CLIP(torch.nn.Module):
        pass

    def _no_tokenizer(*a, **k):
        raise NotImplementedError("the stub clip module has no BPE tokenizer")

    model.VisionTransformer, model.CLIP = VisionTransformer, CLIP
    clip.model = model
    clip.load = lambda name, device="cpu", **kw: (tower, None)
    clip.tokenize = _no_tokenizer
    return clip, model


def load(tower):
    """Import the staged reference modules under private names and return them as a namespace
    (``disturbances_gpu``, ``disturbance_types``, ``clip_ppo_utils``).  ``shared`` and ``clip`` are rebound only while
    the imports run; this repository's own drop-in ``shared`` package is put back afterwards."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run python oracle/build_ref.py where /root/reference exists)")
    saved = {k: v for k, v in sys.modules.items() if k == "shared" or k.startswith("shared.") or k == "clip" or k.startswith("clip.")}
    for k in saved:
        del sys.modules[k]
    try:
        pkg = types.ModuleType("shared")
        pkg.__path__ = [os.path.join(STAGE, "shared")]
        sys.modules["shared"] = pkg
        sys.modules["clip"], sys.modules["clip.model"] = make_stub_clip(tower)
        mods = {}
        for name in ("disturbance_types", "disturbances_gpu", "clip_ppo_utils"):
            spec = importlib.util.spec_from_file_location(f"shared.{name}", os.path.join(STAGE, "shared", f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"shared.{name}"] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k in [m for m in sys.modules if m == "shared" or m.startswith("shared.") or m == "clip" or m.startswith("clip.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return types.SimpleNamespace(**mods)


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "reference tree not found; nothing staged", STAGE)
    if available():
        print(open(os.path.join(STAGE, "MANIFEST.json")).read())
