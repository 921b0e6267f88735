"""Build libclipppo_b200.so (sm_100a only) in-tree with nvcc.

    python clip-ppo_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the gpurun
snapshot.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libclipppo_b200.so")

SOURCES = ["capi.cu", "prof.cu", "disturb.cu", "losses.cu", "preprocess.cu", "layernorm.cu",
           "gemm_tcgen05.cu", "attention.cu", "attention_tc.cu", "vit.cu", "policy.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-I", INCLUDE]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CLIP-PPO B200 kernels cannot be built")
    return nvcc


def _newest_header() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    deps.append(os.path.abspath(__file__))
    return max(os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    # measurement-only GEMM variants (tools/gemm_probe.py): never in the product library unless asked for
    flags = NVCC_FLAGS + (["-DCLIPPPO_BUILD_PROBES"] if os.environ.get("CLIPPPO_BUILD_PROBES", "") not in ("", "0") else [])
    hdr = _newest_header()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for src in sources:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            cmd = [nvcc, *flags, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"--- {src}\n{r.stderr}")
        return src

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))

    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s in sources]
    if force or jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
