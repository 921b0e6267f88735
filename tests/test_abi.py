"""The C-ABI library loads and exports every symbol include/clipppo_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "clipppo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef CLIPPPO_BUILD_PROBES.*?#endif", "", text, flags=re.S)      # measurement-only entry points
    return sorted(set(re.findall(r"\b(clipppo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(native, n), f"{n} declared in the header but not exported"


def test_binding_lists_every_header_symbol():
    from clip_ppo_b200 import _native
    assert sorted(_native.SYMBOLS) == _declared_symbols()


def test_version_and_strerror(native):
    assert native.clipppo_abi_version() == 4
    assert native.clipppo_strerror(0) == b"ok"
    assert b"channels" in native.clipppo_strerror(-2)
    assert native.clipppo_strerror(-12345) == b"unknown status"


def test_argument_validation_without_gpu(native):
    """Entry points reject bad arguments before touching the device."""
    i64x4 = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    taps = (ctypes.c_float * 3)(0.25, 0.5, 0.25)
    # null image pointer
    assert native.clipppo_disturb_f32(None, i64x4, None, i64x4, None, 1, 3, 8, 8, 15, 0.1, 1.0, taps, 3, 0, 0, 2, 2, None) == -4
    # non-positive shape
    assert native.clipppo_disturb_f32(None, i64x4, None, i64x4, None, 0, 3, 8, 8, 15, 0.1, 1.0, taps, 3, 0, 0, 2, 2, None) == -1
    assert native.clipppo_gae_f32(None, None, None, None, None, 4, 4, 0.99, 0.95, None, None, None) == -4
    assert native.clipppo_cosine_loss_fwd(None, None, 4, 512, None, None, None) == -4
    assert native.clipppo_gemm_bf16(None, None, 128, 256, 64, 0, None, None, 0, None, 256, None) == -4


import pytest


@pytest.mark.gpu
def test_header_symbols_and_sass_on_the_gpu_box(native):
    """The same header / export / SASS checks inside the `-m gpu` run, against the library the GPU tests actually load."""
    test_header_symbols_are_exported(native)
    test_binding_lists_every_header_symbol()
    test_library_is_sm100a_native()
    test_sass_keeps_the_round2_issue_and_barrier_properties()


def test_library_is_sm100a_native():
    """The shipped code object targets sm_100a and contains the tcgen05 / TMA instructions."""
    import shutil
    import subprocess
    from clip_ppo_b200 import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    # the long-sequence attention kernel is tcgen05 / TMEM / TMA as well: MMAs, TMEM loads AND stores (P goes back into
    # TMEM as the A operand of P V), 3-D tensor-map loads and stores
    att = sass[sass.index("attention_tc_kernel"):]
    att = att[:att.index("Function :", 10)] if "Function :" in att[10:] else att
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG.3D", "UTMASTG.3D"):
        assert mnemonic in att, f"{mnemonic} missing from attention_tc_kernel"
    assert "HMMA.16816" not in att


def _kernel_sass(sass: str, needle: str) -> str:
    """SASS text of the first function whose mangled name contains `needle`."""
    i = sass.index(needle)
    start = sass.rfind("Function :", 0, i)
    end = sass.find("Function :", i)
    return sass[start:end if end > 0 else len(sass)]


def test_sass_keeps_the_round2_issue_and_barrier_properties():
    """Three properties found with `ncu --page source` in round 2, pinned so that a refactor cannot silently lose them:
    (1) the disturbance kernel's cluster barrier arrives relaxed - ONE MEMBAR.ALL.GPU (the publishing thread's fence), not one
        per arrive (profiles/r02_experiments.md section 10);
    (2) the GEMM / tcgen05 attention staging tiles are accessed in the shared address space (STS / LDS), not through generic
        ST.E / LD.E (section 11);
    (3) the issuer warps run converged: the four tcgen05.mma of a GEMM k-block are consecutive instructions on uniform
        registers, no ELECT / R2UR loop around each (section 12)."""
    import re
    import shutil
    import subprocess
    from clip_ppo_b200 import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    # (1) plain fast kernel, k = 7, 224 wide
    dist = _kernel_sass(sass, "disturb_fast_kernelILi7ELi224ELb0ELb0ELi0EEE")
    assert dist.count("MEMBAR.ALL.GPU") == 1, dist.count("MEMBAR.ALL.GPU")
    assert dist.count("UCGABAR_ARV") == 2 and dist.count("UCGABAR_WAIT") == 2
    # (2) + (3) the QKV / c_fc / residual GEMMs in pair mode and the three attention kernels
    for needle in ("gemm_bf16_kernelILi6ELi2ELi0EEE", "gemm_bf16_kernelILi7ELi2ELi0EEE", "gemm_bf16_kernelILi9ELi2ELi0EEE",
                   "attention_tc_kernelE", "attention_tc_pair_kernelILb1EEE", "attention_tc_pair_kernelILb0EEE"):
        k = _kernel_sass(sass, needle)
        assert not re.search(r"\s(LD|ST)\.E\.128", k), f"generic 128-bit shared-memory access in {needle}"
        assert "STS.128" in k or "STS" in k
    lines = [re.sub(r"/\*.*?\*/", "", l).strip() for l in _kernel_sass(sass, "gemm_bf16_kernelILi6ELi2ELi0EEE").splitlines()]
    ops = [l.split()[0] for l in lines if l and not l.startswith(("Function", ".", "="))]
    first = ops.index("UTCHMMA.2CTA")
    assert ops[first:first + 4] == ["UTCHMMA.2CTA"] * 4, ops[first - 2:first + 6]
