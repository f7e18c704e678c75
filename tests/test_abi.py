"""The drop-in boundary: libbsw.so loads, exports every symbol include/bsw.h declares, and refuses to run without a
CUDA device instead of silently computing on the CPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "bsw.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsw_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_three_levels():
    names = declared_functions()
    for f in ("bsw_init", "bsw_destroy", "bsw_extend_batch", "bsw_extend_batch_flat", "bsw_chain2aln_batch",
              "bsw_fpga_batch", "bsw_tbb_encode", "bsw_rbb_decode", "bsw_submit", "bsw_poll", "bsw_wait"):
        assert f in names


def test_library_exports_every_declared_symbol(B):
    L = B.lib()
    for f in declared_functions():
        assert hasattr(L, f), f"libbsw.so does not export {f}"
    assert b"sm_100a" in L.bsw_version()


def test_struct_sizes_match_header(B):
    assert C.sizeof(B.Params) == 52 and C.sizeof(B.Params2) == 64
    assert C.sizeof(B.Task) == 32 and C.sizeof(B.SeedTask) == 64
    assert B.RESULT_DTYPE.itemsize == 24 and B.ALN_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_device(B):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    with pytest.raises(B.BswError) as e:
        B.Context()
    assert e.value.code == B.BSW_ECUDA


def test_product_library_does_not_link_the_oracle(B):
    import subprocess
    out = subprocess.run(["nm", "-D", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "bswref_" not in out and "bsw_emu_" not in out
