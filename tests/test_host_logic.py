"""Host-side pieces of the product library that have no device counterpart: the base packer (both ISA builds of
csrc/bsw_pack.cpp), the tabulated band clamp, and chunk-size independence of the scheduler.  CPU only."""
import ctypes as C

import pytest


@pytest.mark.parametrize("seed,ntasks,max_len", [(1, 400, 70), (2, 300, 300), (3, 60, 5000), (4, 500, 33)])
def test_packer_builds_agree_with_scalar_and_never_overread(B, seed, ntasks, max_len):
    """Every sequence sits flush against a PROT_NONE page: reading one byte past it would fault.  Nibbles, the zero
    padding up to 32 bases, the N flag and the rejection of codes > 4 are compared with a scalar packer."""
    E = B.emu_lib()
    E.bsw_emu_pack_check.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
    wide = C.c_int(0)
    rc = E.bsw_emu_pack_check(seed, ntasks, max_len, C.byref(wide))
    assert rc == 0, f"pack check failed: rc={rc} (avx512 build exercised: {bool(wide.value)})"


def test_band_clamp_table_equals_the_divisions(B):
    assert B.emu_lib().bsw_emu_band_clamp_check() == 0


def test_results_do_not_depend_on_sse2_vs_avx512_pack(B, O):
    """The run-time dispatch is read once per process, so this runs the SSE2 build in a child process and compares the
    emulated batch with the oracle there."""
    import os, subprocess, sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, bsw_b200 as B, oracle as O\n"
        "t = B.synth_tasks('cfg3_mixed', 600)\n"
        "p, po = B.make_params(), O.make_params()\n"
        "ro, co = O.extend_batch(po, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])\n"
        "re, ce, info = B.emu_extend_batch(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])\n"
        "assert np.array_equal(ro, re) and np.array_equal(co.astype(np.int64), ce.astype(np.int64))\n"
        "print('ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BSW_NO_AVX512="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
