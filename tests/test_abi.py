"""CPU suite: the C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls without a GPU), and the host mirror packs buffers like the reference."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "megapath_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import megapath_b200 as mp
    if not os.path.exists(mp.LIB_PATH):
        mp.build()
    lib = ctypes.CDLL(mp.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s


def test_no_gpu_fails_loudly():
    import megapath_b200 as mp
    import torch
    if torch.cuda.is_available():
        return
    try:
        mp.Context(0)
    except mp.MegapathError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("mp_init must fail without a GPU (no CPU fallback)")


def test_struct_sizes_match_header():
    import megapath_b200 as mp
    assert mp.PAIR_RESULT.itemsize == 104
    assert mp.SINGLE_RESULT.itemsize == 56
    assert ctypes.sizeof(mp.MmpParams) == 48 and ctypes.sizeof(mp.AlignParams) == 48 + 12 * 4
    assert ctypes.sizeof(mp.Results) == 112 and ctypes.sizeof(mp.Stats) == 120
    assert mp.SEEDPOS.itemsize == 16 and mp.CAND.itemsize == 24


def test_pack_queries_layout():
    """appendToQueryArrays layout (QueryParser.cpp:184-203)."""
    import megapath_b200 as mp
    rng = np.random.default_rng(0)
    n, L = 70, 151
    lens = rng.integers(40, L, size=n).astype(np.uint32)
    codes = rng.integers(0, 4, size=(n, L)).astype(np.uint8)
    il, wpq = mp.pack_queries(codes, lens, L)
    assert wpq == 10
    for r in (0, 1, 31, 32, 69):
        for i in range(int(lens[r])):
            w = il[(r // 32) * 32 * wpq + (r % 32) + 32 * (i // 16)]
            assert (int(w) >> (2 * (i % 16))) & 3 == codes[r, i]
        # bases beyond the read length are zero
        last = il[(r // 32) * 32 * wpq + (r % 32) + 32 * (int(lens[r]) // 16)]
        assert int(last) >> (2 * (int(lens[r]) % 16)) == 0
