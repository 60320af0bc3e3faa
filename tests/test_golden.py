"""Committed golden vectors (tests/golden/, generated from the reference itself by tools/make_golden.py):
the oracle restatement on CPU, and the CUDA path (-m gpu), must both reproduce them bit for bit.
These run where neither /root/reference nor oracle/_ref exists."""
import os

import numpy as np
import pytest

from conftest import load_pairs
from oracle import pyoracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = ["clean", "div", "nt2"]


@pytest.fixture(scope="module")
def gix():
    return po.Index(os.path.join(G, "idx"))     # no .lkt in the fixture: the oracle rebuilds it (LTConstruct.c:46-96)


def dp_tasks(z):
    meta, blob = z["dp_meta"], z["dp_bytes"].tobytes()
    o = 0
    for cl, cr, mm, go, dl, rl, co, sc, hl, mc, pl in meta:
        ref = np.frombuffer(blob, dtype=np.uint8, count=dl, offset=o); o += dl
        rd = np.frombuffer(blob, dtype=np.uint8, count=rl, offset=o); o += rl
        pat = blob[o:o + pl]; o += pl
        yield int(cl), int(cr), int(mm), int(go), ref, rd, int(co), int(sc), int(hl), int(mc), pat


def test_oracle_primitives_golden(gix):
    z = np.load(os.path.join(G, "prim.npz"))
    assert (gix.occ(z["idx"], z["c"]) == z["occ"]).all()
    assert (gix.sa(z["sidx"]) == z["sa"]).all()
    for k, l, r in zip(z["keys"], z["l"], z["r"]):
        assert gix.lkt(k) == (l, r)


def test_oracle_dp_golden():
    z = np.load(os.path.join(G, "dp.npz"))
    for k in range(3):
        maxdna, maxread, cl, cr = [int(x) for x in z["shape%d" % k]]
        for t in range(len(z["dl%d" % k])):
            dl, rl = int(z["dl%d" % k][t]), int(z["rl%d" % k][t])
            co = po.dp_cutoff(rl)
            got = po.dp(z["refs%d" % k][t, :dl], z["reads%d" % k][t, :rl], cl, cr, -2, -3, co)
            pl = int(z["plen%d" % k][t])
            want = (int(z["sc%d" % k][t]), int(z["hl%d" % k][t]), int(z["mc%d" % k][t]), bytes(z["pats%d" % k][t, :pl]))
            assert got == want, (k, t)


@pytest.mark.parametrize("name", SETS)
def test_oracle_seeds_candidates_dp_golden(gix, name):
    z = np.load(os.path.join(G, name + ".npz"))
    lopt, nt2 = int(z["lopt"][0]), bool(z["nt2"][0])
    reads, lens = load_pairs(os.path.join(G, name + "_1.fq"), os.path.join(G, name + "_2.fq"), trunc=lopt - 1)
    rp, mp = gix.seed_pairs(reads, lens, po.mmp_params(nt2))
    assert rp.tobytes() == z["readPos"].tobytes() and mp.tobytes() == z["matePos"].tobytes()
    insert_low = max(1, int(lens[0::2].max()), int(lens[1::2].max()))
    cands = po.pair_candidates(rp, mp, lens, insert_low, 750)
    assert cands.tobytes() == z["cand"].tobytes()
    n = 0
    for cl, cr, mm, go, ref, rd, co, sc, hl, mc, pat in dp_tasks(z):
        assert po.dp(ref, rd, cl, cr, mm, go, co) == (sc, hl, mc, pat)
        n += 1
    assert n > 100


# ------------------------------------------------------------------------------------ CUDA path
@pytest.fixture(scope="module")
def gctx():
    import megapath_b200 as mp
    c = mp.Context(0)
    pac = np.fromfile(os.path.join(G, "idx.pac"), dtype=np.uint8)
    n = int(open(os.path.join(G, "idx.ann")).readline().split()[0])
    c.index_build(pac[:(n + 3) // 4], n)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_primitives_golden(gctx):
    z = np.load(os.path.join(G, "prim.npz"))
    assert (gctx.occ(z["idx"], z["c"]) == z["occ"]).all()
    assert (gctx.sa(z["sidx"]) == z["sa"]).all()
    l, r = gctx.lkt(z["keys"])
    assert (l == z["l"]).all() and (r == z["r"]).all()


@pytest.mark.gpu
def test_gpu_dp_golden(gctx):
    import megapath_b200 as mp
    z = np.load(os.path.join(G, "dp.npz"))
    for k in range(3):
        maxdna, maxread, cl, cr = [int(x) for x in z["shape%d" % k]]
        dl, rl = z["dl%d" % k], z["rl%d" % k]
        cut = np.array([po.dp_cutoff(int(x)) for x in rl], dtype=np.int32)
        sc, hl, mc, pats = gctx.dp_batch(mp.pack_dp_interleaved(z["refs%d" % k], dl, maxdna), dl, maxdna,
                                         mp.pack_dp_interleaved(z["reads%d" % k], rl, maxread), rl, maxread, cut, cl, cr)
        assert (sc == z["sc%d" % k]).all() and (hl == z["hl%d" % k]).all() and (mc == z["mc%d" % k]).all()
        for t in range(len(dl)):
            pl = int(z["plen%d" % k][t])
            assert bytes(pats[t, :pl]) == bytes(z["pats%d" % k][t, :pl]), (k, t)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SETS)
def test_gpu_seeds_candidates_golden(gctx, name):
    import megapath_b200 as mp
    z = np.load(os.path.join(G, name + ".npz"))
    lopt, nt2 = int(z["lopt"][0]), bool(z["nt2"][0])
    reads, lens = load_pairs(os.path.join(G, name + "_1.fq"), os.path.join(G, name + "_2.fq"), trunc=lopt - 1)
    insert_low = max(1, int(lens[0::2].max()), int(lens[1::2].max()))
    q, wpq = mp.pack_queries(reads, lens, lopt)
    gctx.batch_upload(q, lens, wpq)
    P = mp.default_params(nt2=nt2, insert_low=insert_low, insert_high=750, max_read_length=lopt)
    gctx.seed_pairs(P)
    rp, mpos = gctx.download_seedpos()
    assert rp.tobytes() == z["readPos"].tobytes() and mpos.tobytes() == z["matePos"].tobytes()
    assert gctx.download_candidates().tobytes() == z["cand"].tobytes()
