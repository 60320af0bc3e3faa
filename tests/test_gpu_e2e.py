"""GPU suite, end to end: the product's soap4 driver (host C++ + libmegapath_b200.so) must print the same
annotated FASTQ as the reference binary on the same index and reads, after the canonical sort of pairs.
Covers all stages: deep DP (S1), single-end DP (S2), mate rescue (S3), unpaired output (S4)."""
import os

import numpy as np
import pytest

from conftest import make_reads, run_ref_soap4, run_our_soap4, canon_fastq, needs_ref

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

E2E_SETS = [
    ("clean", 150, 151, dict(model="clean"), "soap4.ini", ("-F", "-nc")),
    ("mixed", 100, 101, dict(model="divergent", one_random=0.10, unalignable=0.05), "soap4.ini", ("-F", "-nc")),
    ("var", 150, 151, dict(model="clean", varlen=True, n_rate=0.002, one_random=0.05), "soap4.ini", ("-F", "-nc")),
    ("long", 250, 251, dict(model="divergent", one_random=0.05), "soap4.ini", ("-F", "-nc")),
    ("nt2", 150, 151, dict(model="divergent", one_random=0.10, unalignable=0.02), "soap4-nt2.ini", ("-F", "-nc", "-top", "95")),
    ("pmode", 100, 101, dict(model="divergent", one_random=0.10, unalignable=0.05), "soap4.ini", ("-P", "-nc")),
    ("span", 150, 151, dict(model="clean", span_frac=0.05), "soap4.ini", ("-F", "-nc")),
]


def first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return i, x[:300], y[:300]
    return min(len(la), len(lb)), b"<end>", b"<end>"


@needs_ref
@pytest.mark.parametrize("name,rlen,lopt,kw,ini,extra", E2E_SETS)
def test_e2e_stdout_matches_reference(workdir, small_ref, name, rlen, lopt, kw, ini, extra):
    npairs = 3000 if name != "long" else 800
    fq1, fq2 = make_reads(workdir, small_ref, "e2e_" + name, npairs, rlen, seed=31, **kw)
    ref_out, _ = run_ref_soap4(workdir, small_ref["prefix"], fq1, fq2, "e2eref_" + name, lopt, dump=False, ini=ini, threads=4,
                               extra=[x for x in extra if x not in ("-F", "-nc")] + (["-P"] if "-P" in extra else []))
    want = canon_fastq(open(ref_out, "rb").read())
    got = canon_fastq(run_our_soap4(workdir, small_ref["prefix"], fq1, fq2, "e2eour_" + name, lopt, ini=ini, extra=extra))
    assert len(want) > 1000
    assert got == want, first_diff(got, want)


@pytest.mark.parametrize("name", ["clean", "div", "nt2"])
def test_e2e_stdout_matches_golden(tmp_path, name):
    """Same check against the committed reference outputs (no oracle/_ref needed): index built by the GPU builder."""
    import megapath_b200 as mp
    z = np.load(os.path.join(G, name + ".npz"))
    lopt, nt2 = int(z["lopt"][0]), bool(z["nt2"][0])
    pac = np.fromfile(os.path.join(G, "idx.pac"), dtype=np.uint8)
    n = int(open(os.path.join(G, "idx.ann")).readline().split()[0])
    prefix = str(tmp_path / "g.index")
    c = mp.Context(0)
    c.index_build(pac[:(n + 3) // 4], n)
    c.index_save(prefix)
    c.close()
    for ext in ("ann", "amb", "tra"):
        open(prefix + "." + ext, "wb").write(open(os.path.join(G, "idx." + ext), "rb").read())
    got = canon_fastq(run_our_soap4(str(tmp_path), prefix, os.path.join(G, name + "_1.fq"), os.path.join(G, name + "_2.fq"), "g_" + name, lopt,
                                    ini="soap4-nt2.ini" if nt2 else "soap4.ini"))
    want = bytes(z["fastq"])
    assert got == want, first_diff(got, want)
