import os, sys, subprocess, shutil, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools import synth
import conftest as ct
import megapath_b200 as mp
from oracle import pyoracle as po

name, rlen, lopt, kw, ini, extra = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), eval(sys.argv[4]), sys.argv[5], tuple(sys.argv[6].split())
npairs = int(sys.argv[7]) if len(sys.argv) > 7 else 3000
wd = tempfile.mkdtemp()
fa = os.path.join(wd, "ref.fa")
seq, bounds = synth.make_ref(300000, 6, seed=42, repeat_frac=0.05)
synth.write_fasta(fa, seq, bounds)
shutil.copy(os.path.join(ct.REF_DIR, "2bwt-builder.ini"), wd)
subprocess.check_call([os.path.join(ct.REF_DIR, "2bwt-builder"), fa], cwd=wd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
sr = dict(fasta=fa, prefix=fa + ".index", seq=seq, bounds=bounds)
fq1, fq2 = ct.make_reads(wd, sr, "e2e_" + name, npairs, rlen, seed=31, **kw)
ref_out, _ = ct.run_ref_soap4(wd, sr["prefix"], fq1, fq2, "ref", lopt, dump=False, ini=ini, threads=4, extra=[x for x in extra if x not in ("-F", "-nc")])
want = ct.canon_fastq(open(ref_out, "rb").read()).split(b"\n")
got = ct.canon_fastq(ct.run_our_soap4(wd, sr["prefix"], fq1, fq2, "our", lopt, ini=ini, extra=extra)).split(b"\n")
bad = []
for i, (a, b) in enumerate(zip(got, want)):
    if a != b:
        bad.append((i, a, b))
print("differing lines:", len(bad), "of", len(want))
for i, a, b in bad[:12]:
    print(i, a[:200], "| want", b[:200])
# library-level detail for the first bad pairs
reads, lens = ct.load_pairs(fq1, fq2, trunc=lopt - 1)
c = mp.Context(0); c.index_load(sr["prefix"])
q, wpq = mp.pack_queries(reads, lens, lopt)
c.batch_upload(q, lens, wpq)
il = max(1, int(lens[0::2].max()), int(lens[1::2].max()))
P = mp.default_params(nt2="nt2" in ini, insert_low=il, insert_high=750, max_read_length=lopt)
res = c.align_pairs(P)
print({k: v for k, v in res.items() if not hasattr(v, "__len__")})
names = sorted(set(int(a.split(b"\t")[0][2:]) for _, a, _ in bad[:12] if a.startswith(b"@p")))
for pid in names[:6]:
    print("pair", pid, "lens", lens[2 * pid], lens[2 * pid + 1])
    for arr in ("pairs", "rescued"):
        for r in res[arr][res[arr]["readID"] == 2 * pid]:
            print(" ", arr, {f: int(r[f]) for f in ("algnmt_1", "algnmt_2", "score_1", "score_2", "strand_1", "strand_2", "insertSize")},
                  mp.cigar_at(res["cigars"], int(r["cigar_1"])), mp.cigar_at(res["cigars"], int(r["cigar_2"])))
    for r in res["singles"][(res["singles"]["readID"] >> 1) == pid]:
        print("  single", {f: int(r[f]) for f in ("readID", "algnmt", "score", "strand", "seedAlignmentLength")}, mp.cigar_at(res["cigars"], int(r["cigar"])))
print("bounds", bounds)
