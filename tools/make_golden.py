#!/usr/bin/env python3
"""Generates tests/golden/ from the reference itself (oracle/_ref, built by oracle/Makefile.ref from
/root/reference/soap4).  Run here (the reference sources exist only in this container); the outputs are
small and committed so that the oracle and the CUDA path are also checked where oracle/_ref is absent.

  idx.*            index of a 30 kbp synthetic reference written by the reference's 2bwt-builder
                   (.lkt omitted: 512 MiB; the oracle rebuilds it from .pac, the CUDA path builds its own index)
  ref.fa           the FASTA it was built from
  <set>_1.fq/_2.fq read pairs
  <set>.npz        seam dumps of soap4_dump: SeedPos arrays, CandidateInfo list, DP task records (first 400),
                   and the canonically sorted stdout FASTQ of the reference binary
  prim.npz         BWTOccValue / BWTSaValue / LT values of the reference (libref_bwt.so) at fixed queries
  dp.npz           callDP outputs of the reference (libref_dp.so) for random tasks of three table shapes
"""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po          # noqa: E402
from tools import synth                    # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden")
SETS = [("clean", 150, 151, dict(model="clean", varlen=True, n_rate=0.002), "soap4.ini"),
        ("div", 100, 101, dict(model="divergent", one_random=0.08, unalignable=0.03), "soap4.ini"),
        ("nt2", 150, 151, dict(model="divergent", one_random=0.05), "soap4-nt2.ini")]


def canon_fastq(path):
    recs = []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        recs.append(b"\n".join(lines[i:i + 4]))
    # records come in mate1, mate2 adjacent order per pair; sort pairs by name keeping mate order
    pairs = [(recs[i], recs[i + 1]) for i in range(0, len(recs) - 1, 2)]
    pairs.sort(key=lambda p: p[0].split(b"\t")[0])
    return b"\n".join(a + b"\n" + b for a, b in pairs) + b"\n"


def main():
    from test_oracle_vs_ref import random_dp_tasks
    os.makedirs(OUT, exist_ok=True)
    work = "/tmp/mp_golden"
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    seq, bounds = synth.make_ref(30000, 3, seed=123, repeat_frac=0.10)
    fa = os.path.join(work, "ref.fa")
    synth.write_fasta(fa, seq, bounds)
    shutil.copy(os.path.join(REF, "2bwt-builder.ini"), work)
    subprocess.check_call([os.path.join(REF, "2bwt-builder"), fa], cwd=work, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    shutil.copy(fa, os.path.join(OUT, "ref.fa"))
    for ext in ("bwt", "fmv", "sa", "pac", "ann", "amb", "tra"):
        shutil.copy(fa + ".index." + ext, os.path.join(OUT, "idx." + ext))
    prefix = fa + ".index"
    # ---- primitives ----
    rx = po.RefIndex(prefix)
    rng = np.random.default_rng(5)
    idx = np.concatenate([rng.integers(0, rx.n + 2, size=3000), np.arange(0, 600), [rx.n, rx.n + 1]]).astype(np.uint64)
    c = rng.integers(0, 4, size=len(idx)).astype(np.uint32)
    sidx = np.concatenate([rng.integers(0, rx.n + 1, size=2000), [0, rx.n]]).astype(np.uint64)
    keys = rng.integers(0, 4 ** 13, size=3000).astype(np.uint32)
    l, r = rx.lkt(keys)
    np.savez_compressed(os.path.join(OUT, "prim.npz"), idx=idx, c=c, occ=rx.occ(idx, c), sidx=sidx, sa=rx.sa(sidx), keys=keys, l=l, r=r)
    # ---- DP ----
    dp = {}
    for k, (maxdna, maxread, clips) in enumerate([(220, 152, (130, 130)), (752, 152, (130, 130)), (320, 252, (10, 20))]):
        rg = np.random.default_rng(100 + k)
        refs, dl, reads, rl = random_dp_tasks(rg, 96, maxdna, maxread)
        sc, hl, mc, pats = po.ref_dp(refs, dl, reads, rl, maxdna, maxread, clips[0], clips[1])
        plen = np.array([len(po.pattern_bytes(p)) if s >= po.dp_cutoff(int(L)) else 0 for p, s, L in zip(pats, sc, rl)], dtype=np.int32)
        for name, a in (("refs", refs), ("dl", dl), ("reads", reads), ("rl", rl), ("sc", sc), ("hl", hl), ("mc", mc), ("pats", pats), ("plen", plen),
                        ("shape", np.array([maxdna, maxread, clips[0], clips[1]]))):
            dp["%s%d" % (name, k)] = a
    np.savez_compressed(os.path.join(OUT, "dp.npz"), **dp)
    # ---- read sets through the instrumented reference ----
    for name, rlen, lopt, kw, ini in SETS:
        r1, r2 = synth.make_pairs(seq, bounds, 160, rlen, seed=77, **kw)
        p = os.path.join(OUT, name)
        synth.write_fastq(p + "_1.fq", r1, 1)
        synth.write_fastq(p + "_2.fq", r2, 2)
        dump = os.path.join(work, name + ".dump")
        os.makedirs(dump)
        env = dict(os.environ, MPH_DUMP_DIR=dump)
        outfq = os.path.join(work, name + ".out.fq")
        with open(outfq, "wb") as fo:
            subprocess.check_call([os.path.join(REF, "soap4_dump"), "pair", prefix, p + "_1.fq", p + "_2.fq", "-o", os.path.join(work, name),
                                   "-C", os.path.join(REF, ini), "-L", str(lopt), "-T", "2", "-u", "750", "-F", "-nc"],
                                  stdout=fo, stderr=subprocess.DEVNULL, env=env, cwd=work)
        (rp, mp), = po.read_seedpos_dump(dump + "/seedpos.bin")
        cand, = po.read_cand_dump(dump + "/cand.bin")
        tasks = []
        for rec in po.read_dp_dump(dump + "/dp.bin"):
            for t in rec["tasks"]:
                if len(tasks) < 400:
                    tasks.append((rec["clip_lt"], rec["clip_rt"], rec["mismatch"], rec["gap_open"]) + t)
        blob = {"readPos": rp, "matePos": mp, "cand": cand, "lopt": np.array([lopt]), "nt2": np.array([int(ini != "soap4.ini")]),
                "fastq": np.frombuffer(canon_fastq(outfq), dtype=np.uint8),
                "dp_meta": np.array([[t[0], t[1], t[2], t[3], len(t[4]), len(t[5]), t[6], t[7], t[8], t[9], len(t[10])] for t in tasks], dtype=np.int64),
                "dp_bytes": np.frombuffer(b"".join(bytes(t[4]) + bytes(t[5]) + bytes(t[10]) for t in tasks), dtype=np.uint8)}
        np.savez_compressed(p + ".npz", **blob)
    subprocess.call(["du", "-sh", OUT])


if __name__ == "__main__":
    main()
