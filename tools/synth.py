#!/usr/bin/env python3
"""Seeded synthetic references and read pairs for the soap4 hot path.

Everything is ACGT-only unless --n-rate is given (the reference loader maps N to G,
soap4/IndexHandler.cpp:41-45).  Reads are FR pairs: mate 1 forward at the fragment start,
mate 2 reverse-complemented at the fragment end.

Usage:
  synth.py ref   --out ref.fa --len 2000000 --nseq 10 --seed 42
  synth.py reads --ref ref.fa --out-prefix r --pairs 20000 --len 150 --seed 7 [--model clean|divergent]
"""
import argparse
import sys
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


def make_ref(total_len, nseq, seed, repeat_frac=0.0):
    """Uniform i.i.d. ACGT multi-FASTA; optional planted near-duplicate blocks."""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.choice(np.arange(1000, total_len - 1000), size=nseq - 1, replace=False)) if nseq > 1 else np.array([], dtype=np.int64)
    bounds = np.concatenate([[0], cuts, [total_len]]).astype(np.int64)
    seq = ALPHA[rng.integers(0, 4, size=total_len, dtype=np.uint8)]
    if repeat_frac > 0:
        nrep = int(total_len * repeat_frac / 2000)
        for _ in range(nrep):
            ln = int(rng.integers(300, 2000))
            src = int(rng.integers(0, total_len - ln))
            dst = int(rng.integers(0, total_len - ln))
            blk = seq[src:src + ln].copy()
            nmut = int(rng.integers(0, max(1, ln // 100)))
            if nmut:
                p = rng.integers(0, ln, size=nmut)
                blk[p] = ALPHA[rng.integers(0, 4, size=nmut)]
            seq[dst:dst + ln] = blk
    return seq, bounds


def write_fasta(path, seq, bounds, width=60):
    with open(path, "wb") as f:
        for i in range(len(bounds) - 1):
            s = seq[bounds[i]:bounds[i + 1]]
            f.write(b">seq%d synthetic\n" % (i + 1))
            n = len(s)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def read_fasta(path):
    seqs = []
    cur = []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur))
                    cur = []
            else:
                cur.append(line.strip())
    if cur:
        seqs.append(b"".join(cur))
    arrs = [np.frombuffer(s, dtype=np.uint8) for s in seqs]
    bounds = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
    return np.concatenate(arrs), bounds


def _mutate_indel(frag, rng, sub, ins, dele):
    out = []
    i = 0
    n = len(frag)
    while i < n:
        r = rng.random()
        if r < dele:
            i += int(rng.integers(1, 4))
            continue
        if r < dele + ins:
            k = int(rng.integers(1, 4))
            out.extend(ALPHA[rng.integers(0, 4, size=k)].tolist())
        c = frag[i]
        if rng.random() < sub:
            c = ALPHA[(np.searchsorted(ALPHA, c) + int(rng.integers(1, 4))) % 4]
        out.append(int(c))
        i += 1
    return np.array(out, dtype=np.uint8)


def make_pairs(seq, bounds, npairs, rlen, seed, model="clean", ins_lo=250, ins_hi=500,
               unalignable=0.0, one_random=0.0, varlen=False, n_rate=0.0, span_frac=0.0):
    """Returns (reads1, reads2) as lists of uint8 arrays (mate 2 already reverse-complemented)."""
    rng = np.random.default_rng(seed)
    total = len(seq)
    nseq = len(bounds) - 1
    r1, r2 = [], []
    if model == "clean" and not varlen:
        # vectorised path
        isz = rng.integers(max(ins_lo, rlen), max(ins_hi, rlen + 1), size=npairs)
        sid = rng.integers(0, nseq, size=npairs)
        lo = bounds[sid]
        hi = bounds[sid + 1] - isz
        bad = hi <= lo
        hi = np.where(bad, lo + 1, hi)
        start = lo + (rng.random(npairs) * (hi - lo)).astype(np.int64)
        if span_frac > 0:
            sp = rng.random(npairs) < span_frac
            start = np.where(sp, np.clip(bounds[np.minimum(sid + 1, nseq - 1)] - rlen // 2, 0, total - isz - 1), start)
        start = np.clip(start, 0, total - isz - 1)
        idx = start[:, None] + np.arange(rlen)[None, :]
        a = seq[idx]
        idx2 = (start + isz)[:, None] - 1 - np.arange(rlen)[None, :]
        b = COMP[seq[idx2]]
        for arr in (a, b):
            nsub = rng.choice(np.array([0, 0, 0, 1, 1, 2]), size=npairs)
            for k in (1, 2):
                rows = np.nonzero(nsub >= k)[0]
                cols = rng.integers(0, rlen, size=len(rows))
                old = arr[rows, cols]
                arr[rows, cols] = ALPHA[(np.searchsorted(ALPHA, old) + rng.integers(1, 4, size=len(rows))) % 4]
        rnd = rng.random(npairs)
        ua = rnd < unalignable
        a[ua] = ALPHA[rng.integers(0, 4, size=(int(ua.sum()), rlen))]
        b[ua] = ALPHA[rng.integers(0, 4, size=(int(ua.sum()), rlen))]
        one = (rnd >= unalignable) & (rnd < unalignable + one_random)
        b[one] = ALPHA[rng.integers(0, 4, size=(int(one.sum()), rlen))]
        if n_rate > 0:
            for arr in (a, b):
                m = rng.random(arr.shape) < n_rate
                arr[m] = ord("N")
        return list(a), list(b)
    sub, ins, dele = (0.04, 0.005, 0.005) if model == "divergent" else (0.0, 0.0, 0.0)
    for _ in range(npairs):
        l1 = int(rng.integers(50, rlen + 1)) if varlen else rlen
        l2 = int(rng.integers(50, rlen + 1)) if varlen else rlen
        isz = int(rng.integers(max(ins_lo, max(l1, l2) + 10), max(ins_hi, max(l1, l2) + 11)))
        sid = int(rng.integers(0, nseq))
        lo, hi = int(bounds[sid]), int(bounds[sid + 1]) - isz - 16
        if hi <= lo:
            lo, hi = 0, total - isz - 16
        st = int(rng.integers(lo, hi))
        f1 = seq[st:st + l1 + 12]
        f2 = COMP[seq[st + isz - l2 - 12:st + isz][::-1]]
        if model == "divergent":
            f1 = _mutate_indel(f1, rng, sub, ins, dele)
            f2 = _mutate_indel(f2, rng, sub, ins, dele)
        else:
            f1 = f1.copy()
            f2 = f2.copy()
            for f in (f1, f2):
                for _k in range(int(rng.choice([0, 0, 0, 1, 1, 2]))):
                    p = int(rng.integers(0, len(f)))
                    f[p] = ALPHA[(np.searchsorted(ALPHA, f[p]) + int(rng.integers(1, 4))) % 4]
        f1 = f1[:l1]
        f2 = f2[:l2]
        u = rng.random()
        if u < unalignable:
            f1 = ALPHA[rng.integers(0, 4, size=l1)]
            f2 = ALPHA[rng.integers(0, 4, size=l2)]
        elif u < unalignable + one_random:
            if rng.random() < 0.5:
                f2 = ALPHA[rng.integers(0, 4, size=l2)]
            else:
                f1 = ALPHA[rng.integers(0, 4, size=l1)]
        if n_rate > 0:
            for f in (f1, f2):
                m = rng.random(len(f)) < n_rate
                f[m] = ord("N")
        r1.append(f1)
        r2.append(f2)
    return r1, r2


def write_fastq(path, reads, mate, qual=b"I", prefix=b"p"):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            s = r.tobytes()
            f.write(b"@" + prefix + b"%d/%d\n" % (i, mate) + s + b"\n+\n" + qual * len(s) + b"\n")


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("ref")
    a.add_argument("--out", required=True)
    a.add_argument("--len", type=int, required=True)
    a.add_argument("--nseq", type=int, default=10)
    a.add_argument("--seed", type=int, default=42)
    a.add_argument("--repeat-frac", type=float, default=0.0)
    b = sub.add_parser("reads")
    b.add_argument("--ref", required=True)
    b.add_argument("--out-prefix", required=True)
    b.add_argument("--pairs", type=int, required=True)
    b.add_argument("--len", type=int, default=150)
    b.add_argument("--seed", type=int, default=7)
    b.add_argument("--model", default="clean", choices=["clean", "divergent"])
    b.add_argument("--unalignable", type=float, default=0.0)
    b.add_argument("--one-random", type=float, default=0.0)
    b.add_argument("--varlen", action="store_true")
    b.add_argument("--n-rate", type=float, default=0.0)
    b.add_argument("--span-frac", type=float, default=0.0)
    args = ap.parse_args(argv)
    if args.cmd == "ref":
        seq, bounds = make_ref(args.len, args.nseq, args.seed, args.repeat_frac)
        write_fasta(args.out, seq, bounds)
    else:
        seq, bounds = read_fasta(args.ref)
        r1, r2 = make_pairs(seq, bounds, args.pairs, args.len, args.seed, args.model,
                            unalignable=args.unalignable, one_random=args.one_random,
                            varlen=args.varlen, n_rate=args.n_rate, span_frac=args.span_frac)
        write_fastq(args.out_prefix + "_1.fq", r1, 1)
        write_fastq(args.out_prefix + "_2.fq", r2, 2)
    return 0


if __name__ == "__main__":
    sys.exit(main())
