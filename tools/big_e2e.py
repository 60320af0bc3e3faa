#!/usr/bin/env python3
"""Multi-batch end-to-end check on the GPU box: > 1 Mi pairs (two driver batches) against a 20 Mbp reference built by the GPU
builder; the driver's stdout (1 GPU, and -G 2 when two GPUs are visible) must equal the reference binary's after the canonical sort."""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import megapath_b200 as mp
from conftest import canon_fastq, REF_DIR

d = "/tmp/mp_big"
os.makedirs(d, exist_ok=True)
n, npairs = 20_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_200_000
dev = torch.device("cuda", 0)
codes = bench.gen_ref_codes(n, 5, dev)
bounds = bench.ref_bounds(n, 12, 5)
ctx = mp.Context(0)
prefix = os.path.join(d, "ref.index")
ctx.index_build_codes(codes, bounds, prefix)
ctx.close()
reads = bench.gen_batch(codes, torch.from_numpy(bounds).to(dev), npairs, 99, unalignable=0.02, one_random=0.03).cpu().numpy()
t0 = time.time()
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
for mate in (0, 1):
    rows = lut[reads[mate::2]]
    L = rows.shape[1]
    names = np.char.add(np.char.add("@p", np.arange(npairs).astype(str)), "/%d" % (mate + 1)).astype("S")
    with open(os.path.join(d, "r_%d.fq" % (mate + 1)), "wb") as f:
        qual = b"I" * L
        for i in range(npairs):
            f.write(names[i] + b"\n" + rows[i].tobytes() + b"\n+\n" + qual + b"\n")
print("fastq written in %.1f s" % (time.time() - t0), flush=True)
fq1, fq2 = os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")
t0 = time.time()
with open(os.path.join(d, "ref.out"), "wb") as fo:
    subprocess.check_call([os.path.join(REF_DIR, "soap4"), "pair", prefix, fq1, fq2, "-o", os.path.join(d, "refo"), "-C", os.path.join(REF_DIR, "soap4.ini"),
                           "-L", "151", "-T", str(os.cpu_count()), "-u", "750", "-F", "-nc"], stdout=fo, stderr=subprocess.DEVNULL, cwd=d)
t_ref = time.time() - t0
want = hashlib.md5(canon_fastq(open(os.path.join(d, "ref.out"), "rb").read())).hexdigest()
print("reference: %.1f s (%d pairs/s)" % (t_ref, npairs / t_ref), flush=True)
exe = os.path.join(ROOT, "megapath_b200", "bin", "soap4")
ini = os.path.join(ROOT, "megapath_b200", "ini", "soap4.ini")
for g in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
    t0 = time.time()
    p = subprocess.run([exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", "4", "-u", "750", "-F", "-nc", "-G", str(g)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    t_our = time.time() - t0
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    got = hashlib.md5(canon_fastq(p.stdout)).hexdigest()
    tail = [l for l in p.stderr.decode().splitlines() if "Overall" in l or "Loading time" in l]
    print("ours -G %d: %.1f s wall (%d pairs/s incl. index load and FASTQ parsing) md5 %s %s" % (g, t_our, npairs / t_our, got, "OK" if got == want else "MISMATCH"), tail, flush=True)
    assert got == want
# throughput of the driver itself: stdout to a file, all host threads
for T in (4, os.cpu_count()):
    t0 = time.time()
    with open(os.path.join(d, "our.out"), "wb") as fo:
        p = subprocess.run([exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", str(T), "-u", "750", "-F", "-nc"],
                           stdout=fo, stderr=subprocess.PIPE, timeout=600, env=dict(os.environ, MP_DRIVER_TIMING="1"))
    t_our = time.time() - t0
    lines = p.stderr.decode().splitlines()
    tail = [l for l in lines if "Overall" in l or "Loading time" in l or "Elapsed time" in l or "[timing]" in l]
    print("ours -T %d to a file: %.1f s wall" % (T, t_our), tail, flush=True)
print("big e2e ok")
