#!/usr/bin/env python3
"""Steady-state throughput of the soap4 driver itself (FASTQ files in, annotated FASTQ out) on the GPU box:
20 Mbp GPU-built reference, 1.2 M synthetic pairs replicated 8x (9.6 M pairs, ~3 GB per mate file)."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import megapath_b200 as mp

d = "/tmp/mp_thr"
os.makedirs(d, exist_ok=True)
n, npairs, rep = 20_000_000, 1_200_000, int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
codes = bench.gen_ref_codes(n, 5, dev)
bounds = bench.ref_bounds(n, 12, 5)
ctx = mp.Context(0)
prefix = os.path.join(d, "ref.index")
ctx.index_build_codes(codes, bounds, prefix)
ctx.close()
reads = bench.gen_batch(codes, torch.from_numpy(bounds).to(dev), npairs, 99, unalignable=0.02, one_random=0.03).cpu().numpy()
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
for mate in (0, 1):
    rows = lut[reads[mate::2]]
    L = rows.shape[1]
    names = np.char.add(np.char.add("@p", np.arange(npairs).astype(str)), "/%d" % (mate + 1)).astype("S")
    qual = b"I" * L
    blob = b"".join(names[i] + b"\n" + rows[i].tobytes() + b"\n+\n" + qual + b"\n" for i in range(npairs))
    with open(os.path.join(d, "r_%d.fq" % (mate + 1)), "wb") as f:
        for _ in range(rep):
            f.write(blob)
fq1, fq2 = os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")
os.sync()          # the read files were written a moment ago: without this their write-back runs beside the first timed run
exe = os.path.join(ROOT, "megapath_b200", "bin", "soap4")
ini = os.path.join(ROOT, "megapath_b200", "ini", "soap4.ini")
total = npairs * rep
if os.environ.get("MP_THR_NCU"):
    # launch list of the ingest / egress kernels of one run of the driver (per-launch times under ncu are cold-cache and serialised)
    out = os.path.join(ROOT, "gpurun_out", "launches_cli_io.csv")
    with open("/dev/null", "wb") as fo:
        subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:k_fq|k_fmt", "-c", "60", "--csv", "--log-file", out,
                        exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", "16", "-u", "750", "-F", "-nc"],
                       stdout=fo, stderr=subprocess.DEVNULL, timeout=900, env=dict(os.environ, MP_CONTEXTS_PER_GPU="1"))
    print(open(out).read()[-3000:])
    sys.exit(0)
if os.environ.get("MP_THR_LSAM"):
    # -lsam (the fused fastq2lsam consumer) with stdout in a regular file: the text a consumer has to read shrinks, so the page-cache
    # write of the annotated FASTQ no longer bounds the loop
    for mode in ("0", "1"):
        sink = os.path.join(d, "our.lsam")
        with open(sink, "wb") as fo:
            p = subprocess.run([exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", "16", "-u", "750", "-F", "-nc", "-lsam", mode],
                               stdout=fo, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, MP_DRIVER_TIMING="1"))
        lines = p.stderr.decode().splitlines()
        loop = [float(l.split(":")[1].split()[0]) for l in lines if "Overall alignment time" in l][0]
        print("-lsam %s to a file (%d MB): loop %.3f s = %.2f M pairs/s; %s" % (mode, os.path.getsize(sink) >> 20, loop, total / loop / 1e6,
                                                                                 [l for l in lines if "formatting on the" in l]), flush=True)
    sys.exit(0)
if os.environ.get("MP_THR_SWEEP"):
    for ctxs, st, extra in ((3, 16, {}), (3, 8, {}), (3, 4, {}), (3, 4, {}), (3, 2, {}), (4, 4, {}), (2, 4, {}), (3, 8, {})):
        with open("/dev/null", "wb") as fo:
            p = subprocess.run([exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", "16", "-u", "750", "-F", "-nc"],
                               stdout=fo, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, MP_CONTEXTS_PER_GPU=str(ctxs), MP_STAGE_THREADS=str(st), **extra))
        lines = p.stderr.decode().splitlines()
        loop = [float(l.split(":")[1].split()[0]) for l in lines if "Overall alignment time" in l][0]
        print("contexts %d, stage threads %d %s: loop %.3f s = %.2f M pairs/s" % (ctxs, st, extra, loop, total / loop / 1e6), flush=True)
    sys.exit(0)
cfgs = ((16, 3, "/dev/null", 0),) if os.environ.get("MP_THR_ALL") else ((16, 3, "/dev/null", 0), (16, 3, os.path.join(d, "our.out"), 0), (16, 3, "/dev/null", 1), (16, 3, os.path.join(d, "our.out"), 1))
for T, ctxs, sink, host in cfgs:
    t0 = time.time()
    with open(sink, "wb") as fo:
        p = subprocess.run([exe, "pair", prefix, fq1, fq2, "-o", os.path.join(d, "ouro"), "-C", ini, "-L", "151", "-T", str(T), "-u", "750", "-F", "-nc"],
                           stdout=fo, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, MP_CONTEXTS_PER_GPU=str(ctxs), MP_DRIVER_TIMING="1", **({"MP_HOST_IO": "1"} if host else {}), **({"MP_TRACE": "2"} if os.environ.get("MP_THR_TRACE") else {})))
    wall = time.time() - t0
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = p.stderr.decode().splitlines()
    loop = [float(l.split(":")[1].split()[0]) for l in lines if "Overall alignment time" in l][0]
    load = [l for l in lines if "Elapsed time on host" in l]
    tim = [l for l in lines if "[timing]" in l]
    if os.environ.get("MP_THR_ALL"):
        print("\n".join(l for l in lines if "[timing]" in l or "Elapsed time on host" in l), flush=True)
    if os.environ.get("MP_THR_TRACE"):
        print("\n".join([l for l in lines if "mp_trace" in l or "[timing]" in l][:70]), flush=True)
    print([l for l in lines if "formatting on the" in l])
    print("-T %d, %d contexts, out=%s: wall %.1f s; batch loop %.2f s = %.2f M pairs/s; reader per batch %s; last batches %s" % (
        T, ctxs, sink, wall, loop, total / loop / 1e6, [l.split(":")[1].split()[0] for l in load[-4:-1]], tim[-4:]), flush=True)
