#!/usr/bin/env python3
"""bench.py -- read pairs aligned/sec of the soap4 alignment hot path on B200 (BASELINE.json metric).

One "step" = one pass of the whole hot path (MMP seeding -> SA resolution -> pairing -> deep DP with
traceback -> single-end DP -> mate rescue -> per-pair reduction) over one batch of synthetic read pairs,
through the C-ABI of libmegapath_b200.so (include/megapath_b200.h).

  value  : pairs/s with the batch already resident in HBM (mp_batch_upload outside the timed region),
           timed with CUDA events on the library's launching stream (mp_results.ms_total), max over ranks
  e2e    : pairs/s through the same C-ABI with HOST buffers: pinned-host -> HBM upload of the packed reads
           and the host-resident result arrays inside the timed region (wall clock between device syncs)
  roofline : the seeding kernel (k_mmp): algorithmic bytes = 64 B per occ evaluation + 16 B per LKT jump
            / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference's own soap4 (oracle/_ref/soap4, compiled from /root/reference by
           oracle/Makefile.ref) on a bounded sample of the same workload, all host cores

`--impl reference` times only that CPU arm.  Multi-GPU: one process per GPU (torchrun), index replicated,
disjoint read batches per rank, no collective on the data path ("weak" scaling).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
MAX_READ_LEN_OPT = 151          # soap4 -L 151: reads are truncated to L-1 = 150 (QueryParser.cpp:188)
INSERT_LO, INSERT_HI = 250, 500
INSERT_HIGH_OPT = 750           # soap4 -u 750


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-mbp", type=float, default=float(os.environ.get("MP_BENCH_REF_MBP", "3100")))
    ap.add_argument("--pairs-per-step", type=int, default=int(os.environ.get("MP_BENCH_PAIRS", str(1 << 20))))
    ap.add_argument("--cpu-sample-pairs", type=int, default=int(os.environ.get("MP_BENCH_CPU_PAIRS", "200000")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=int(os.environ.get("MP_BENCH_CONTEXTS", "3")),
                    help="contexts (host thread + stream each) per GPU sharing one resident index; batches alternate between them")
    ap.add_argument("--workdir", default=os.environ.get("MP_BENCH_DIR", "/tmp/mpbench"))
    ap.add_argument("--profile-step", action="store_true",
                    help="profiling aid (ncu --profile-from-start off): after the warm-up run ONE step on one context between "
                         "cudaProfilerStart/Stop and exit; prints no bench value")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload (SURVEY.md 8d cfg2 model: uniform i.i.d. ACGT reference in 24 sequences;
# FR pairs, insert U[250,500), substitutions per read from {0,0,0,1,1,2}, 1% unalignable pairs,
# 1% pairs with one random mate)
# ------------------------------------------------------------------------------------------------
def ref_bounds(n, nseq, seed):
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.choice(np.arange(1000, n - 1000), size=nseq - 1, replace=False))
    return np.concatenate([[0], cuts, [n]]).astype(np.int64)


def gen_ref_codes(n, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    CH = 1 << 28
    for o in range(0, n, CH):
        m = min(CH, n - o)
        out[o:o + m] = torch.randint(0, 4, (m,), dtype=torch.uint8, device=device, generator=g)
    return out


def gen_batch(ref_codes, bounds_t, npairs, seed, unalignable=0.01, one_random=0.01):
    """-> (codes (2*npairs, READ_LEN) uint8 on device, mate1 = even rows)."""
    import torch
    dev = ref_codes.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = READ_LEN
    nseq = bounds_t.numel() - 1
    isz = torch.randint(INSERT_LO, INSERT_HI, (npairs,), device=dev, generator=g)
    sid = torch.randint(0, nseq, (npairs,), device=dev, generator=g)
    lo = bounds_t[sid]
    hi = torch.maximum(bounds_t[sid + 1] - isz, lo + 1)
    start = lo + (torch.rand(npairs, device=dev, generator=g, dtype=torch.float64) * (hi - lo).double()).long()
    start = torch.clamp(start, 0, ref_codes.numel() - INSERT_HI - 1)
    ar = torch.arange(L, device=dev)
    out = torch.empty((2 * npairs, L), dtype=torch.uint8, device=dev)
    CH = 1 << 18
    for o in range(0, npairs, CH):
        s = start[o:o + CH]
        z = isz[o:o + CH]
        a = ref_codes[s[:, None] + ar[None, :]]
        b = 3 - ref_codes[(s + z)[:, None] - 1 - ar[None, :]]
        out[2 * o:2 * (o + len(s)):2] = a
        out[2 * o + 1:2 * (o + len(s)):2] = b
    nreads = 2 * npairs
    nsub = torch.tensor([0, 0, 0, 1, 1, 2], device=dev)[torch.randint(0, 6, (nreads,), device=dev, generator=g)]
    rows = torch.arange(nreads, device=dev)
    for k in (1, 2):
        sel = rows[nsub >= k]
        cols = torch.randint(0, L, (sel.numel(),), device=dev, generator=g)
        delta = torch.randint(1, 4, (sel.numel(),), device=dev, generator=g).to(torch.uint8)
        out[sel, cols] = (out[sel, cols] + delta) & 3
    u = torch.rand(npairs, device=dev, generator=g)
    ua = torch.nonzero(u < unalignable).flatten()
    if ua.numel():
        out[2 * ua] = torch.randint(0, 4, (ua.numel(), L), dtype=torch.uint8, device=dev, generator=g)
        out[2 * ua + 1] = torch.randint(0, 4, (ua.numel(), L), dtype=torch.uint8, device=dev, generator=g)
    one = torch.nonzero((u >= unalignable) & (u < unalignable + one_random)).flatten()
    if one.numel():
        out[2 * one + 1] = torch.randint(0, 4, (one.numel(), L), dtype=torch.uint8, device=dev, generator=g)
    return out


def pack_queries_torch(codes, max_len_opt):
    """appendToQueryArrays layout (QueryParser.cpp:184-203) built with torch on the device:
    2-bit, 16 bases per word LSB-first, 32-read interleaved.  -> pinned host int32 tensor, wpq."""
    import torch
    n, L = codes.shape
    wpq = (max_len_opt + 15) // 16
    npad = (n + 31) // 32 * 32
    c = torch.zeros((npad, wpq * 16), dtype=torch.int64, device=codes.device)
    c[:n, :L] = codes
    sh = (2 * (torch.arange(wpq * 16, device=codes.device) % 16))
    w = (c << sh[None, :]).view(npad, wpq, 16).sum(dim=2)
    w = w.view(npad // 32, 32, wpq).permute(0, 2, 1).contiguous().view(-1)
    w32 = (w & 0xFFFFFFFF).to(torch.int64)
    w32 = torch.where(w32 >= (1 << 31), w32 - (1 << 32), w32).to(torch.int32)
    host = torch.empty(w32.shape, dtype=torch.int32, pin_memory=True)
    host.copy_(w32)
    return host, wpq


def write_fastq_sample(path_prefix, codes_np):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    qual = b"I" * codes_np.shape[1]
    for mate in (0, 1):
        with open("%s_%d.fq" % (path_prefix, mate + 1), "wb") as f:
            rows = lut[codes_np[mate::2]]
            for i in range(rows.shape[0]):
                f.write(b"@p%d/%d\n" % (i, mate + 1) + rows[i].tobytes() + b"\n+\n" + qual + b"\n")


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("MP_BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("MP_BENCH_SAMPLE_MS", "200")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload cache: index files (reference format, so the reference binary runs on the same index)
# ------------------------------------------------------------------------------------------------
def workload_dir(args):
    n = int(args.ref_mbp * 1e6)
    return os.path.join(args.workdir, "ref%d" % n), n


def ensure_index(args, ctx, device, rank, world):
    """Builds the synthetic reference + FM-index once per box (rank 0), in HBM, with the library's own GPU
    builder (mp_index_build), and saves it in the reference's on-disk format; other ranks load the files."""
    import torch
    d, n = workload_dir(args)
    prefix = os.path.join(d, "ref.index")
    ready = os.path.join(d, "READY")
    nseq = 24
    bounds = ref_bounds(n, nseq, 42)
    ref_codes = gen_ref_codes(n, 42, device)
    t0 = time.time()
    if rank == 0 and not os.path.exists(ready):
        os.makedirs(d, exist_ok=True)
        ctx.index_build_codes(ref_codes, bounds, prefix)
        open(ready, "w").write("ok\n")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if not ctx.has_index():
        ctx.index_load(prefix)
    return prefix, ref_codes, torch.from_numpy(bounds).to(device), time.time() - t0


def run_reference_soap4(prefix, fq_prefix, out_prefix, threads):
    """-> (align_seconds, wall_seconds, stderr text); align = the reference's own
    'Overall alignment time (excl. read loading)' (SOAP4.cpp:613)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(ref_dir, "soap4")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/soap4 is not built (oracle/Makefile.ref)")
    cmd = [exe, "pair", prefix, fq_prefix + "_1.fq", fq_prefix + "_2.fq", "-o", out_prefix, "-C", os.path.join(ref_dir, "soap4.ini"),
           "-L", str(MAX_READ_LEN_OPT), "-T", str(threads), "-u", str(INSERT_HIGH_OPT), "-F", "-nc"]
    t0 = time.time()
    with open(out_prefix + ".stdout.fq", "wb") as fo:
        p = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, cwd=os.path.dirname(out_prefix))
    wall = time.time() - t0
    err = p.stderr.decode(errors="replace")
    if p.returncode != 0:
        raise RuntimeError("reference soap4 failed (%d): %s" % (p.returncode, err[-400:]))
    m = re.search(r"Overall alignment time \(excl\. read loading\)\s*:\s*([0-9.]+)", err)
    if not m:
        raise RuntimeError("could not parse the reference's timing line")
    try:
        os.remove(out_prefix + ".stdout.fq")
    except OSError:
        pass
    return float(m.group(1)), wall, err


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full summary."""
    best = None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))) if os.path.isdir(os.path.join(ROOT, "profiles")) else []:
        if not (name.startswith("r") and "ncu_full" in name and name.endswith(".txt")):
            continue
        cur, rd, wr = None, None, None
        for ln in open(os.path.join(ROOT, "profiles", name)):
            if ln.startswith("== "):
                cur, rd, wr = ln, None, None
            elif cur and kernel in cur:
                f = ln.split()
                if ln.startswith("dram__bytes_read.sum") and len(f) >= 3:
                    rd = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(f[2], 1.0)
                if ln.startswith("dram__bytes_write.sum") and len(f) >= 3:
                    wr = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(f[2], 1.0)
                if rd is not None and wr is not None:
                    best = {"bytes_per_launch": rd + wr, "source": "profiles/" + name}
                    cur = None
    return best


def clocks_hint():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        return 1965.0


def workload_name(args):
    n = int(args.ref_mbp * 1e6)
    return "%d x %d-pair batches of synthetic 150bp pairs (-L 151 -u 750 soap4.ini) vs %.0f Mbp synthetic reference" % (
        1, args.pairs_per_step, n / 1e6)


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: anything libraries print while we work (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return run(args, saved_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)


def emit(saved_stdout, obj):
    sys.stdout.flush()
    os.write(saved_stdout, (json.dumps(obj) + "\n").encode())


def run(args, saved_stdout):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    cfg = {"workload": "cfg2: 150bp pairs vs %.0f Mbp synthetic reference (24 seqs), %d pairs per step, soap4.ini -L 151 -u 750; "
                       "1%% unalignable pairs, 1%% one-mate-random" % (args.ref_mbp, args.pairs_per_step),
           "pairs_per_step_per_gpu": args.pairs_per_step, "ref_mbp": args.ref_mbp,
           "l2": "",
           "parallelism": "replicated index, disjoint read batches per GPU, no data-path collective"}

    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import megapath_b200 as mp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    ctx = mp.Context(local)
    prefix, ref_codes, bounds_t, t_index = ensure_index(args, ctx, device, rank, world if args.impl == "ours" else 1)
    info = ctx.index_info()
    cfg["l2"] = ("inputs larger than L2: %.1f GB HBM index gathered at random; every step also writes ~40 GB of traceback tables, "
                 "which flushes the 126 MB L2 between steps" % (info["hbmBytes"] / 1e9))
    cfg["contexts_per_gpu"] = max(1, args.contexts)
    d, _ = workload_dir(args)

    sample_codes = None
    if rank == 0 and (args.impl == "reference" or (world == 1 and not args.no_cpu_baseline)):
        sample_codes = gen_batch(ref_codes, bounds_t, args.cpu_sample_pairs, 7_000_003).cpu().numpy()

    def cpu_leg(npairs, tag):
        fqp = os.path.join(d, "sample_%s_%d" % (tag, npairs))
        if not os.path.exists(fqp + "_2.fq"):
            write_fastq_sample(fqp, sample_codes[:2 * npairs])
        align_s, wall_s, _ = run_reference_soap4(prefix, fqp, os.path.join(d, "refout_" + tag), ncores)
        return npairs / align_s, align_s, wall_s

    if args.impl == "reference":
        npairs = args.cpu_sample_pairs
        vals = []
        for i in range(args.warmup + args.steps):
            v, a, w = cpu_leg(npairs, "ref")
            if i >= args.warmup:
                vals.append((v, a))
        tot_align = sum(a for _, a in vals)
        value = npairs * len(vals) / tot_align
        cb = {"value": value, "unit": "pairs/s", "cores": ncores, "kind": "reference",
              "sample": "%d pairs per step of the same workload, oracle/_ref/soap4 -T %d, its own 'Overall alignment time (excl. read loading)'" % (npairs, ncores)}
        emit(saved_stdout, {"impl": "reference", "metric": "read pairs aligned/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_align / max(1, len(vals)),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u8", "data": "synthetic",
                          "config": cfg, "cpu_baseline": cb,
                          "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0

    # ---------------- our arm ----------------
    P = mp.default_params(insert_low=READ_LEN, insert_high=INSERT_HIGH_OPT, max_read_length=MAX_READ_LEN_OPT)
    nb = min(args.warmup + args.steps, 6)
    batches = []
    lens = np.full(2 * args.pairs_per_step, READ_LEN, dtype=np.uint32)
    for b in range(nb):
        codes = gen_batch(ref_codes, bounds_t, args.pairs_per_step, 1000 + 97 * rank + b)
        host, wpq = pack_queries_torch(codes, MAX_READ_LEN_OPT)
        batches.append(host)
        del codes
    del ref_codes
    torch.cuda.empty_cache()
    h2d = batches[0].numel() * 4 + lens.nbytes

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    nctx = max(1, args.contexts)
    ctxs = [ctx]

    def run_steps(c, ids, upload, acc_out):
        """steps `ids` on context c; upload=True: host buffers in (pinned -> HBM) every step"""
        for i in ids:
            if upload:
                c.batch_upload_ptr(batches[i % nb].data_ptr(), lens, wpq)
            t_call = time.perf_counter()
            s = c.align_pairs_summary(P)
            if os.environ.get("MP_BENCH_VERBOSE"):
                sys.stderr.write("  ctx %d step %d upload=%d: call %.1f ms (lib wall %.1f, seed %.1f, dp %.1f, fill %.1f, tb %.1f)\n" % (
                    ctxs.index(c) if c in ctxs else -1, i, int(upload), (time.perf_counter() - t_call) * 1e3, s["ms_wall"], s["ms_seed"], s["ms_dp"], s["ms_fill"], s["ms_tb"]))
            for k, v in s.items():
                acc_out[k] = acc_out.get(k, 0) + v

    # ---- warm-up (the K-mer filter is built once, before the other contexts borrow it) ----
    ctx.index_prepare(P)
    run_steps(ctx, range(min(1, args.warmup)), True, {})
    for _ in range(nctx - 1):
        ctxs.append(ctx.clone())
    for ci, c in enumerate(ctxs):
        run_steps(c, range(ci, ci + max(args.warmup - (1 if ci == 0 else 0), 1)), True, {})
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run_steps(ctx, [args.warmup], True, {})
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        emit(saved_stdout, {"profile_step": True, "note": "one step under the profiler; not a bench value"})
        return 0
    sampler = ClockSampler(local)
    sampler.start()

    # ---- loop R: one context, steps back to back: per-kernel device times for the roofline (no overlap between kernels) ----
    barrier()
    l0 = mp.launch_count()
    acc = {}
    run_steps(ctx, range(args.warmup, args.warmup + args.steps), True, acc)
    barrier()
    launches = mp.launch_count() - l0
    acc["n_fill_launches"] = 0
    if os.environ.get("MP_BENCH_VERBOSE"):
        sys.stderr.write("rank %d loop R: %s\n" % (rank, json.dumps(dict({k: round(v / args.steps, 3) for k, v in acc.items() if k.startswith("ms_")}, exact=acc.get("dp_tasks_exact", 0) / max(1, acc.get("dp_tasks", 1))))))

    def pipelined(upload):
        """K steps pulled from one queue by the contexts (one host thread each), so that a K that is not a multiple of the
        number of contexts leaves no context idle longer than one step; -> wall seconds between device syncs"""
        accs = [dict() for _ in ctxs]
        lock = threading.Lock()
        nxt = [args.warmup]

        def ids():
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= args.warmup + args.steps:
                    return
                yield i
        ths = [threading.Thread(target=run_steps, args=(c, ids(), upload, accs[ci])) for ci, c in enumerate(ctxs)]
        barrier()
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        barrier()
        return time.perf_counter() - t0, accs

    # ---- loop A (value): every context's batch is resident in HBM before the clock starts; each step re-runs the whole
    #      hot path on it (the 10 GB of traceback tables a step writes flush L2 between steps) ----
    for ci, c in enumerate(ctxs):
        c.batch_upload_ptr(batches[(args.warmup + ci) % nb].data_ptr(), lens, wpq)
    dev_s, accsA = pipelined(False)
    dev_ms = dev_s * 1e3
    pairs_aligned_A = sum(a.get("pairs_aligned", 0) for a in accsA)
    # ---- loop B (e2e): the same through host buffers: pinned-host -> HBM upload of every batch, results in host memory ----
    e2e_s, accsB = pipelined(True)
    d2h = sum(a.get("result_bytes", 0) for a in accsB) / max(1, args.steps)
    clocks = sampler.stop()

    from megapath_b200 import shard
    dev_ms_max, e2e_ms_max = shard.max_over_ranks([dev_ms, e2e_s * 1e3], device=device)     # slowest rank decides
    tot = shard.sum_counters({"pairs_aligned": pairs_aligned_A}, device=device)
    acc_pairs_all = tot["pairs_aligned"]
    total_pairs = args.pairs_per_step * args.steps * world
    value = total_pairs / (dev_ms_max / 1e3)
    e2e = total_pairs / (e2e_ms_max / 1e3)

    peak, peak_kind = measured_peak()
    # ---- roofline denominators MEASURED_PEAKS.json does not hold, measured now on this GPU ----
    try:
        gather32, gather64, dpx_peak = ctx.microbench(0), ctx.microbench(1), ctx.microbench(2)
    except Exception:
        gather32 = gather64 = dpx_peak = None
    steps = args.steps
    n_fill_launches = max(1, acc.get("n_fill_launches", 0))
    # dominant kernel: k_dp_fill (DP table fill).  Algorithmic bytes = one traceback byte written per DP cell.
    # cells the fill kernel really computed: tasks whose read occurs unchanged in its window are answered by the exact-occurrence
    # test (k_dp_exact, bit-identical results) and never reach it; dp_cells counts what the reference computes
    cells_filled = float(acc.get("dp_cells_filled") or acc["dp_cells"])
    fill_bytes = cells_filled
    fill_ach = fill_bytes / (acc["ms_fill"] / 1e3) / 1e9
    gcups_fill = cells_filled / (acc["ms_fill"] / 1e3) / 1e9
    gcups_dp = acc["dp_cells"] / ((acc["ms_fill"] + acc["ms_tb"] + acc.get("ms_exact", 0.0)) / 1e3) / 1e9
    # seeding kernel: bytes it must gather = 8-byte filter probes + 16-byte LKT pairs + 64-byte occ blocks
    # (two per backward-search step that is not a text-compare step) + 4-byte SA values + 1 text byte per compare
    occ_steps = max(0.0, (acc["n_occ"] - 2.0 * acc["n_text"]) / 2.0)
    seed_bytes = 8.0 * acc["n_probe"] + 16.0 * acc["n_lkt"] + 128.0 * occ_steps + 4.0 * acc["n_sa"] + 1.0 * acc["n_text"]
    seed_sectors = 32.0 * acc["n_probe"] + 32.0 * acc["n_lkt"] + 128.0 * occ_steps + 32.0 * acc["n_sa"] + 32.0 * acc["n_text"] / 32.0
    seed_ach = seed_bytes / (acc["ms_seed"] / 1e3) / 1e9
    traffic = ncu_traffic("k_dp_fill")
    roof = {"kernel": "k_dp_fill<5,-2,-3> (packed 16-bit DPX table fill, the kernel with the largest share of the step)",
            "bound": "hbm", "achieved": fill_ach, "peak": peak, "unit": "GB/s", "frac": fill_ach / peak,
            # DRAM bytes (read + write) of ONE k_dp_fill launch from the committed ncu --set full capture; the captured launch is the
            # first of a step: 2^18 left-leg tasks offered, about half of them answered by the exact-occurrence test, so it fills
            # ~4.1 G cells (= algorithmic bytes) and moves 4.55 GB of writes + 1.42 GB of write-allocate reads
            "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
            "peak_kind": peak_kind, "algorithmic": "1 traceback byte written per DP cell the kernel computes (SURVEY 8d cells = sum refLen*readLen over the tasks it is given)",
            "ms_per_step": acc["ms_fill"] / steps,
            "note": "integer-ALU bound, not bandwidth bound: ncu ALU pipe 90.8% busy, issue slots 79.7%, top stall math_pipe_throttle "
                    "(profiles/r01_ncu_full_v8_cfg2.txt); per-kernel times come from a single-context pass (loop R), value/e2e from the pipelined passes",
            "compute": {"gcups_fill": gcups_fill, "gcups_reference_equivalent_all_dp_kernels": gcups_dp,
                        "exact_occurrence_test": {"tasks_per_step": acc.get("dp_tasks_exact", 0) / steps, "of_tasks_per_step": acc["dp_tasks"] / steps,
                                                  "ms_per_step": acc.get("ms_exact", 0.0) / steps,
                                                  "cells_filled_per_step": cells_filled / steps, "cells_reference_per_step": acc["dp_cells"] / steps},
                        "dpx_peak_ginstr_s": dpx_peak, "dpx_instr_per_cell": 4.5,   # 9 packed min/max/add-max instructions per cell PAIR in the steady loop
                        "dpx_frac": (gcups_fill * 4.5 / dpx_peak) if dpx_peak else None,
                        # ncu smsp__inst_executed.sum of one launch x 32 lanes / (cells / 2): includes idle lanes of the wavefront ramps
                        "warp_instr_lane_slots_per_cell_pair": 42.6,
                        "issue_peak_gcups_at_that_instr_count": 148 * 4 * 32 * 2 * (clocks_hint() / 1e3) / 42.6},
            "seeding": {"kernel": "k_mmp", "ms_per_step": acc["ms_seed"] / steps, "bytes_gathered_gbs": seed_ach,
                        "sector_gbs": seed_sectors / (acc["ms_seed"] / 1e3) / 1e9, "gather32_peak_gbs": gather32, "gather64_peak_gbs": gather64,
                        "frac_of_gather32_peak": (seed_sectors / (acc["ms_seed"] / 1e3) / 1e9 / gather32) if gather32 else None,
                        "reference_algorithm_equiv_gbs": (64.0 * acc["n_occ"] + 16.0 * acc["n_lkt"] + 8.0 * acc["n_sa"]) / (acc["ms_seed"] / 1e3) / 1e9,
                        "counters_per_step": {k: acc[k] / steps for k in ("n_probe", "n_lkt", "n_occ", "n_sa", "n_text", "n_lf") if k in acc},
                        "note": "n_occ counts the occ evaluations the reference makes for the executed steps; starts rejected by the K-mer filter are not walked at all; "
                                "n_probe counts every filter probe issued, including re-probes that hit in L1"},
            "sa_lookup_gbs": (64.0 * acc["n_lf"] + 8.0 * acc["n_sa"]) / (acc["ms_sa"] / 1e3) / 1e9 if acc["ms_sa"] else None,
            "stage_ms_per_step": {k: acc[k] / steps for k in ("ms_seed", "ms_sa", "ms_pair", "ms_dp", "ms_fill", "ms_tb", "ms_exact", "ms_total", "ms_wall") if k in acc}}
    out = {"metric": "read pairs aligned/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u8 (integer DP, 2-bit FM-index)",
           "data": "synthetic", "config": cfg, "clocks": clocks,
           "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": int(launches), "roofline": roof,
           "aligned_fraction": acc_pairs_all / (args.pairs_per_step * args.steps * world),
           "index_prepare_s": t_index}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            v, a, w = cpu_leg(args.cpu_sample_pairs, "cpu")
            out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": ncores, "kind": "reference",
                                   "sample": "%d pairs of the same workload (same index files), oracle/_ref/soap4 -T %d, align loop %.1f s, wall %.1f s" % (
                                       args.cpu_sample_pairs, ncores, a, w)}
        except Exception as e:  # the baseline is reported, never the target: say why it is missing
            out["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": ncores, "kind": "reference", "sample": "unavailable: %s" % e}
    if rank == 0:
        emit(saved_stdout, out)
    for c in ctxs[1:]:
        c.close()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
