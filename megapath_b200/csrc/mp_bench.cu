// mp_bench.cu -- micro-benchmarks for the two roofline denominators MEASURED_PEAKS.json does not hold
// (SURVEY.md 8d): the random-gather ceiling of HBM (32-byte and 64-byte granules over a multi-GB table,
// millions of independent requests in flight) and the issue peak of the packed 16-bit DPX instructions.
// Called by bench.py through mp_microbench(); nothing here is on the product path.
#include "mp_context.h"

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
// every thread issues `rounds` x 8 independent loads at pseudo-random granule addresses
// POLICY: how the gather is issued -- 0 ld.global.nc (__ldg), 1 ld.global.cg, 2 ld.global.cs, 3 ld.global.nc.L1::no_allocate, 4 ld.global.nc.L1::no_allocate.L2::64B... see load16
template <int POLICY>
__device__ __forceinline__ uint4 load16(const uint4 *p)
{
    if (POLICY == 1) return __ldcg(p);
    if (POLICY == 2) return __ldcs(p);
    if (POLICY == 3) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
    if (POLICY == 4) { uint4 v; asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
    if (POLICY == 5) return __ldlu(p);
    if (POLICY == 6) return __ldcv(p);
    return __ldg(p);
}
template <int GRANULE, int POLICY = 0>
__global__ void k_gather_bench(const uint4 *__restrict__ table, uint64_t nGranules, int rounds, uint32_t *__restrict__ sink)
{
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        uint4 v[8], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = __umul64hi(mix64(tid * 8191u + (uint64_t)r * 8 + u + 1), nGranules);
            const uint4 *p = table + g * (GRANULE / 16);
            v[u] = load16<POLICY>(p);
            if (GRANULE == 64) w[u] = load16<POLICY>(p + 2); else w[u] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].w ^ w[u].y;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
// 8 independent chains of VIADDMNMX.U16x2 + VIMNMX3.U16x2 + VIMNMX.U16x2 per thread
__global__ void k_dpx_bench(int iters, uint32_t seed, uint32_t *__restrict__ sink)
{
    uint32_t a[8], b = seed * 0x10001u + threadIdx.x, c = 0x00030003u;
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed + u * 0x00010001u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a[u] = __viaddmax_u16x2(a[u], 0xFFFFFFFFu, b);
            a[u] = __vimax3_u16x2(a[u], b, c);
            a[u] = __vminu2(a[u], 0x7FFF7FFFu);
        }
        b += 0x00010001u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) r ^= a[u];
    if (r == 0x12345678u) sink[0] = r;
}

// kind 0: random 32-byte gathers, 1: random 64-byte gathers -> result = GB/s of requested bytes
// kind 2: packed 16-bit DPX instruction issue rate -> result = 1e9 thread-instructions per second (two 16-bit lanes each)
extern "C" int mp_microbench(mp_context *ctx, int kind, double *result)
{
    if (!ctx || !result) { mp_set_error("mp_microbench: null argument"); return MP_ERR_ARG; }
    MP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    DevBuf sink; if (sink.reserve(64)) return MP_ERR_CUDA;
    float ms = 0;
    if (kind == 0 || kind == 1 || (kind >= 10 && kind <= 29)) {
        const size_t bytes = (size_t)8 << 30;
        DevBuf table; if (table.reserve(bytes)) return MP_ERR_CUDA;
        MP_CUDA(cudaMemsetAsync(table.p, 1, bytes, st));
        // kinds 10+p / 20+p: the 32- / 64-byte gather issued with load policy p (see load16), to choose the policy of the index gathers
        const int granule = (kind == 0 || (kind >= 10 && kind < 20)) ? 32 : 64, rounds = 16;
        const int policy = kind >= 20 ? kind - 20 : kind >= 10 ? kind - 10 : 0;
        const uint64_t nGran = bytes / granule;
        const unsigned blocks = 148 * 64, threads = 256;
        for (int rep = 0; rep < 3; ++rep) {
            MP_CUDA(cudaEventRecord(e0, st));
#define GB(G, PL) k_gather_bench<G, PL><<<blocks, threads, 0, st>>>(table.as<uint4>(), nGran, rounds, sink.as<uint32_t>())
#define GBP(G) do { switch (policy) { case 1: GB(G, 1); break; case 2: GB(G, 2); break; case 3: GB(G, 3); break; case 4: GB(G, 4); break; \
                                      case 5: GB(G, 5); break; case 6: GB(G, 6); break; default: GB(G, 0); } } while (0)
            if (granule == 32) GBP(32); else GBP(64);
#undef GBP
#undef GB
            MP_CUDA(cudaEventRecord(e1, st));
            MP_CUDA(cudaStreamSynchronize(st));
            MP_CUDA(cudaGetLastError());
            float t; cudaEventElapsedTime(&t, e0, e1); if (rep == 0 || t < ms) ms = t;
        }
        *result = (double)blocks * threads * rounds * 8 * granule / (ms * 1e-3) / 1e9;
        table.release();
    } else if (kind == 2) {
        const int iters = 4096; const unsigned blocks = 148 * 16, threads = 256;
        for (int rep = 0; rep < 3; ++rep) {
            MP_CUDA(cudaEventRecord(e0, st));
            k_dpx_bench<<<blocks, threads, 0, st>>>(iters, 7u + rep, sink.as<uint32_t>());
            MP_CUDA(cudaEventRecord(e1, st));
            MP_CUDA(cudaStreamSynchronize(st));
            MP_CUDA(cudaGetLastError());
            float t; cudaEventElapsedTime(&t, e0, e1); if (rep == 0 || t < ms) ms = t;
        }
        *result = (double)blocks * threads * iters * 8 * 3 / (ms * 1e-3) / 1e9;
    } else { mp_set_error("mp_microbench: unknown kind %d", kind); sink.release(); return MP_ERR_ARG; }
    sink.release();
    return 0;
}
