// gather_bench3.cu -- does any load PATH on B200 fetch a random 64-byte row without paying for the whole 128-byte
// line?  gather_bench2 showed that every ld.global flavour moves 2x the useful bytes through L2 and DRAM and that
// the ceiling is ~37 G random rows/s for rows of 32, 64 and 128 bytes alike.  Here the same random 64-byte rows are
// fetched with
//   ldg   : ld.global.nc.L1::no_allocate.v4 (4 lanes per row)               -- the path kernel #2 uses
//   ldgsts: cp.async.cg.shared.global 16 B (4 lanes per row)                -- LDGSTS, bypasses registers
//   bulk  : cp.async.bulk.shared::cluster.global.mbarrier (one 64 B copy per lane) -- UBLKCP, the TMA unit
// Output: JSON, useful GB/s.  Run under ncu with dram__bytes_read.sum to see the fetch granularity per path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench3 gather_bench3.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int WARPS = 4;

template <int ROW>
__global__ void __launch_bounds__(32 * WARPS) gather_ldg(const uint8_t *tab, uint64_t n_rows, uint32_t iters, uint32_t *sink)
{
    constexpr int LPR = ROW / 16, G = 32 / LPR, U = 4;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, col = lane % LPR;
    const uint64_t wid = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it)
    {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const uint64_t key = mix(wid * 0x9e3779b97f4a7c15ULL + ((uint64_t)it * U + u) * G + sub);
            const uint64_t row = (uint64_t)(((__uint128_t)key * n_rows) >> 64);
            const uint8_t *p = tab + row * ROW + col * 16;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678)
        *sink = acc;
}

// LDGSTS: STAGES groups in flight, each G rows (16 B per lane)
template <int ROW, int STAGES>
__global__ void __launch_bounds__(32 * WARPS) gather_ldgsts(const uint8_t *tab, uint64_t n_rows, uint32_t iters, uint32_t *sink)
{
    constexpr int LPR = ROW / 16, G = 32 / LPR;
    __shared__ alignas(16) uint4 buf[WARPS][STAGES][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / LPR, col = lane % LPR;
    const uint64_t wid = (uint64_t)blockIdx.x * WARPS + wib;
    uint32_t acc = 0;
    auto issue = [&](uint32_t it)
    {
        const uint64_t key = mix(wid * 0x9e3779b97f4a7c15ULL + (uint64_t)it * G + sub);
        const uint64_t row = (uint64_t)(((__uint128_t)key * n_rows) >> 64);
        const uint8_t *p = tab + row * ROW + col * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&buf[wib][it % STAGES][lane])), "l"(p));
        asm volatile("cp.async.commit_group;");
    };
    for (uint32_t it = 0; it < STAGES - 1 && it < iters; ++it)
        issue(it);
    for (uint32_t it = 0; it < iters; ++it)
    {
        if (it + STAGES - 1 < iters)
            issue(it + STAGES - 1);
        else
            asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1));
        const uint4 v = buf[wib][it % STAGES][lane];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678)
        *sink = acc;
}

// bulk copies: every lane fetches one whole row per stage (32 rows per warp and stage)
template <int ROW, int STAGES>
__global__ void __launch_bounds__(32 * WARPS) gather_bulk(const uint8_t *tab, uint64_t n_rows, uint32_t iters, uint32_t *sink)
{
    __shared__ alignas(128) uint8_t buf[WARPS][STAGES][32 * ROW];
    __shared__ alignas(8) uint64_t bar[WARPS][STAGES];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t wid = (uint64_t)blockIdx.x * WARPS + wib;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[wib][s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    __syncthreads();
    uint32_t acc = 0;
    auto issue = [&](uint32_t it)
    {
        const int s = it % STAGES;
        const uint64_t key = mix(wid * 0x9e3779b97f4a7c15ULL + (uint64_t)it * 32 + lane);
        const uint64_t row = (uint64_t)(((__uint128_t)key * n_rows) >> 64);
        const uint8_t *p = tab + row * ROW;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[wib][s])), "r"(32 * ROW));
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(&buf[wib][s][lane * ROW])),
                     "l"(p), "r"(ROW), "r"(smem_u32(&bar[wib][s]))
                     : "memory");
    };
    for (uint32_t it = 0; it < STAGES && it < iters; ++it)
        issue(it);
    for (uint32_t it = 0; it < iters; ++it)
    {
        const int s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar[wib][s])), "r"(parity) : "memory");
        const uint4 *r = reinterpret_cast<const uint4 *>(&buf[wib][s][lane * ROW]);
#pragma unroll
        for (int q = 0; q < ROW / 16; ++q)
        {
            // rotate the 16-byte chunk order per lane: conflict-free LDS.128
            const uint4 v = r[(q + lane) % (ROW / 16)];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (it + STAGES < iters)
            issue(it + STAGES);
    }
    if (acc == 0x12345678)
        *sink = acc;
}

static cudaEvent_t ev_a, ev_b;
template <typename F>
float time_best(F launch)
{
    launch(true);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        CK(cudaEventRecord(ev_a));
        launch(false);
        CK(cudaEventRecord(ev_b));
        CK(cudaEventSynchronize(ev_b));
        float ms; CK(cudaEventElapsedTime(&ms, ev_a, ev_b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int ROW>
void run_all(const uint8_t *tab, uint64_t bytes, uint32_t *sink, int sms, bool first)
{
    const uint64_t n_rows = bytes / ROW;
    const double target = 8e9;
    auto gbps = [&](uint64_t rows_per_iter, int blocks, auto kernel)
    {
        const uint32_t iters = (uint32_t)(target / ((double)rows_per_iter * ROW));
        const float ms = time_best([&](bool warm) { kernel<<<blocks, 32 * WARPS>>>(tab, n_rows, warm ? iters / 8 + 1 : iters, sink); });
        return (double)rows_per_iter * iters * ROW / (ms * 1e-3) / 1e9;
    };
    constexpr int G = 32 / (ROW / 16);
    const int b8 = sms * 8, b4 = sms * 4;
    const double ldg = gbps((uint64_t)b8 * WARPS * G * 4, b8, gather_ldg<ROW>);
    const double sts4 = gbps((uint64_t)b8 * WARPS * G, b8, gather_ldgsts<ROW, 4>);
    const double sts8 = gbps((uint64_t)b8 * WARPS * G, b8, gather_ldgsts<ROW, 8>);
    const double bulk2 = gbps((uint64_t)b4 * WARPS * 32, b4, gather_bulk<ROW, 2>);
    const double bulk4 = gbps((uint64_t)b4 * WARPS * 32, b4, gather_bulk<ROW, 4>);
    const double bulk4b = gbps((uint64_t)b8 * WARPS * 32, b8, gather_bulk<ROW, 4>);
    printf("%s  {\"row_bytes\": %d, \"ldg\": %.0f, \"ldgsts_4stage\": %.0f, \"ldgsts_8stage\": %.0f, \"bulk_2stage_4cta\": %.0f, \"bulk_4stage_4cta\": %.0f, \"bulk_4stage_8cta\": %.0f}",
           first ? "" : ",\n", ROW, ldg, sts4, sts8, bulk2, bulk4, bulk4b);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 8ull) << 30;
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaEventCreate(&ev_a)); CK(cudaEventCreate(&ev_b));
    uint8_t *tab; uint32_t *sink;
    CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(tab, 1, bytes));
    printf("{\"table_gib\": %llu, \"sms\": %d, \"useful_GBps\": [\n", (unsigned long long)(bytes >> 30), sms);
    run_all<64>(tab, bytes, sink, sms, true);
    run_all<32>(tab, bytes, sink, sms, false);
    printf("\n]}\n");
    return 0;
}
