// gather_bench.cu -- measures the B200 random-row-gather ceiling that bounds kernel #2: rows of R bytes
// (R = 64..4096, aligned) at uniformly random positions of a table much larger than L2, 16 bytes per lane,
// several independent rows in flight per lane.  Reports useful GB/s per row size and per
// cudaLimitMaxL2FetchGranularity setting.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
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

template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint8_t *p)
{
    uint4 r;
    if (MODE == 0)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 1)
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 2)
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else
    {
        // sm_100 256-bit load (32 bytes per lane) with an L2 evict-first hint
        uint4 q;
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w), "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
        r.x ^= q.x; r.y ^= q.y; r.z ^= q.z; r.w ^= q.w;
    }
    return r;
}

// each warp gathers `iters` steps; per step (32/lpr)*UNROLL rows of lpr*16 bytes
template <int MODE, int UNROLL>
__global__ void gather(const uint8_t *tab, uint64_t n_rows, uint32_t row_bytes, uint32_t iters, uint32_t *sink)
{
    const int lane = threadIdx.x & 31;
    constexpr uint32_t LB = MODE == 3 ? 32 : 16; // bytes per lane per load
    const uint32_t lpr = row_bytes >= 32 * LB ? 32 : row_bytes / LB;
    const uint32_t chunks = row_bytes >= 32 * LB ? row_bytes / (32 * LB) : 1;
    const uint32_t sub = lane / lpr, col = lane % lpr, G = 32 / lpr;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it)
    {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const uint64_t key = mix(wid * 0x9e3779b97f4a7c15ULL + ((uint64_t)it * UNROLL + u) * G + sub);
            const uint64_t row = (uint64_t)(((__uint128_t)key * n_rows) >> 64);
            const uint8_t *p = tab + row * row_bytes + col * LB;
            v[u] = ld16<MODE>(p);
            for (uint32_t c = 1; c < chunks; ++c)
            {
                uint4 w = ld16<MODE>(p + c * 32 * LB);
                v[u].x ^= w.x; v[u].y ^= w.y;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678)
        *sink = acc;
}

template <int MODE, int UNROLL>
double run(const uint8_t *tab, uint64_t bytes, uint32_t row_bytes, uint32_t *sink, int sms, int wps)
{
    const uint64_t n_rows = bytes / row_bytes;
    constexpr uint32_t LB = MODE == 3 ? 32 : 16;
    const uint32_t lpr = row_bytes >= 32 * LB ? 32 : row_bytes / LB;
    const uint32_t G = 32 / lpr;
    const int blocks = sms * (wps / 4), threads = 128;
    const double target = 24e9; // useful bytes per launch
    const uint64_t rows_per_iter = (uint64_t)blocks * 4 * G * UNROLL;
    uint32_t iters = (uint32_t)(target / ((double)rows_per_iter * row_bytes));
    if (iters < 4) iters = 4;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gather<MODE, UNROLL><<<blocks, threads>>>(tab, n_rows, row_bytes, iters / 4 + 1, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        CK(cudaEventRecord(a));
        gather<MODE, UNROLL><<<blocks, threads>>>(tab, n_rows, row_bytes, iters, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return (double)rows_per_iter * iters * row_bytes / (best * 1e-3) / 1e9;
}

int main(int argc, char **argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 8ull) << 30;
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint8_t *tab; uint32_t *sink;
    CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(tab, 1, bytes));
    printf("{\"table_gib\": %llu, \"sms\": %d, \"results\": [\n", (unsigned long long)(bytes >> 30), sms);
    const size_t grans[] = {0, 32, 64, 128};
    bool first = true;
    for (size_t g : grans)
    {
        if (g) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g));
        size_t cur = 0; CK(cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity));
        for (uint32_t rb : {64u, 128u, 256u, 512u, 1024u, 4096u})
        {
            double r0 = run<0, 2>(tab, bytes, rb, sink, sms, 32);
            double r1 = run<0, 4>(tab, bytes, rb, sink, sms, 32);
            double r2 = run<1, 4>(tab, bytes, rb, sink, sms, 32);
            double r3 = run<3, 4>(tab, bytes, rb, sink, sms, 32);
            double r4 = run<0, 6>(tab, bytes, rb, sink, sms, 48);
            printf("%s {\"l2_fetch_granularity\": %zu, \"row_bytes\": %u, \"nc_u2_w32\": %.0f, \"nc_u4_w32\": %.0f, \"cg_u4_w32\": %.0f, \"nc_v8_evict_first_u4_w32\": %.0f, \"nc_u6_w48\": %.0f}",
                   first ? "" : ",\n", cur, rb, r0, r1, r2, r3, r4);
            first = false;
            fflush(stdout);
        }
    }
    printf("\n]}\n");
    return 0;
}
