// gather_bench2.cu -- second microbenchmark for kernel #2's design decisions on B200 (sm_100a), narrow rows only
// (64 B = the T=64 indexes of BASELINE configs[1..2], 128 B, 32 B):
//   (1) ld.global L2 prefetch-size qualifiers (.L2::64B / .L2::128B / .L2::256B) -- does any of them stop the
//       128-byte DRAM fetch per 64-byte row seen in profiles/r1_e_ncu_full.txt?
//   (2) resident warps per SM x loads in flight per lane (occupancy against memory-level parallelism)
//   (3) the XOR-filter probe shape: 3 rows per probe, all random ("rrr") against row 0 taken in ascending slot
//       order ("srr": what sorting/bucketing the hashes of a read group by their segment-0 slot would give)
// Output: one JSON object; GB/s are useful bytes (rows x row bytes), Gprobes/s for the probe shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench2 gather_bench2.cu
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
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 3)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 4)
    {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    }
    else
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// plain random rows: per step a warp gathers (32/lpr)*UNROLL rows
template <int MODE, int UNROLL>
__global__ void gather(const uint8_t *tab, uint64_t n_rows, uint32_t row_bytes, uint32_t iters, uint32_t *sink)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lpr = row_bytes / 16;
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
            v[u] = ld16<MODE>(tab + row * row_bytes + col * 16);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678)
        *sink = acc;
}

// probe shape: 3 rows per probe in three segments of n_rows/3.  SORTED: row 0 of probe j (global numbering over
// the launch, warps own contiguous probe ranges) is floor(j * density) -- ascending slot order.
template <int UNROLL, bool SORTED>
__global__ void probe3(const uint8_t *tab, uint64_t seg_rows, uint32_t row_bytes, uint32_t iters, double density, uint32_t *sink)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lpr = row_bytes / 16;
    const uint32_t sub = lane / lpr, col = lane % lpr, G = 32 / lpr;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t per_warp = (uint64_t)iters * UNROLL * G;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it)
    {
        uint4 a[UNROLL], b[UNROLL], c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const uint64_t j = wid * per_warp + ((uint64_t)it * UNROLL + u) * G + sub;
            const uint64_t key = mix(j * 0x9e3779b97f4a7c15ULL + 1);
            uint64_t r0;
            if (SORTED)
                r0 = (uint64_t)((double)j * density) % seg_rows;
            else
                r0 = (uint64_t)(uint32_t)key * seg_rows >> 32;
            const uint64_t r1 = ((uint64_t)(uint32_t)(key >> 21) * seg_rows >> 32) + seg_rows;
            const uint64_t r2 = ((uint64_t)(uint32_t)(key >> 42 | key << 22) * seg_rows >> 32) + 2 * seg_rows;
            a[u] = ld16<0>(tab + r0 * row_bytes + col * 16);
            b[u] = ld16<0>(tab + r1 * row_bytes + col * 16);
            c[u] = ld16<0>(tab + r2 * row_bytes + col * 16);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            acc += (a[u].x ^ b[u].x ^ c[u].x) + (a[u].y ^ b[u].y ^ c[u].y) + (a[u].z ^ b[u].z ^ c[u].z) + (a[u].w ^ b[u].w ^ c[u].w);
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

template <int MODE, int UNROLL>
double run_gather(const uint8_t *tab, uint64_t bytes, uint32_t row_bytes, uint32_t *sink, int sms, int wps)
{
    const uint64_t n_rows = bytes / row_bytes;
    const uint32_t G = 32 / (row_bytes / 16);
    const int blocks = sms * (wps / 4);
    const uint64_t rows_per_iter = (uint64_t)blocks * 4 * G * UNROLL;
    uint32_t iters = (uint32_t)(12e9 / ((double)rows_per_iter * row_bytes));
    float ms = time_best([&](bool warm) { gather<MODE, UNROLL><<<blocks, 128>>>(tab, n_rows, row_bytes, warm ? iters / 8 + 1 : iters, sink); });
    return (double)rows_per_iter * iters * row_bytes / (ms * 1e-3) / 1e9;
}

template <int UNROLL, bool SORTED>
double run_probe(const uint8_t *tab, uint64_t bytes, uint32_t row_bytes, double density, uint32_t *sink, int sms, int wps)
{
    const uint64_t seg_rows = bytes / row_bytes / 3;
    const uint32_t G = 32 / (row_bytes / 16);
    const int blocks = sms * (wps / 4);
    const uint64_t probes_per_iter = (uint64_t)blocks * 4 * G * UNROLL;
    uint32_t iters = (uint32_t)(12e9 / ((double)probes_per_iter * 3 * row_bytes));
    float ms = time_best([&](bool warm) { probe3<UNROLL, SORTED><<<blocks, 128>>>(tab, seg_rows, row_bytes, warm ? iters / 8 + 1 : iters, density, sink); });
    return (double)probes_per_iter * iters / (ms * 1e-3) / 1e9; // Gprobes/s
}

int main(int argc, char **argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 8ull) << 30;
    const bool quick = argc > 2; // for ncu: only the qualifier sweep at one shape
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaEventCreate(&ev_a)); CK(cudaEventCreate(&ev_b));
    uint8_t *tab; uint32_t *sink;
    CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(tab, 1, bytes));
    printf("{\"table_gib\": %llu, \"sms\": %d,\n", (unsigned long long)(bytes >> 30), sms);
    printf(" \"qualifiers_GBps\": [\n");
    bool first = true;
    for (uint32_t rb : {32u, 64u, 128u})
    {
        if (quick && rb != 64u) continue;
        printf("%s  {\"row_bytes\": %u, \"nc_noalloc\": %.0f, \"L2_64B\": %.0f, \"L2_128B\": %.0f, \"L2_256B\": %.0f, \"evict_first\": %.0f, \"nc_l1alloc\": %.0f}",
               first ? "" : ",\n", rb,
               run_gather<0, 4>(tab, bytes, rb, sink, sms, 32), run_gather<1, 4>(tab, bytes, rb, sink, sms, 32),
               run_gather<2, 4>(tab, bytes, rb, sink, sms, 32), run_gather<3, 4>(tab, bytes, rb, sink, sms, 32),
               run_gather<4, 4>(tab, bytes, rb, sink, sms, 32), run_gather<5, 4>(tab, bytes, rb, sink, sms, 32));
        first = false;
        fflush(stdout);
    }
    printf("\n ]");
    if (!quick)
    {
        printf(",\n \"occupancy_GBps_row64\": [\n");
        first = true;
        for (int wps : {16, 24, 32, 40, 48, 56, 64})
        {
            printf("%s  {\"warps_per_sm\": %d, \"u1\": %.0f, \"u2\": %.0f, \"u3\": %.0f, \"u4\": %.0f, \"u6\": %.0f, \"u8\": %.0f}", first ? "" : ",\n", wps,
                   run_gather<0, 1>(tab, bytes, 64, sink, sms, wps), run_gather<0, 2>(tab, bytes, 64, sink, sms, wps),
                   run_gather<0, 3>(tab, bytes, 64, sink, sms, wps), run_gather<0, 4>(tab, bytes, 64, sink, sms, wps),
                   run_gather<0, 6>(tab, bytes, 64, sink, sms, wps), run_gather<0, 8>(tab, bytes, 64, sink, sms, wps));
            first = false;
            fflush(stdout);
        }
        printf("\n ],\n \"probe3_Gprobes\": [\n");
        first = true;
        for (uint32_t rb : {64u, 128u, 256u})
            for (int wps : {32, 48})
            {
                printf("%s  {\"row_bytes\": %u, \"warps_per_sm\": %d, \"rrr_u1\": %.2f, \"rrr_u2\": %.2f, \"srr_d0.05_u2\": %.2f, \"srr_d0.4_u2\": %.2f, \"srr_d1_u2\": %.2f, \"srr_d4_u2\": %.2f}",
                       first ? "" : ",\n", rb, wps,
                       run_probe<1, false>(tab, bytes, rb, 0, sink, sms, wps), run_probe<2, false>(tab, bytes, rb, 0, sink, sms, wps),
                       run_probe<2, true>(tab, bytes, rb, 0.05, sink, sms, wps), run_probe<2, true>(tab, bytes, rb, 0.4, sink, sms, wps),
                       run_probe<2, true>(tab, bytes, rb, 1.0, sink, sms, wps), run_probe<2, true>(tab, bytes, rb, 4.0, sink, sms, wps));
                first = false;
                fflush(stdout);
            }
        printf("\n ]");
    }
    printf("\n}\n");
    return 0;
}
