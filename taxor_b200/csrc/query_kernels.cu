// query_kernels.cu -- kernel #2 of the `taxor search` hot path for sm_100a.
//
// Replaces, for a whole batch of reads and one level of the HIXF tree at a time,
//     interleaved_xor_filter<uint8_t>::counting_agent<uint32_t>().bulk_count(values)   (call site hixf.hpp:307-309)
//     the bin scan of membership_agent::bulk_contains_impl                              (hixf.hpp:311-339)
// A work item is (read, IXF).  The reference's recursion (hixf.hpp:322) becomes a level-synchronous queue:
// merged bins whose count reaches the read's threshold append (read, child IXF) to the next level's queue,
// split-bin runs that reach it append (read, user bin, count) to the hit list.  Hits of one read are put into
// the reference's DFS pre-order on the host (engine.cu).
//
// HBM-bound random gather: per hash three rows of `tbins` bytes.  A lane owns 16 consecutive bins (one 16-byte
// load per row); tbins/16 lanes cover a row, so a warp probes 32/(tbins/16) hashes per step.  Hits are detected
// with a branch-free zero-byte test on r0^r1^r2^splat(f) and accumulated in byte-packed registers, spilled to
// 32-bit shared-memory counters at most every 255 steps.  Loads of the next step are issued before the current
// step is reduced (register double buffering); the index has no reuse, so rows are fetched with
// ld.global.nc.L1::no_allocate.
#include "device_types.cuh"
#include "ixf_arith.cuh"

#include <cstdlib>

namespace txr
{
namespace
{
__device__ __forceinline__ uint4 ldg_row16(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 0x01 in every byte of x that is zero (exact, no carries across bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
{
    const uint32_t t = ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x; // bit 7 of a byte set <=> byte != 0
    return (~t & 0x80808080u) >> 7;
}

struct Probe
{
    uint4 r0, r1, r2;
    uint32_t fs; // fingerprint splatted over 4 bytes; 0 matches are suppressed through `live`
    bool live;
};

__device__ __forceinline__ void probe_issue(Probe &p, const IxfDev &d, const uint8_t *col_base, uint64_t key, bool live)
{
    p.live = live;
    if (live)
    {
        const uint64_t h = ixf_mix(key, d.seed);
        uint32_t p0, p1, p2;
        ixf_slots(h, d.seg_len, p0, p1, p2);
        p.fs = ixf_fingerprint(h) * 0x01010101u;
        p.r0 = ldg_row16(col_base + (uint64_t)p0 * d.tbins);
        p.r1 = ldg_row16(col_base + (uint64_t)p1 * d.tbins);
        p.r2 = ldg_row16(col_base + (uint64_t)p2 * d.tbins);
    }
}

__device__ __forceinline__ void probe_reduce(const Probe &p, uint32_t (&acc)[4])
{
    if (p.live)
    {
        acc[0] += zero_bytes(p.r0.x ^ p.r1.x ^ p.r2.x ^ p.fs);
        acc[1] += zero_bytes(p.r0.y ^ p.r1.y ^ p.r2.y ^ p.fs);
        acc[2] += zero_bytes(p.r0.z ^ p.r1.z ^ p.r2.z ^ p.fs);
        acc[3] += zero_bytes(p.r0.w ^ p.r1.w ^ p.r2.w ^ p.fs);
    }
}

// add the 16 byte counters of a lane to the 32-bit shared counters of its 16 bins
__device__ __forceinline__ void acc_flush(uint32_t (&acc)[4], uint32_t *cnt16)
{
#pragma unroll
    for (int wd = 0; wd < 4; ++wd)
    {
        const uint32_t a = acc[wd];
        if (a)
        {
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
                const uint32_t c = (a >> (8 * b)) & 0xffu;
                if (c)
                    atomicAdd(&cnt16[4 * wd + b], c);
            }
        }
        acc[wd] = 0;
    }
}

// One warp probes all hashes of a read against a column chunk of an IXF.
//   chunk_off : first byte (bin) of the chunk inside a row, multiple of 16
//   lpr       : lanes per row for this chunk (chunk bytes / 16), 1..32
//   cnt       : shared counters of the chunk's first bin
template <int UNROLL>
__device__ __forceinline__ void probe_chunk(const IxfDev &d, const uint64_t *__restrict__ hp, uint32_t H,
                                            uint32_t chunk_off, uint32_t lpr, uint32_t *cnt, int lane)
{
    const uint32_t G = 32u / lpr;         // hashes per step
    const uint32_t sub = (uint32_t)lane / lpr;
    const uint32_t col = (uint32_t)lane - sub * lpr;
    const bool active = sub < G;
    const uint8_t *col_base = d.fp + chunk_off + 16u * col;
    uint32_t acc[4] = {0, 0, 0, 0};
    uint32_t steps = 0;
    Probe pr[UNROLL];
    for (uint32_t h0 = 0; h0 < H; h0 += G * UNROLL)
    {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const uint32_t idx = h0 + u * G + sub;
            const bool live = active && idx < H;
            const uint64_t key = live ? hp[idx] : 0;
            probe_issue(pr[u], d, col_base, key, live);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            probe_reduce(pr[u], acc);
        steps += UNROLL;
        if (steps + UNROLL > 255)
        {
            if (active)
                acc_flush(acc, cnt + 16u * col);
            steps = 0;
        }
    }
    if (active)
        acc_flush(acc, cnt + 16u * col);
}

// bin scan of hixf.hpp:311-339 for one (read, IXF) from the shared counters; warp- or block-wide
// (`tid`/`nt` = thread index / count of the cooperating group, all of its lanes must call this).
__device__ __forceinline__ void scan_bins(const QueryArgs &a, const IxfDev &d, uint32_t read, uint64_t thr,
                                          const uint32_t *cnt, uint32_t tid, uint32_t nt)
{
    const int lane = threadIdx.x & 31;
    for (uint32_t b0 = 0; b0 < d.bins; b0 += nt)
    {
        const uint32_t b = b0 + tid;
        bool is_hit = false, is_desc = false;
        uint32_t sum = 0;
        int32_t target = 0;
        if (b < d.bins)
        {
            const uint8_t kind = a.bin_kind[d.meta_off + b];
            if (kind == kBinMerged)
            {
                sum = cnt[b];
                is_desc = (uint64_t)sum >= thr;                   // hixf.hpp:321
                target = a.bin_child[d.meta_off + b];
            }
            else if (kind == kBinRunEnd)
            {
                for (uint32_t q = a.bin_run_begin[d.meta_off + b]; q <= b; ++q)
                    sum += cnt[q];                                 // hixf.hpp:315 over the split run
                is_hit = (uint64_t)sum >= thr;                    // hixf.hpp:328
                target = a.bin_ub[d.meta_off + b];
            }
        }
        // warp-aggregated appends
        const uint32_t hit_bal = __ballot_sync(0xffffffffu, is_hit);
        if (hit_bal)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(a.n_hits, (uint32_t)__popc(hit_bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is_hit)
            {
                const uint32_t at = base + __popc(hit_bal & ((1u << lane) - 1u));
                if (at < a.hit_cap)
                {
                    a.hit_read[at] = read;
                    a.hit_ub[at] = target;
                    a.hit_cnt[at] = sum;
                }
            }
        }
        const bool to_small = is_desc && a.ixf[target].tbins <= kSmallRowBytes;
        const bool to_large = is_desc && !to_small;
        const uint32_t sm_bal = __ballot_sync(0xffffffffu, to_small);
        if (sm_bal)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(a.next_small_n, (uint32_t)__popc(sm_bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (to_small)
            {
                const uint32_t at = base + __popc(sm_bal & ((1u << lane) - 1u));
                if (at < a.next_cap)
                    a.next_small[at] = make_uint2(read, (uint32_t)target);
            }
        }
        const uint32_t lg_bal = __ballot_sync(0xffffffffu, to_large);
        if (lg_bal)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(a.next_large_n, (uint32_t)__popc(lg_bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (to_large)
            {
                const uint32_t at = base + __popc(lg_bal & ((1u << lane) - 1u));
                if (at < a.next_cap)
                    a.next_large[at] = make_uint2(read, (uint32_t)target);
            }
        }
    }
}
} // namespace

constexpr int kQueryWarps = 4;
constexpr int kQueryUnroll = 2;

// ---- IXFs with tbins <= 512: one warp per (read, IXF) ----
__global__ void __launch_bounds__(32 * kQueryWarps) ixf_query_small_kernel(QueryArgs a)
{
    __shared__ uint32_t s_cnt[kQueryWarps][kSmallRowBytes];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *cnt = s_cnt[wib];
    for (int i = lane; i < (int)kSmallRowBytes; i += 32)
        cnt[i] = 0;
    __syncwarp();
    const uint32_t n_items = a.items ? *a.n_items_ptr : a.n_items_direct;
    unsigned long long bytes = 0, items = 0;
    while (true)
    {
        uint32_t it = 0;
        if (lane == 0)
            it = atomicAdd(a.cursor, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= min(n_items, a.items_cap))
            break;
        uint32_t read = it, x = 0;
        if (a.items)
        {
            const uint2 w = a.items[it];
            read = w.x;
            x = w.y;
        }
        const IxfDev d = a.ixf[x];
        const uint32_t H = a.hash_count[read];
        const uint64_t *hp = a.hashes + a.hash_off[read];
        const uint64_t thr = H < a.lut_len ? a.thr_lut[H] : ~0ULL;
        probe_chunk<kQueryUnroll>(d, hp, H, 0u, d.tbins >> 4, cnt, lane);
        __syncwarp();
        scan_bins(a, d, read, thr, cnt, (uint32_t)lane, 32u);
        __syncwarp();
        for (uint32_t i = lane; i < d.tbins; i += 32)
            cnt[i] = 0;
        __syncwarp();
        bytes += (unsigned long long)H * 3ull * d.tbins + 8ull * H;
        ++items;
    }
    if (lane == 0 && items)
    {
        atomicAdd(a.stat_bytes, bytes);
        atomicAdd(a.stat_items, items);
    }
}

// ---- IXFs with tbins > 512: one CTA per (read, IXF); warps take 512-byte column chunks of the rows ----
__global__ void __launch_bounds__(256) ixf_query_large_kernel(QueryArgs a)
{
    extern __shared__ uint32_t s_cnt_dyn[]; // tbins counters
    __shared__ uint32_t s_item;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t n_items = a.items ? *a.n_items_ptr : a.n_items_direct;
    unsigned long long bytes = 0, items = 0;
    while (true)
    {
        if (threadIdx.x == 0)
            s_item = atomicAdd(a.cursor, 1u);
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= min(n_items, a.items_cap))
            break;
        uint32_t read = it, x = 0;
        if (a.items)
        {
            const uint2 w = a.items[it];
            read = w.x;
            x = w.y;
        }
        const IxfDev d = a.ixf[x];
        for (uint32_t i = threadIdx.x; i < d.tbins; i += blockDim.x)
            s_cnt_dyn[i] = 0;
        __syncthreads();
        const uint32_t H = a.hash_count[read];
        const uint64_t *hp = a.hashes + a.hash_off[read];
        const uint64_t thr = H < a.lut_len ? a.thr_lut[H] : ~0ULL;
        const uint32_t n_chunks = (d.tbins + kSmallRowBytes - 1) / kSmallRowBytes;
        for (uint32_t c = wib; c < n_chunks; c += nwarps)
        {
            const uint32_t off = c * kSmallRowBytes;
            const uint32_t width = min(kSmallRowBytes, d.tbins - off);
            probe_chunk<kQueryUnroll>(d, hp, H, off, width >> 4, s_cnt_dyn + off, lane);
        }
        __syncthreads();
        scan_bins(a, d, read, thr, s_cnt_dyn, threadIdx.x, blockDim.x);
        __syncthreads();
        bytes += (unsigned long long)H * 3ull * d.tbins + 8ull * H;
        ++items;
    }
    if (threadIdx.x == 0 && items)
    {
        atomicAdd(a.stat_bytes, bytes);
        atomicAdd(a.stat_items, items);
    }
}

// bulk_count of a single IXF for a single value list (parity entry point txr_ixf_bulk_count)
__global__ void __launch_bounds__(256) ixf_bulk_count_kernel(IxfDev d, const uint64_t *values, uint32_t n, uint32_t *counts)
{
    extern __shared__ uint32_t s_cnt_dyn[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t i = threadIdx.x; i < d.tbins; i += blockDim.x)
        s_cnt_dyn[i] = 0;
    __syncthreads();
    const uint32_t n_chunks = (d.tbins + kSmallRowBytes - 1) / kSmallRowBytes;
    for (uint32_t c = wib; c < n_chunks; c += nwarps)
    {
        const uint32_t off = c * kSmallRowBytes;
        const uint32_t width = min(kSmallRowBytes, d.tbins - off);
        probe_chunk<kQueryUnroll>(d, values, n, off, width >> 4, s_cnt_dyn + off, lane);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < d.bins; i += blockDim.x)
        counts[i] = s_cnt_dyn[i];
}

// ---- launchers ----
// CTAs per SM of the persistent query grids (tuning knob, TXR_QUERY_CTAS_PER_SM; 8 = all 32 warps an SM can hold at
// 61 registers, fewer leaves room for the compute-bound hash/dedup kernels of the neighbouring pipeline slots)
static int query_ctas_per_sm()
{
    static int v = 0;
    if (!v)
    {
        const char *e = getenv("TXR_QUERY_CTAS_PER_SM");
        v = e ? atoi(e) : 8;
        if (v < 1 || v > 16)
            v = 8;
    }
    return v;
}

cudaError_t launch_query_small(const QueryArgs &a, int sm_count, cudaStream_t st)
{
    ixf_query_small_kernel<<<sm_count * query_ctas_per_sm(), 32 * kQueryWarps, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_query_large(const QueryArgs &a, int sm_count, uint32_t max_tbins, cudaStream_t st)
{
    const size_t smem = (size_t)max_tbins * 4;
    cudaFuncSetAttribute(ixf_query_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ixf_query_large_kernel<<<sm_count * 4, 256, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bulk_count(const IxfDev &d, const uint64_t *values, uint32_t n, uint32_t *counts, cudaStream_t st)
{
    const size_t smem = (size_t)d.tbins * 4;
    cudaFuncSetAttribute(ixf_bulk_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ixf_bulk_count_kernel<<<1, 256, smem, st>>>(d, values, n, counts);
    return cudaGetLastError();
}
} // namespace txr
