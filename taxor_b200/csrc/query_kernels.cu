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
// HBM-bound random gather: per hash three rows of `tbins` bytes.
//   rows <= 512 bytes (ixf_query_small_kernel): one warp per item; a lane owns 16 consecutive bins (one 16-byte load per
//     row); tbins/16 lanes cover a row, so a warp probes 32/(tbins/16) hashes per step.  UNROLL steps are in flight per warp;
//     the keys of a step are loaded one step ahead (the hash lists of a batch are GBs, i.e. DRAM).  Builds: two steps in
//     flight at 64 / 70 registers (8 / 7 CTAs per SM) for a kernel that owns the GPU, one step at 40 / 32 registers for
//     the probes that share the SMs with the hash stage of the next batch (engine.cu: overlap).
//   rows > 512 bytes (ixf_query_large_kernel): one CTA per item, the hash list dealt to the 8 warps, every warp covers whole
//     rows in 2 KB passes (the wide roots of GTDB-scale indexes).
// Hits are detected with a branch-free zero-byte test on r0^r1^r2^splat(f) and accumulated in byte-packed registers,
// spilled to 32-bit shared-memory counters at most every 255 steps.  The index has no reuse at the root, so rows are fetched
// with ld.global.nc.L1::no_allocate (64-byte rows: with the .L2::64B prefetch-size hint, two DRAM sectors instead of a
// line).  The probe arithmetic is either the prototype's with folded constants or driven by the descriptor the index was
// uploaded with (probe_address<GEN>, ixf_arith.cuh).
#include "device_types.cuh"
#include "ixf_arith.cuh"

#include <algorithm>
#include <cstdlib>

namespace txr
{
constexpr int kQueryWarps = 4;
constexpr int kQueryUnroll = 2;

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

// the same load with the 64-byte L2 prefetch-size hint: a miss then brings two sectors from DRAM instead of the whole 128-byte
// line (ncu: half the dram__bytes for 64-byte rows, profiles/r1_gather_bench2_ncu.txt).  Rows of T = 64 indexes only.
__device__ __forceinline__ uint4 ldg_row16_s64(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
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

// the same load with an L2 eviction-priority policy (createpolicy): rows worth keeping against rows that stream
__device__ __forceinline__ uint4 ldg_row16_hint(const uint8_t *p, uint64_t policy)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// L2 residency plan of an IXF.  What survives in L2 between two touches is far less than the nominal 126 MB
// (profiles/r1_launches_root_partitioned_10GB.txt: a 33 MB reuse distance already misses half the time), so only
// IXFs whose FIRST segment is small are worth it: its rows are loaded evict_last and the other two segments'
// rows evict_first, which keeps one of the three lines of a probe in L2 while the level queue (grouped by IXF)
// works through that IXF.
struct L2Plan
{
    bool split;        // segment 0 evict_last, segments 1 and 2 evict_first
    uint64_t keep, stream;
    bool sector64;     // 64-byte rows: ask L2 for 64 bytes per miss, not for the 128-byte line
};
constexpr uint64_t kL2KeepBytes = 28ull << 20;
__device__ __forceinline__ L2Plan l2_plan_for(const IxfDev &d, bool enabled, bool sector64 = false)
{
    L2Plan p;
    p.sector64 = sector64 && d.tbins == 64;
    p.split = enabled && (uint64_t)d.seg_len * d.tbins <= kL2KeepBytes && (uint64_t)d.seg_len * d.tbins * 3 > kL2KeepBytes;
    p.keep = l2_policy_evict_last();
    p.stream = l2_policy_evict_first();
    return p;
}

// the three rows + fingerprint of a key.  GEN == false: the prototype's arithmetic with folded constants (the default
// scheme); GEN == true: driven by the descriptor the index was uploaded with (ixf_arith.cuh)
template <bool GEN>
__device__ __forceinline__ uint32_t probe_address(const IxfDev &d, const IxfScheme &sch, uint64_t key, uint32_t &p0, uint32_t &p1,
                                                  uint32_t &p2)
{
    if constexpr (GEN)
    {
        const uint64_t h = ixf_mix_g(key, d.seed, sch);
        ixf_slots_g(h, d.seg_len, d.count_len, sch, p0, p1, p2);
        return ixf_fingerprint_g(h, sch) * 0x01010101u;
    }
    else
    {
        const uint64_t h = ixf_mix(key, d.seed);
        ixf_slots(h, d.seg_len, p0, p1, p2);
        return ixf_fingerprint(h) * 0x01010101u;
    }
}

template <bool GEN>
__device__ __forceinline__ void probe_issue(Probe &p, const IxfDev &d, const IxfScheme &sch, const uint8_t *col_base, uint64_t key,
                                            bool live, const L2Plan &l2 = L2Plan{false, 0, 0, false})
{
    p.live = live;
    if (live)
    {
        uint32_t p0, p1, p2;
        p.fs = probe_address<GEN>(d, sch, key, p0, p1, p2);
        if (l2.split)
        {
            p.r0 = ldg_row16_hint(col_base + (uint64_t)p0 * d.tbins, l2.keep);
            p.r1 = ldg_row16_hint(col_base + (uint64_t)p1 * d.tbins, l2.stream);
            p.r2 = ldg_row16_hint(col_base + (uint64_t)p2 * d.tbins, l2.stream);
        }
        else if (l2.sector64)
        {
            p.r0 = ldg_row16_s64(col_base + (uint64_t)p0 * d.tbins);
            p.r1 = ldg_row16_s64(col_base + (uint64_t)p1 * d.tbins);
            p.r2 = ldg_row16_s64(col_base + (uint64_t)p2 * d.tbins);
        }
        else
        {
            p.r0 = ldg_row16(col_base + (uint64_t)p0 * d.tbins);
            p.r1 = ldg_row16(col_base + (uint64_t)p1 * d.tbins);
            p.r2 = ldg_row16(col_base + (uint64_t)p2 * d.tbins);
        }
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
// Early exit (exact).  A hit or a descent needs a run of technical bins (one user bin, or one merged bin) whose
// counts sum to >= thr.  After h of H hashes a run of at most max_run bins holds at most
// max_run * (largest bin count so far + (H - h)), because every remaining hash adds at most 1 to each of its bins.
// Once that is below thr nothing of this item can be reported or descended into, so the remaining probes cannot
// change the output and are skipped (the reference computes them, hixf.hpp:307-309, and throws the counts away).
// Typical: a read of an organism that is not in the index stops after (1 - ratio) of its hashes.  The bound uses
// the lane-local byte counters (times the G hash slots that share a bin) plus the flushed shared counters.
__device__ __forceinline__ uint32_t max_byte4(const uint32_t (&acc)[4])
{
    const uint32_t t = __vmaxu4(__vmaxu4(acc[0], acc[1]), __vmaxu4(acc[2], acc[3]));
    return max(max(t & 0xffu, (t >> 8) & 0xffu), max((t >> 16) & 0xffu, t >> 24));
}

// One warp probes all hashes of a read against a column chunk of an IXF.
//   chunk_off : first byte (bin) of the chunk inside a row, multiple of 16
//   lpr       : lanes per row for this chunk (chunk bytes / 16), 1..32
//   cnt       : shared counters of the chunk's first bin
//   exit_thr  : != 0: the chunk is the whole row and the item may stop early against this threshold
// Returns the number of hashes probed.
template <int UNROLL, bool GEN>
__device__ __forceinline__ uint32_t probe_chunk(const IxfDev &d, const IxfScheme &sch, const uint64_t *__restrict__ hp, uint32_t H,
                                                uint32_t chunk_off, uint32_t lpr, uint32_t *cnt, int lane,
                                                const L2Plan &l2 = L2Plan{false, 0, 0, false}, uint64_t exit_thr = 0)
{
    const uint32_t G = 32u / lpr;         // hashes per step
    const uint32_t sub = (uint32_t)lane / lpr;
    const uint32_t col = (uint32_t)lane - sub * lpr;
    const bool active = sub < G;
    const uint8_t *col_base = d.fp + chunk_off + 16u * col;
    uint32_t acc[4] = {0, 0, 0, 0};
    uint32_t steps = 0;
    // first hash index at which even all-zero counts could rule the item out: max_run * (H - h) < thr
    const uint64_t slack = exit_thr ? (exit_thr - 1) / d.max_run : 0;
    if (exit_thr && slack >= H)
        return 0; // max_run * H < thr: hopeless before the first probe (e.g. the k-mer model's wrapped thresholds)
    uint32_t next_check = exit_thr ? H - (uint32_t)slack : 0xffffffffu;
    bool flushed = false;
    uint32_t probed = H;
    Probe pr[UNROLL];
    // the keys of a step are fetched one step ahead: a batch's hash lists are GBs (DRAM, not L2), and a key load in front of
    // every step's row loads would put two dependent DRAM latencies into each step of a warp
    uint64_t keys[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
    {
        const uint32_t idx = u * G + sub;
        keys[u] = active && idx < H ? hp[idx] : 0;
    }
    for (uint32_t h0 = 0; h0 < H; h0 += G * UNROLL)
    {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const uint32_t idx = h0 + u * G + sub;
            const bool live = active && idx < H;
            probe_issue<GEN>(pr[u], d, sch, col_base, keys[u], live, l2);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const uint32_t idx = h0 + G * UNROLL + u * G + sub;
            keys[u] = active && idx < H ? hp[idx] : 0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            probe_reduce(pr[u], acc);
        steps += UNROLL;
        const uint32_t done = h0 + G * UNROLL;
        if (done >= next_check && done < H)
        {
            uint32_t m = __reduce_max_sync(0xffffffffu, max_byte4(acc)) * G;
            if (flushed)
            {
                uint32_t ms = 0;
                for (uint32_t i = lane; i < lpr * 16u; i += 32)
                    ms = max(ms, cnt[i]);
                m += __reduce_max_sync(0xffffffffu, ms);
            }
            if ((uint64_t)d.max_run * ((uint64_t)m + (H - done)) < exit_thr)
            {
                probed = done;
                break;
            }
            next_check = done + 4 * G * UNROLL;
        }
        if (steps + UNROLL > 255)
        {
            if (active)
                acc_flush(acc, cnt + 16u * col);
            __syncwarp();
            flushed = true;
            steps = 0;
        }
    }
    if (active)
        acc_flush(acc, cnt + 16u * col);
    return probed;
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

// ---- IXFs with tbins <= 512: one warp per (read, IXF) ----
// MINB: CTAs per SM the register allocation must allow (8 = 64 registers, all 32 warps an SM's registers can hold at this
// CTA size; 1 = whatever the code wants, 72 at UNROLL 2 -> 7 CTAs)
template <bool GEN, int UNROLL, int MINB>
__global__ void __launch_bounds__(32 * kQueryWarps, MINB) ixf_query_small_kernel(QueryArgs a)
{
    if (!sm_filter_keep(a.smf))
        return;
    __shared__ uint32_t s_cnt[kQueryWarps][kSmallRowBytes];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *cnt = s_cnt[wib];
    for (int i = lane; i < (int)kSmallRowBytes; i += 32)
        cnt[i] = 0;
    __syncwarp();
    const uint32_t n_items = a.items ? *a.n_items_ptr : a.n_items_direct;
    unsigned long long bytes = 0, items = 0, skipped = 0;
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
        const uint64_t thr = a.thr_read ? a.thr_read[read] : (H < a.lut_len ? a.thr_lut[H] : ~0ULL);
        const uint32_t Hp = probe_chunk<UNROLL, GEN>(d, a.scheme, hp, H, 0u, d.tbins >> 4, cnt, lane,
                                                     l2_plan_for(d, a.l2_hints != 0, a.l2_sector64 != 0), a.early_exit ? thr : 0);
        __syncwarp();
        scan_bins(a, d, read, thr, cnt, (uint32_t)lane, 32u);
        __syncwarp();
        for (uint32_t i = lane; i < d.tbins; i += 32)
            cnt[i] = 0;
        __syncwarp();
        bytes += (unsigned long long)Hp * 3ull * d.tbins + 8ull * Hp;
        skipped += H - Hp;
        ++items;
    }
    if (lane == 0 && items)
    {
        atomicAdd(a.stat_bytes, bytes);
        atomicAdd(a.stat_items, items);
        if (skipped)
            atomicAdd(a.stat_skipped, skipped);
    }
}

// ---- partitioned root: the first level of a batch, probes grouped by their segment-0 slot ----
// Every read probes the root IXF, which is GBs for a 1,000-genome index: three random rows per hash, each a whole
// 128-byte DRAM line on B200 whatever the row size (profiles/r1_gather_bench3_ncu.txt) -- the kernel above already
// runs at that random-line ceiling.  The slot in segment 0 is a monotone function of the low hash word, so grouping
// ALL hashes of a batch by its top bits makes the segment-0 rows of one group a contiguous block of a few MB that
// stays in L2 while the group is processed: one of the three DRAM lines per probe disappears (microbenchmark:
// 12.4 -> 17.3 G probes/s, profiles/r1_gather_bench2.json).  The price is one streaming pass over the hashes (8 B
// read, 12 B written, 12 B read per hash against 384 B of line traffic per probe) and per-(read, bin) counters in
// global memory (16 bit, L2 resident, touched only on a match) instead of shared memory.
//   hist -> scatter (counting sort by partition, CTA-aggregated) -> probe (partition-major chunks) -> scan (per read)
constexpr uint32_t kPartReadsPerUnit = 16;  // reads a CTA groups at a time (about 15k hashes)
constexpr uint32_t kPartMaxLog2 = 10;
constexpr uint32_t kProbeChunk = 256;       // elements per work unit of the probe kernel: ~1.2 M elements in flight on 148 SMs

__device__ __forceinline__ uint32_t root_partition_of(const IxfDev &d, uint64_t key, uint32_t log2_parts)
{
    // ixf_slots: p0 = ((uint32)h * seg_len) >> 32 is monotone in (uint32)h, so its top bits select a slot range
    return log2_parts ? (uint32_t)ixf_mix(key, d.seed) >> (32u - log2_parts) : 0u;
}

__global__ void __launch_bounds__(256) root_part_hist_kernel(RootPartArgs a)
{
    __shared__ uint32_t s_hist[1u << kPartMaxLog2];
    __shared__ uint32_t s_unit;
    const uint32_t parts = 1u << a.log2_parts;
    for (uint32_t i = threadIdx.x; i < parts; i += blockDim.x)
        s_hist[i] = 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t n_units = (a.n_reads + kPartReadsPerUnit - 1) / kPartReadsPerUnit;
    while (true)
    {
        __syncthreads();
        if (threadIdx.x == 0)
            s_unit = atomicAdd(&a.work[0], 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units)
            break;
        for (uint32_t r = u * kPartReadsPerUnit + wib; r < min(a.n_reads, (u + 1) * kPartReadsPerUnit); r += nw)
        {
            const uint32_t H = a.hash_count[r];
            const uint64_t *hp = a.hashes + a.hash_off[r];
            for (uint32_t i = lane; i < H; i += 32)
                atomicAdd(&s_hist[root_partition_of(a.root, hp[i], a.log2_parts)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < parts; i += blockDim.x)
        if (s_hist[i])
            atomicAdd(&a.hist[i], s_hist[i]);
}

// one CTA: cursor = exclusive scan of hist, hist[parts] = total
__global__ void __launch_bounds__(1024) root_part_scan_kernel(RootPartArgs a)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t parts = 1u << a.log2_parts; // <= 1024
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t v = threadIdx.x < parts ? a.hist[threadIdx.x] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += up;
    }
    if (lane == 31)
        s_warp[wib] = incl;
    __syncthreads();
    if (wib == 0)
    {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t up = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d)
                w += up;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t before = wib ? s_warp[wib - 1] : 0;
    if (threadIdx.x < parts)
        a.cursor[threadIdx.x] = before + incl - v;
    if (threadIdx.x == 1023)
        a.hist[parts] = before + incl;
}

__global__ void __launch_bounds__(256) root_part_scatter_kernel(RootPartArgs a)
{
    __shared__ uint32_t s_cnt[1u << kPartMaxLog2];  // elements of this unit per partition, then the local cursor
    __shared__ uint32_t s_base[1u << kPartMaxLog2]; // reserved global position per partition
    __shared__ uint32_t s_unit;
    const uint32_t parts = 1u << a.log2_parts;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t n_units = (a.n_reads + kPartReadsPerUnit - 1) / kPartReadsPerUnit;
    while (true)
    {
        __syncthreads();
        if (threadIdx.x == 0)
            s_unit = atomicAdd(&a.work[1], 1u);
        for (uint32_t i = threadIdx.x; i < parts; i += blockDim.x)
            s_cnt[i] = 0;
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units)
            break;
        const uint32_t r_end = min(a.n_reads, (u + 1) * kPartReadsPerUnit);
        for (uint32_t r = u * kPartReadsPerUnit + wib; r < r_end; r += nw)
        {
            const uint32_t H = a.hash_count[r];
            const uint64_t *hp = a.hashes + a.hash_off[r];
            for (uint32_t i = lane; i < H; i += 32)
                atomicAdd(&s_cnt[root_partition_of(a.root, hp[i], a.log2_parts)], 1u);
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < parts; i += blockDim.x)
        {
            const uint32_t c = s_cnt[i];
            s_base[i] = c ? atomicAdd(&a.cursor[i], c) : 0;
            s_cnt[i] = 0;
        }
        __syncthreads();
        for (uint32_t r = u * kPartReadsPerUnit + wib; r < r_end; r += nw)
        {
            const uint32_t H = a.hash_count[r];
            const uint64_t *hp = a.hashes + a.hash_off[r];
            for (uint32_t i = lane; i < H; i += 32)
            {
                const uint64_t key = hp[i];
                const uint32_t b = root_partition_of(a.root, key, a.log2_parts);
                const uint32_t at = s_base[b] + atomicAdd(&s_cnt[b], 1u);
                a.part_hash[at] = key;
                a.part_read[at] = r;
            }
        }
    }
}

// probes the grouped elements; a match adds 1 to the 16-bit counter of (read, bin)
__global__ void __launch_bounds__(32 * kQueryWarps) root_part_probe_kernel(RootPartArgs a)
{
    const int lane = threadIdx.x & 31;
    const IxfDev d = a.root;
    const uint32_t total = a.hist[1u << a.log2_parts];
    const uint32_t lpr = d.tbins >> 4, G = 32u / lpr;
    const uint32_t sub = (uint32_t)lane / lpr, col = (uint32_t)lane - sub * lpr;
    const bool active = sub < G;
    const uint8_t *col_base = d.fp + 16u * col;
    while (true)
    {
        uint32_t c0 = 0;
        if (lane == 0)
            c0 = atomicAdd(&a.work[2], 1u);
        c0 = __shfl_sync(0xffffffffu, c0, 0);
        if ((uint64_t)c0 * kProbeChunk >= total)
            break;
        const uint32_t e0 = c0 * kProbeChunk, e1 = min(total, e0 + kProbeChunk);
        Probe pr[kQueryUnroll];
        uint32_t rd[kQueryUnroll];
        for (uint32_t e = e0; e < e1; e += G * kQueryUnroll)
        {
#pragma unroll
            for (int u = 0; u < kQueryUnroll; ++u)
            {
                const uint32_t idx = e + u * G + sub;
                const bool live = active && idx < e1;
                const uint64_t key = live ? a.part_hash[idx] : 0;
                rd[u] = live ? a.part_read[idx] : 0;
                probe_issue<false>(pr[u], d, IxfScheme{}, col_base, key, live);
            }
#pragma unroll
            for (int u = 0; u < kQueryUnroll; ++u)
            {
                if (!pr[u].live)
                    continue;
                const uint32_t m[4] = {zero_bytes(pr[u].r0.x ^ pr[u].r1.x ^ pr[u].r2.x ^ pr[u].fs),
                                       zero_bytes(pr[u].r0.y ^ pr[u].r1.y ^ pr[u].r2.y ^ pr[u].fs),
                                       zero_bytes(pr[u].r0.z ^ pr[u].r1.z ^ pr[u].r2.z ^ pr[u].fs),
                                       zero_bytes(pr[u].r0.w ^ pr[u].r1.w ^ pr[u].r2.w ^ pr[u].fs)};
                if (m[0] | m[1] | m[2] | m[3])
                {
                    // bins 16*col + 4*wd + b; counters of bins 2j, 2j+1 share one 32-bit word (low, high half)
                    uint32_t *cw = a.counts + ((size_t)rd[u] * d.tbins + 16u * col) / 2;
#pragma unroll
                    for (int wd = 0; wd < 4; ++wd)
                    {
                        if (!m[wd])
                            continue;
                        const uint32_t lo = (m[wd] & 1u) | ((m[wd] >> 8 & 1u) << 16);        // bins 4wd, 4wd+1
                        const uint32_t hi = (m[wd] >> 16 & 1u) | ((m[wd] >> 24 & 1u) << 16); // bins 4wd+2, 4wd+3
                        if (lo)
                            atomicAdd(&cw[2 * wd], lo);
                        if (hi)
                            atomicAdd(&cw[2 * wd + 1], hi);
                    }
                }
            }
        }
    }
}

// per read: counters -> bin scan (thresholds, split runs, descents, hits)
__global__ void __launch_bounds__(32 * kQueryWarps) root_part_scan_kernel2(QueryArgs a, RootPartArgs p)
{
    __shared__ uint32_t s_cnt[kQueryWarps][kSmallRowBytes];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *cnt = s_cnt[wib];
    const IxfDev d = p.root;
    unsigned long long bytes = 0, items = 0;
    while (true)
    {
        uint32_t read = 0;
        if (lane == 0)
            read = atomicAdd(&p.work[3], 1u);
        read = __shfl_sync(0xffffffffu, read, 0);
        if (read >= p.n_reads)
            break;
        const uint32_t *cw = p.counts + (size_t)read * d.tbins / 2;
        for (uint32_t i = lane; i < d.tbins / 2; i += 32)
        {
            const uint32_t w = cw[i];
            cnt[2 * i] = w & 0xffffu;
            cnt[2 * i + 1] = w >> 16;
        }
        __syncwarp();
        const uint32_t H = a.hash_count[read];
        const uint64_t thr = a.thr_read ? a.thr_read[read] : (H < a.lut_len ? a.thr_lut[H] : ~0ULL);
        scan_bins(a, d, read, thr, cnt, (uint32_t)lane, 32u);
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

// ---- IXFs with tbins > 512: one CTA per (read, IXF) ----
// Real Taxor indexes with >= 10k user bins have wide upper levels (root t_max up to 4096, taxor_build.cpp:173-187): a
// row is 1-4 KB, so one probe is three contiguous multi-line reads and the kernel is a streaming gather.  Work split:
// the HASH LIST of the read is dealt to the 8 warps in blocks (kWideHashesPerWarp consecutive hashes per warp and
// block), every warp covers whole rows -- a lane owns 16 bytes of every 512-byte column chunk, CP chunks (CP*3 16-byte
// loads) in flight per hash, rows wider than CP*512 bytes in several passes of CP*512 contiguous bytes.  Counters are
// byte-packed per lane and flushed to the CTA's 32-bit shared counters once per block (at most 8 per byte).  Between
// blocks the CTA can take the exact early exit of the small kernel (same bound, max over the shared counters).
constexpr int kWideWarps = 8;
constexpr uint32_t kWideHashesPerWarp = 8;

template <int CP, int U, bool GEN>
__device__ __forceinline__ void wide_block(const IxfDev &d, const IxfScheme &sch, const uint64_t *__restrict__ hp, uint32_t h_begin,
                                           uint32_t h_end, uint32_t col0, uint32_t *s_cnt, int lane)
{
    uint32_t acc[CP][4];
    bool col_ok[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c)
    {
        acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0;
        col_ok[c] = col0 + 512u * c + 16u * lane < d.tbins;
    }
    const uint8_t *base = d.fp + col0 + 16u * lane;
    for (uint32_t h = h_begin; h < h_end; h += U)
    {
        uint4 r[U][CP][3];
        uint32_t fs[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const bool live = h + u < h_end;
            uint32_t p0, p1, p2;
            fs[u] = probe_address<GEN>(d, sch, live ? hp[h + u] : 0, p0, p1, p2);
            const uint8_t *q0 = base + (uint64_t)p0 * d.tbins, *q1 = base + (uint64_t)p1 * d.tbins, *q2 = base + (uint64_t)p2 * d.tbins;
#pragma unroll
            for (int c = 0; c < CP; ++c)
            {
                if (live && col_ok[c])
                {
                    r[u][c][0] = ldg_row16(q0 + 512 * c);
                    r[u][c][1] = ldg_row16(q1 + 512 * c);
                    r[u][c][2] = ldg_row16(q2 + 512 * c);
                }
                else // r0 ^ r1 ^ r2 ^ fs != 0 in every byte unless fs == 0: make the row itself differ from fs
                {
                    r[u][c][0] = make_uint4(~fs[u], ~fs[u], ~fs[u], ~fs[u]);
                    r[u][c][1] = r[u][c][2] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CP; ++c)
            {
                acc[c][0] += zero_bytes(r[u][c][0].x ^ r[u][c][1].x ^ r[u][c][2].x ^ fs[u]);
                acc[c][1] += zero_bytes(r[u][c][0].y ^ r[u][c][1].y ^ r[u][c][2].y ^ fs[u]);
                acc[c][2] += zero_bytes(r[u][c][0].z ^ r[u][c][1].z ^ r[u][c][2].z ^ fs[u]);
                acc[c][3] += zero_bytes(r[u][c][0].w ^ r[u][c][1].w ^ r[u][c][2].w ^ fs[u]);
            }
    }
#pragma unroll
    for (int c = 0; c < CP; ++c)
        if (col_ok[c])
            acc_flush(acc[c], s_cnt + col0 + 512u * c + 16u * lane);
}

template <int CP, int U, bool GEN>
__device__ __forceinline__ uint32_t wide_item(const IxfDev &d, const IxfScheme &sch, const uint64_t *__restrict__ hp, uint32_t H,
                                              uint32_t *s_cnt, uint32_t *s_red, uint64_t exit_thr)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr uint32_t HB = kWideWarps * kWideHashesPerWarp;
    const uint64_t slack = exit_thr ? (exit_thr - 1) / d.max_run : 0;
    if (exit_thr && slack >= H)
        return 0; // hopeless before the first probe (see probe_chunk)
    uint32_t next_check = exit_thr ? H - (uint32_t)slack : 0xffffffffu;
    for (uint32_t b0 = 0; b0 < H; b0 += HB)
    {
        const uint32_t hb = min(H, b0 + wib * kWideHashesPerWarp), he = min(H, hb + kWideHashesPerWarp);
        if (hb < he)
            for (uint32_t col0 = 0; col0 < d.tbins; col0 += 512u * CP)
                wide_block<CP, U, GEN>(d, sch, hp, hb, he, col0, s_cnt, lane);
        const uint32_t done = min(H, b0 + HB);
        if (done >= next_check && done < H) // CTA-uniform
        {
            __syncthreads();
            uint32_t m = 0;
            for (uint32_t i = threadIdx.x; i < d.bins; i += blockDim.x)
                m = max(m, s_cnt[i]);
            m = __reduce_max_sync(0xffffffffu, m);
            if (lane == 0)
                s_red[wib] = m;
            __syncthreads();
            m = 0;
#pragma unroll
            for (int q = 0; q < kWideWarps; ++q)
                m = max(m, s_red[q]);
            if ((uint64_t)d.max_run * ((uint64_t)m + (H - done)) < exit_thr)
                return done;
            next_check = done + 2 * HB;
        }
    }
    return H;
}

template <bool GEN>
__global__ void __launch_bounds__(32 * kWideWarps, 2) ixf_query_large_kernel(QueryArgs a)
{
    if (!sm_filter_keep(a.smf))
        return;
    extern __shared__ uint32_t s_cnt_dyn[]; // tbins counters
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_red[kWideWarps];
    const uint32_t n_items = a.items ? *a.n_items_ptr : a.n_items_direct;
    unsigned long long bytes = 0, items = 0, skipped = 0;
    while (true)
    {
        __syncthreads(); // the previous item's bin scan is over: counters and s_item may be rewritten
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
        const uint64_t thr = a.thr_read ? a.thr_read[read] : (H < a.lut_len ? a.thr_lut[H] : ~0ULL);
        const uint64_t exit_thr = a.early_exit ? thr : 0;
        uint32_t Hp;
        if (d.tbins <= 1024)
            Hp = wide_item<2, 2, GEN>(d, a.scheme, hp, H, s_cnt_dyn, s_red, exit_thr);
        else
            Hp = wide_item<4, 1, GEN>(d, a.scheme, hp, H, s_cnt_dyn, s_red, exit_thr);
        __syncthreads();
        scan_bins(a, d, read, thr, s_cnt_dyn, threadIdx.x, blockDim.x);
        bytes += (unsigned long long)Hp * 3ull * d.tbins + 8ull * Hp;
        skipped += H - Hp;
        ++items;
    }
    if (threadIdx.x == 0 && items)
    {
        atomicAdd(a.stat_bytes, bytes);
        atomicAdd(a.stat_items, items);
        if (skipped)
            atomicAdd(a.stat_skipped, skipped);
    }
}

// bulk_count of a single IXF for a single value list (parity entry point txr_ixf_bulk_count)
template <bool GEN>
__global__ void __launch_bounds__(256) ixf_bulk_count_kernel(IxfDev d, IxfScheme sch, const uint64_t *values, uint32_t n, uint32_t *counts)
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
        probe_chunk<kQueryUnroll, GEN>(d, sch, values, n, off, width >> 4, s_cnt_dyn + off, lane);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < d.bins; i += blockDim.x)
        counts[i] = s_cnt_dyn[i];
}

// ---- level queues grouped by IXF ----
// Work items of a level are appended in the order reads finish, i.e. all child IXFs interleaved: the rows a warp
// gathers then come from the whole level (GBs) and nothing is reused.  Grouping the queue by IXF makes the few
// thousand warps in flight work on two or three IXFs at a time; a child IXF of a 1,000-genome index is tens of MB
// and stays in the 126 MB L2 while its reads are processed, so most of its rows are fetched from HBM once per
// batch instead of once per probe.  Counting sort in three small launches (counts are device-resident).
__global__ void __launch_bounds__(256) items_hist_kernel(const uint2 *items, const uint32_t *n_ptr, uint32_t cap, uint32_t *hist)
{
    const uint32_t n = min(*n_ptr, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&hist[items[i].y], 1u);
}

// exclusive scan of hist[0..n) in place, one CTA
__global__ void __launch_bounds__(1024) items_scan_kernel(uint32_t *hist, uint32_t n)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? hist[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d)
                incl += up;
        }
        if (lane == 31)
            s_warp[wib] = incl;
        __syncthreads();
        if (wib == 0)
        {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d)
                    w += up;
            }
            s_warp[lane] = w; // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = s_carry + (wib ? s_warp[wib - 1] : 0);
        if (i < n)
            hist[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023)
            s_carry = before + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) items_scatter_kernel(const uint2 *items, const uint32_t *n_ptr, uint32_t cap, uint32_t *offs, uint2 *out)
{
    const uint32_t n = min(*n_ptr, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint2 it = items[i];
        out[atomicAdd(&offs[it.y], 1u)] = it;
    }
}

cudaError_t launch_sort_items(const uint2 *items, const uint32_t *n_ptr, uint32_t cap, uint32_t *hist, uint32_t n_ixf, uint2 *out,
                              int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)n_ixf * 4, st);
    if (e != cudaSuccess)
        return e;
    items_hist_kernel<<<sm_count * 2, 256, 0, st>>>(items, n_ptr, cap, hist);
    items_scan_kernel<<<1, 1024, 0, st>>>(hist, n_ixf);
    items_scatter_kernel<<<sm_count * 2, 256, 0, st>>>(items, n_ptr, cap, hist, out);
    return cudaGetLastError();
}

// ---- launchers ----
// CTAs per SM of the persistent probe grids (QueryArgs::ctas_per_sm): 8 = all 32 warps an SM can hold at 56 registers.
// The random-line rate of HBM is already saturated by 16 warps per SM (profiles/r1_gather_bench2.json); the engine
// lowers the value when the hash / dedup kernels of the next batch run beside the probes.
static int query_ctas(const QueryArgs &a) { return a.ctas_per_sm <= 0 ? 8 : a.ctas_per_sm > 16 ? 16 : a.ctas_per_sm; }

cudaError_t launch_query_small(const QueryArgs &a, int sm_count, cudaStream_t st)
{
    // probe steps in flight per warp (QueryArgs::unroll): 2 fills the memory system when 7-8 CTAs per SM run; with fewer
    // CTAs (the hash kernel of the next batch beside the probes) 3 or 4 keep as many bytes in flight
    const int grid = sm_count * query_ctas(a);
    if (a.generic)
        ixf_query_small_kernel<true, kQueryUnroll, 1><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else if (a.unroll == 1 && (query_ctas(a) >= 16 || a.regs32)) // 32 registers: all 64 warps an SM can hold
        ixf_query_small_kernel<false, 1, 16><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else if (a.unroll == 1) // one step in flight per warp, 40 registers: up to 12 CTAs (48 warps) per SM
        ixf_query_small_kernel<false, 1, 12><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else if (a.unroll == 3)
        ixf_query_small_kernel<false, 3, 1><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else if (a.unroll >= 4)
        ixf_query_small_kernel<false, 4, 1><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else if (query_ctas(a) >= 8)
        ixf_query_small_kernel<false, kQueryUnroll, 8><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    else
        ixf_query_small_kernel<false, kQueryUnroll, 7><<<grid, 32 * kQueryWarps, 0, st>>>(a);
    return cudaGetLastError();
}

// partitioned root level: `a.work` and `a.hist` must be zero on entry (the engine clears them with the counters)
cudaError_t launch_root_partitioned(const QueryArgs &q, const RootPartArgs &a, int sm_count, cudaStream_t st)
{
    root_part_hist_kernel<<<sm_count * 4, 256, 0, st>>>(a);
    root_part_scan_kernel<<<1, 1024, 0, st>>>(a);
    root_part_scatter_kernel<<<sm_count * 4, 256, 0, st>>>(a);
    root_part_probe_kernel<<<sm_count * query_ctas(q), 32 * kQueryWarps, 0, st>>>(a);
    root_part_scan_kernel2<<<sm_count * 8, 32 * kQueryWarps, 0, st>>>(q, a);
    return cudaGetLastError();
}

// widest IXF the CTA-per-item kernel can hold counters for (32-bit counters in at most 200 KB of dynamic shared memory)
uint32_t query_large_max_tbins() { return 200u * 1024u / 4u; }

cudaError_t launch_query_large(const QueryArgs &a, int sm_count, uint32_t max_tbins, cudaStream_t st)
{
    const size_t smem = (size_t)max_tbins * 4;
    if (max_tbins > query_large_max_tbins())
        return cudaErrorInvalidValue; // txr_index_upload rejects such indexes
    if (smem > 48 * 1024)
    {
        const cudaError_t e = a.generic ? cudaFuncSetAttribute(ixf_query_large_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                        : cudaFuncSetAttribute(ixf_query_large_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    // CTAs per SM: as many as the counters allow, up to 8 warps x 6 = 48 warps (registers cap it at 4-5 anyway)
    const int by_smem = (int)std::max<size_t>(1, (200u * 1024u) / std::max<size_t>(smem + 64, 1));
    const int ctas = std::min(a.ctas_per_sm > 0 ? a.ctas_per_sm : 6, std::min(by_smem, 6));
    if (a.generic)
        ixf_query_large_kernel<true><<<sm_count * ctas, 32 * kWideWarps, smem, st>>>(a);
    else
        ixf_query_large_kernel<false><<<sm_count * ctas, 32 * kWideWarps, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bulk_count(const IxfDev &d, const IxfScheme &sch, const uint64_t *values, uint32_t n, uint32_t *counts, cudaStream_t st)
{
    const size_t smem = (size_t)d.tbins * 4;
    const bool gen = !ixf_scheme_is_default(sch);
    if (smem > 48 * 1024)
    {
        const cudaError_t e = gen ? cudaFuncSetAttribute(ixf_bulk_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(ixf_bulk_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    if (gen)
        ixf_bulk_count_kernel<true><<<1, 256, smem, st>>>(d, sch, values, n, counts);
    else
        ixf_bulk_count_kernel<false><<<1, 256, smem, st>>>(d, sch, values, n, counts);
    return cudaGetLastError();
}
} // namespace txr
