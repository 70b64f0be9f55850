// device_types.cuh -- plain structs shared by the kernels and the host engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ixf_arith.cuh"

namespace txr
{
constexpr int kTileWindows = 1024;         // windows per warp tile in the hash kernels (32 per lane)
constexpr uint64_t kEmptyKey = ~0ULL;      // sentinel of the dedup tables
constexpr int kMaxMinimiserValues = 96;    // taxor build --window-size <= 96 (taxor_build.cpp:78-80)
constexpr uint32_t kSmallRowBytes = 512;   // IXFs with tbins <= this are probed by one warp per (read, IXF)

// SM partition for the persistent (work-stealing) kernels: a CTA keeps running iff mod == 0 or (its SM id % mod) lies
// in [lo, hi); otherwise it returns before taking any work.  With complementary filters the ALU-bound hash / dedup
// kernels of batch i+1 and the DRAM-bound probe kernels of batch i run side by side on disjoint SMs (engine.cu).
struct SmFilter
{
    uint32_t mod, lo, hi;
};
#if defined(__CUDACC__)
// adaptive grid of the hash-stage kernels: true = this CTA should leave (the probes own the GPU and this CTA is beyond the share
// the hash stage gets beside them)
__device__ __forceinline__ bool yield_to_probes(const uint32_t *probe_flag, uint32_t small_grid)
{
    return probe_flag != nullptr && blockIdx.x >= small_grid && *reinterpret_cast<const volatile uint32_t *>(probe_flag) != 0u;
}
__device__ __forceinline__ bool sm_filter_keep(const SmFilter &f)
{
    if (f.mod == 0)
        return true;
    uint32_t id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    const uint32_t r = id % f.mod;
    return r >= f.lo && r < f.hi;
}
#endif

// ---- kernel #1 (hashing) ----
struct HashArgs
{
    const uint64_t *words;     // 2-bit packed reads, MSB-first, one zero pad word per read
    const uint64_t *word_off;  // [n_reads] first word of each read
    const uint32_t *len;       // [n_reads] bases
    const uint64_t *out_off;   // [n_reads+1] capacity offsets into `out`
    uint64_t *out;             // hashes
    uint32_t *n_out;           // [n_reads] number of hashes written (raw emissions / k-mers)
    uint32_t n_reads;
    uint32_t *work_counter;    // dynamic read assignment
    uint32_t *overflow;        // set to 1 when a read needed more than its capacity
    uint64_t kmer_seed;        // k-mer mode: adjust_seed(k)
    int k, s, t;               // generic kernel only
    SmFilter smf;
    int ctas_per_sm;           // host side: CTAs per SM of the persistent grid (0: fill the SM)
    int window;                // minimiser mode: k-mer values per window, window_size - k + 1 (2..kMaxMinimiserValues)
    int min_blocks;            // host side: register variant of the syncmer kernel (0/4 default, 5 = 102 registers)
    // adaptive grid (overlap): the kernel is launched with a grid for the whole GPU; CTAs beyond `small_grid` leave at once when the
    // probe kernels of the batch before are running (`*probe_flag` != 0, set and cleared in stream order around every query stage)
    const uint32_t *probe_flag;
    uint32_t small_grid;
    // fused per-read distinct set (syncmer_kernel only; the ankerl::set of syncmer.cpp:145 built while hashing):
    uint32_t fuse_dedup;       // 1: reads whose capacity is <= kFuseMaxCap write DISTINCT hashes + hash_count directly
    uint32_t fuse_max_keys;    // distinct keys the warp table takes before the read is handed over (<= kWarpMaxKeys; tests lower it)
    uint32_t *hash_count;      // [n_reads] distinct (after the FracMin scaling filter)
    uint32_t *deferred;        // reads whose distinct set outgrew the warp table: raw tail appended, dedup'd by a CTA later
    uint32_t *n_deferred;
    uint32_t scaling;          // FracMin scaling (1 = off), applied before the set insert
    double scaling_limit;
};
constexpr uint32_t kFuseMaxCap = 2048;      // == the `ids_small` class of the engine
constexpr int kWarpSlots = 2048;            // per-warp dedup table (u32 slots in shared memory)
constexpr uint32_t kWarpMaxKeys = kWarpSlots * 3 / 4; // 1536: load factor 3/4; index + 1 fits the low 11 bits of a slot

struct DedupArgs
{
    const uint32_t *read_ids;  // reads of this size class (nullptr: identity)
    uint32_t n_ids;
    const uint64_t *out_off;
    uint64_t *hashes;          // in place: raw -> distinct
    const uint32_t *n_raw;
    uint32_t *hash_count;      // [n_reads] distinct (after the FracMin scaling filter)
    uint64_t *gtable;          // global-memory tables (class C), pre-filled with kEmptyKey
    const uint64_t *gtable_off;
    uint32_t scaling;          // 1 = off
    double scaling_limit;      // double(UINT64_MAX) / double(scaling)
    SmFilter smf;
    int ctas_per_sm;           // host side: CTAs per SM of the warp-per-read grid (0: default)
    const uint32_t *probe_flag; // adaptive grid, as in HashArgs
    uint32_t small_grid;
};

// ---- build side: distinct hash set of a user bin from the raw hash lists of its sequence segments ----
struct BinSetArgs
{
    const uint64_t *hashes;      // raw hash lists of the segments (kernel #1 output, capacity layout)
    const uint64_t *out_off;     // [n_segments+1] capacity offsets
    const uint32_t *n_raw;       // [n_segments]
    const uint32_t *seg_bin;     // [n_segments] user bin (batch-local) of a segment
    uint32_t n_segments;
    uint64_t *tables;            // open-addressing tables of all bins, filled with kEmptyKey
    const uint64_t *table_off;   // [n_bins+1] first slot of a bin's table (sizes are powers of two)
    uint32_t n_bins;
    uint32_t *bin_has_empty_key; // [n_bins] the key equal to the sentinel was seen
    uint64_t *out;               // distinct hashes, bin b at out[out_bin_off[b] ..)
    const uint64_t *out_bin_off; // [n_bins+1] (capacity: raw count of the bin)
    uint32_t *bin_count;         // [n_bins] distinct hashes written
    uint32_t *work;              // work-stealing cursor (segments)
    uint32_t scaling;
    double scaling_limit;
};

// ---- kernel #2 (IXF probe / count / threshold / compaction) ----
struct IxfDev
{
    const uint8_t *fp;         // fp[slot * tbins + bin], 256-byte aligned, tbins % 64 == 0
    uint64_t seed;
    uint32_t seg_len;
    uint32_t tbins;            // row stride in bytes
    uint32_t bins;             // counting-vector size
    uint32_t meta_off;         // first entry of this IXF in the per-bin metadata arrays
    uint32_t max_run;          // longest run of technical bins that belong to one user bin (>= 1)
    uint32_t count_len;        // FUSE3 only: segment_count * seg_len, the range of the first slot
};

enum : uint8_t { kBinMid = 0, kBinRunEnd = 1, kBinMerged = 2 };

struct QueryArgs
{
    const IxfDev *ixf;
    const int32_t *bin_ub;         // user bin id of a run end
    const int32_t *bin_child;      // child IXF of a merged bin
    const uint32_t *bin_run_begin; // first bin of the split run that ends here
    const uint8_t *bin_kind;

    const uint64_t *hashes;
    const uint64_t *hash_off;      // [n_reads+1] (capacity offsets)
    const uint32_t *hash_count;
    const uint64_t *thr_lut;       // threshold by hash_count
    uint32_t lut_len;
    const uint64_t *thr_read;      // per-read thresholds (FracMinHash model: depends on the read length too); nullptr: use the LUT

    const uint2 *items;            // (read, ixf) work items of this level; nullptr: item i = (i, 0)
    const uint32_t *n_items_ptr;   // device counter written by the previous level
    uint32_t n_items_direct;
    uint32_t items_cap;            // capacity of `items` (a level that overflowed is re-run by the host)
    uint32_t *cursor;              // work-stealing cursor of this launch

    uint2 *next_small;             // next level, children with tbins <= kSmallRowBytes
    uint32_t *next_small_n;
    uint2 *next_large;
    uint32_t *next_large_n;
    uint32_t next_cap;

    uint32_t *hit_read;
    int32_t *hit_ub;
    uint32_t *hit_cnt;
    uint32_t *n_hits;
    uint32_t hit_cap;

    SmFilter smf;
    int ctas_per_sm;                // host side: CTAs per SM of the persistent probe grid (0: fill the SM)
    uint32_t early_exit;            // 1: stop probing an item once no user bin can reach the threshold any more
    uint32_t l2_hints;              // 1: items are grouped by IXF, use the L2 eviction-priority plan (query_kernels.cu)
    uint32_t unroll;                // probe steps in flight per warp of the small kernel (0/2 default, 1, 3, 4)
    uint32_t regs32;                // unroll == 1: take the 32-register build (16 CTAs per SM fit) whatever the grid
    uint32_t l2_sector64;           // 1: rows of 64-byte IXFs are loaded with the .L2::64B prefetch-size hint; 2: only below the root
    uint32_t generic;               // 1: the index was uploaded with a non-default probe arithmetic (`scheme`)
    IxfScheme scheme;
    unsigned long long *stat_bytes; // algorithmic bytes: sum H*3*tbins + 8*H
    unsigned long long *stat_items;
    unsigned long long *stat_skipped; // hashes NOT probed thanks to the early exit (their bytes are not in stat_bytes)
};

// ---- kernel #2, root level of a batch, slot-partitioned (query_kernels.cu: "partitioned root") ----
struct RootPartArgs
{
    IxfDev root;
    const uint64_t *hashes;        // per-read lists (as in QueryArgs)
    const uint64_t *hash_off;
    const uint32_t *hash_count;
    uint32_t n_reads;
    uint32_t log2_parts;           // partitions = 1 << log2_parts, by the top bits of the segment-0 slot
    uint32_t *hist;                // [parts + 1] element counts per partition, then their sum
    uint32_t *cursor;              // [parts] write cursors (start at the exclusive scan of hist)
    uint64_t *part_hash;           // elements grouped by partition ...
    uint32_t *part_read;           // ... and the read each one belongs to
    uint32_t *counts;              // [n_reads * root.tbins / 2] 16-bit counters packed in pairs
    uint32_t *work;                // work-stealing cursors: [0] hist, [1] scatter, [2] probe, [3] scan
};
} // namespace txr
