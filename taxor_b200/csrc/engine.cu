// engine.cu -- host engine + C ABI (include/taxor_b200.h) of the B200-native `taxor search` hot path.
//
// Replaces the chunk loop + worker of search_single (src/main/taxor_search.cpp:196-326) for batches of reads:
//   H2D packed reads -> kernel #1 (hash) -> dedup -> kernel #2 once per HIXF level -> D2H hits -> host: order the
//   hits of every read in the reference's DFS pre-order (hixf.hpp:313-338), thresholds, 0.8*max filter flags.
// The index is re-laid-out once into HBM (txr_index_upload: staged, multi-threaded; txr_index_clone: device to device);
// reads stream through `n_slots` pipeline slots: each slot's own stream carries its copies, the probe kernels of all
// batches run in batch order on one compute stream, and the hash + dedup kernels of batch i+1 run on a second stream beside
// the probes of batch i where that was measured to win (enqueue_kernels: overlap, adaptive grids).
#include "../../include/taxor_b200.h"
#include "device_types.cuh"
#include "ixf_arith.cuh"
#include "threshold.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <vector>

namespace txr
{
cudaError_t launch_syncmer(const HashArgs &a, int sm_count, cudaStream_t st);
bool syncmer_has_fast_kernel(int k, int s, int t);
cudaError_t launch_kmer(const HashArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_minimiser(const HashArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_dedup_warp(const DedupArgs &a, int sm_count, uint32_t *work_counter, uint32_t *deferred, uint32_t *n_deferred,
                              cudaStream_t st);
cudaError_t launch_dedup_deferred(const DedupArgs &a, int sm_count, const uint32_t *n_deferred, cudaStream_t st);
cudaError_t launch_dedup_medium(const DedupArgs &a, cudaStream_t st);
cudaError_t launch_dedup_global(const DedupArgs &a, cudaStream_t st);
cudaError_t launch_filter(const DedupArgs &a, cudaStream_t st);
cudaError_t launch_query_small(const QueryArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_query_large(const QueryArgs &a, int sm_count, uint32_t max_tbins, cudaStream_t st);
uint32_t query_large_max_tbins();
size_t pack_plain_words(const char *ascii, size_t n_words, uint64_t *dst); // pack_simd.cpp
cudaError_t launch_sort_items(const uint2 *items, const uint32_t *n_ptr, uint32_t cap, uint32_t *hist, uint32_t n_ixf, uint2 *out,
                              int sm_count, cudaStream_t st);
cudaError_t launch_root_partitioned(const QueryArgs &q, const RootPartArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_binset(const BinSetArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_bulk_count(const IxfDev &d, const IxfScheme &sch, const uint64_t *values, uint32_t n, uint32_t *counts, cudaStream_t st);
} // namespace txr

using namespace txr;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int set_error(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define CU(call)                                                                                          \
    do                                                                                                    \
    {                                                                                                     \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return set_error(TXR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(call)           \
    do                      \
    {                       \
        int rc_ = (call);   \
        if (rc_ != TXR_OK)  \
            return rc_;     \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// small RAII buffers that only ever grow
// ---------------------------------------------------------------------------------------------------------
struct DevBuf
{
    void *p{nullptr};
    size_t cap{0};
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return TXR_OK;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return TXR_OK;
    }
    int ensure_exact(size_t bytes) // no growth slack: the index arena (an eighth of a 100 GB index is not small change)
    {
        if (bytes <= cap)
            return TXR_OK;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        CU(cudaMalloc(&p, bytes ? bytes : 1));
        cap = bytes;
        return TXR_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct PinBuf
{
    void *p{nullptr};
    size_t cap{0};
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return TXR_OK;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
        return TXR_OK;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// ---------------------------------------------------------------------------------------------------------
// index in HBM
// ---------------------------------------------------------------------------------------------------------
struct DeviceIndex
{
    bool loaded{false};
    std::vector<IxfDev> ixf;            // host copy of the descriptors
    DevBuf arena;                       // all fingerprint rows
    DevBuf d_ixf, d_bin_ub, d_bin_child, d_bin_run_begin, d_bin_kind;
    std::vector<uint32_t> dfs_rank;     // user bin -> position in the reference's DFS pre-order
    uint32_t depth{0};                  // number of tree levels
    uint32_t max_tbins{0};
    bool any_large{false};
    uint64_t n_user_bins{0};
    uint64_t fp_bytes{0};
    IxfScheme scheme{kIxfSlotsXor3, kIxfMixAddSeed, kIxfFpFold32, 21u, 42u}; // probe arithmetic the index was built with
    bool generic{false};                // scheme != the prototype's: the descriptor-driven kernels run
    void release()
    {
        arena.release();
        d_ixf.release();
        d_bin_ub.release();
        d_bin_child.release();
        d_bin_run_begin.release();
        d_bin_kind.release();
        loaded = false;
    }
};

// ---------------------------------------------------------------------------------------------------------
// a batch of reads as the kernels see it
// ---------------------------------------------------------------------------------------------------------
struct BatchMeta // host side description, built once per batch
{
    uint64_t first_read{0};
    uint32_t n_reads{0};
    uint64_t n_bases{0};
    uint64_t first_word{0}, n_words{0};
    uint64_t total_cap{0};
    uint64_t max_cap{0};
    std::vector<uint64_t> word_off; // relative to first_word
    std::vector<uint32_t> len;
    std::vector<uint64_t> out_off;  // n_reads + 1
    std::vector<uint32_t> ids_small, ids_medium, ids_global;
    std::vector<uint64_t> gtable_off; // ids_global.size() + 1
};

struct BatchDev // device copies of the metadata
{
    DevBuf word_off, len, out_off, ids_small, ids_medium, ids_global, gtable_off;
};

enum CounterSlot : int
{
    C_NHITS = 0,
    C_HASH_OVERFLOW = 1,
    C_HASH_WORK = 2,
    C_DEDUP_WORK = 3,
    C_DEDUP_DEFERRED = 4,
    C_LEVEL0 = 8, // per level: n_small, n_large, cursor_small, cursor_large
    C_PER_LEVEL = 4,
    C_MAX_LEVELS = 32,
    C_STATS = C_LEVEL0 + C_PER_LEVEL * C_MAX_LEVELS, // three u64: bytes, items, skipped hashes (8-byte aligned: index is even)
    C_TOTAL = C_STATS + 6
};

struct Slot
{
    cudaStream_t stream{nullptr};
    // events: 0 h2d start, 1 h2d done (copy stream) | 6 compute start, 2 hash done, 3 dedup done, 4 query done,
    // 7 compute done (compute stream) | 5 d2h done (copy stream)
    cudaEvent_t ev[10]{}; // ... 8 query start (compute stream), 9 hash stage + thresholds done (hash stream)
    DevBuf words;
    BatchDev meta;
    DevBuf hashes, n_raw, hash_count, gtable, deferred;
    DevBuf queues; // per level >= 1: small[cap], large[cap]; then two more arrays: the current level grouped by IXF
    DevBuf ixf_hist;
    DevBuf part_hash, part_read, part_counts, part_ctl; // partitioned root level (query_kernels.cu)
    DevBuf hit_read, hit_ub, hit_cnt;
    DevBuf counters;
    DevBuf thr;    // per-read thresholds (FracMinHash model only)
    PinBuf h_meta, h_counters, h_hash_count, h_hits, h_thr;
    uint32_t queue_cap{0}, hit_cap{0};
    // state of the batch in flight
    const BatchMeta *bm{nullptr};
    const BatchDev *bd{nullptr};
    const uint64_t *d_words{nullptr};
    bool busy{false};
    bool split_used{false}, ran_query{true};
    bool fused{false}; // the hash kernel of the batch in flight built the distinct sets itself
    bool first{true}, last{true}; // position of the batch in its search call (overlap shapes)
};

struct ResultStore
{
    std::vector<uint32_t> hash_count;
    std::vector<uint64_t> threshold, hit_begin;
    std::vector<int64_t> user_bin;
    std::vector<uint32_t> count;
    std::vector<uint8_t> keep;
    void clear()
    {
        hash_count.clear();
        threshold.clear();
        hit_begin.clear();
        user_bin.clear();
        count.clear();
        keep.clear();
    }
};

struct txr_reads
{
    DevBuf words;
    uint64_t n_reads{0};
    std::vector<BatchMeta> batches;
    std::vector<std::unique_ptr<BatchDev>> dev;
};

struct txr_ctx
{
    int device{0};
    int sm_count{148};
    uint64_t max_batch_reads{262144};
    uint64_t max_batch_bases{3000000000ull};
    int n_slots{4}; // pipeline slots: copy of batch i+2, hash stage of batch i+1 and probes of batch i in flight together
    std::vector<std::unique_ptr<Slot>> slots;
    DeviceIndex index;
    bool have_params{false};
    txr_params params{};
    Thresholder thresholder;
    uint64_t kmer_seed{0};
    bool sort_items{true};     // group level queues by IXF (TXR_SORT_ITEMS=0 disables, for A/B measurements)
    bool early_exit{true};     // exact early exit of kernel #2 (TXR_EARLY_EXIT=0 disables)
    bool l2_hints{true};       // L2 eviction-priority plan for small child IXFs (TXR_L2_HINTS=0 disables)
    // distinct set built inside the syncmer kernel (TXR_FUSE_DEDUP=1).  Off by default: measured slower (hash 13.1 + dedup
    // 7.4 ms -> 25.0 ms per 1 M reads, profiles/r2_b_*): the hash kernel runs 16 warps per SM at 124 registers and is issue
    // bound, so the claim rounds cost more there than in the 24-warp dedup kernel; neither side is DRAM bound, the saved
    // 15 GB round trip buys nothing.
    bool fuse_dedup{false};
    // .L2::64B loads for 64-byte rows (TXR_L2_SECTOR64: 0 off, 1 all levels, 2 below the root only).  On: the rows of a T = 64
    // index then cost two DRAM sectors instead of a 128-byte line; the step gains little (93.5 -> 92.9 ms, the root level is bound by
    // random accesses per second, not by bytes) but the DRAM traffic per probe byte drops from 1.8x to about 1x.
    uint32_t l2_sector64{1};
    uint32_t query_regs32{0};  // TXR_QUERY_REGS=32: the 32-register build of the one-step probe kernel whatever the CTA count
    bool adaptive{true};       // TXR_ADAPTIVE=0: fixed small grids for the hash stage beside the probes
    double ramp{1.8};          // growth of the batch sizes of a host-fed search (TXR_RAMP)
    double ramp_cum{0.35};     // TXR_RAMP_CUM: cap of a batch as a fraction of the reads submitted before it (0: off; 93.5 -> 91.7 ms e2e, profiles/r2_m_ramp_cum.txt)
    int hash_regs{0};          // TXR_HASH_REGS=5: the 102-register variant of the syncmer kernel (5 CTAs per SM)
    uint32_t query_unroll{0};  // TXR_QUERY_UNROLL: probe steps in flight per warp (experiments with fewer probe CTAs per SM)
    uint32_t fuse_max_keys{kWarpMaxKeys}; // TXR_FUSE_MAX_KEYS lowers it (tests: forces the hand-over to the CTA-per-read kernel)
    int root_partition{0};     // root level grouped by segment-0 slot: 0 off (default: measured slower), 1 auto, 2 always (TXR_ROOT_PARTITION)
    bool per_read_thr{false};  // FracMinHash model: the threshold depends on hash_count AND the read length
    std::vector<uint64_t> lut; // threshold by hash_count (host)
    DevBuf d_lut;
    uint64_t d_lut_len{0};
    ResultStore result;
    txr_timing timing{};
    // scratch for the parity entry points
    std::vector<uint64_t> hb_off, hb_hashes;
    std::vector<uint64_t> ub_off, ub_hashes; // txr_hash_user_bins result
    DevBuf scratch_a, scratch_b;
    DevBuf probe_flag;         // != 0 while the probe kernels of some batch run (adaptive grids of the hash stage, overlap)
    DevBuf binset[8]; // txr_hash_user_bins: segment bins, tables, table offsets, flags, output, output offsets, counts, cursor
    cudaStream_t primary{nullptr};     // caller's stream: every search forks from it and joins back into it
    cudaStream_t compute{nullptr};     // the query kernels of all batches run here, in batch order; slot streams only carry copies
    cudaStream_t compute_hash{nullptr}; // overlap mode: hash + dedup (ALU bound) of batch i+1 beside the query (DRAM bound) of batch i
    int overlap_mode{2};               // TXR_OVERLAP: 0 off, 1 always, 2 auto (overlap_applies): hash stage of batch i+1 beside the probes of batch i
    int shape_unroll{0};               // probe steps in flight per warp for the batch being enqueued
    int adaptive_small_hash{0}, adaptive_small_dedup{0}; // != 0: CTAs per SM the hash / dedup kernel keeps while probes run (adaptive grid)
    int query_ctas{0}, hash_ctas{0}, dedup_ctas{0}; // CTAs per SM (0: defaults for the mode)
    int level_ctas{0};                              // probe kernels of the levels below the root (0: same as query_ctas)
    // overlap by SM partition: of every `sm_mod` consecutive SM ids the first `sm_hash` run hash + dedup, the rest the
    // probe kernels (0: share the SMs with small grids instead)
    uint32_t sm_mod{0}, sm_hash{0};
    SmFilter smf_hash{0, 0, 0}, smf_query{0, 0, 0}; // filters of the batch being enqueued ...
    int shape_query{0}, shape_level{0}, shape_hash{0}, shape_dedup{0}; // ... and its CTAs per SM (0: fill the SM)
    cudaEvent_t fork_ev{nullptr}, join_ev{nullptr};
    bool trace{false};                 // TXR_TRACE=1
    cudaEvent_t trace_ev{nullptr};
    std::chrono::steady_clock::time_point trace_t0{};
};

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
static inline uint64_t windows_of(uint64_t len, int k) { return len >= (uint64_t)k ? len - k + 1 : 0; }

// Upper bound on the hashes kernel #1 can emit for a read.  Two selected windows j1 < j2 (states p = j+t-1) are at
// least min(t-1, k-s+1-t)+1 apart: p2 became the state either by a strictly smaller arrival -- then it arrived at
// window p2-(k-s) > j1, since from its arrival on the state can only be at or right of p2, so
// j2-j1 > k-s+1-t -- or by a rescan, which happens at window (previous state)+1 >= p1+1 = j1+t, so j2-j1 >= t.
static uint64_t hash_capacity(const txr_params &p, uint64_t len)
{
    const uint64_t w = windows_of(len, p.kmer_size);
    if (!p.use_syncmer)
        return w;
    const int wn = p.kmer_size - p.syncmer_size + 1;
    const int spacing = std::min<int>(p.t_syncmer - 1, wn - p.t_syncmer) + 1;
    return (w + spacing - 1) / std::max(spacing, 1) + 1;
}

static uint64_t next_pow2(uint64_t x)
{
    uint64_t p = 1;
    while (p < x)
        p <<= 1;
    return p;
}

static int ensure_slots(txr_ctx *c)
{
    if (!c->compute)
    {
        CU(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->compute_hash, cudaStreamNonBlocking));
    }
    if (!c->probe_flag.p)
    {
        TRY(c->probe_flag.ensure(4));
        CU(cudaMemset(c->probe_flag.p, 0, 4));
    }
    while ((int)c->slots.size() < c->n_slots)
    {
        auto s = std::make_unique<Slot>();
        CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        for (auto &e : s->ev)
            CU(cudaEventCreate(&e));
        c->slots.push_back(std::move(s));
    }
    return TXR_OK;
}

// threshold LUT covering hash counts [0, need)
static int ensure_lut(txr_ctx *c, uint64_t need)
{
    if (need <= c->lut.size() && c->d_lut_len >= need)
        return TXR_OK;
    const uint64_t old = c->lut.size();
    if (need > old)
    {
        uint64_t grow = std::max<uint64_t>(need, old + old / 2 + 1024);
        c->lut.resize(grow);
        for (uint64_t i = old; i < grow; ++i)
            c->lut[i] = c->thresholder.get(i, 1.0);
    }
    // all streams must be idle before the table is replaced
    CU(cudaDeviceSynchronize());
    TRY(c->d_lut.ensure(c->lut.size() * 8));
    CU(cudaMemcpy(c->d_lut.p, c->lut.data(), c->lut.size() * 8, cudaMemcpyHostToDevice));
    c->d_lut_len = c->lut.size();
    return TXR_OK;
}

// Builds the host metadata for reads [first, first+n): word offsets relative to the first word of the batch,
// hash capacities and dedup size classes.
static void build_batch_meta(const txr_ctx *c, const uint64_t *word_off, const uint32_t *len, uint64_t first,
                             uint32_t n, BatchMeta &m)
{
    m.first_read = first;
    m.n_reads = n;
    m.first_word = word_off[first];
    m.word_off.resize(n);
    m.len.assign(len + first, len + first + n);
    m.out_off.resize((size_t)n + 1);
    m.ids_small.clear();
    m.ids_medium.clear();
    m.ids_global.clear();
    m.gtable_off.assign(1, 0);
    uint64_t bases = 0, cap_sum = 0, max_cap = 0;
    const bool dedup = c->params.use_syncmer;
    for (uint32_t i = 0; i < n; ++i)
    {
        m.word_off[i] = word_off[first + i] - m.first_word;
        const uint64_t cap = hash_capacity(c->params, len[first + i]);
        m.out_off[i] = cap_sum;
        cap_sum += cap;
        max_cap = std::max(max_cap, cap);
        bases += len[first + i];
        if (dedup)
        {
            if (cap <= 2048)
                m.ids_small.push_back(i);
            else if (cap <= 8192)
                m.ids_medium.push_back(i);
            else
            {
                m.ids_global.push_back(i);
                m.gtable_off.push_back(m.gtable_off.back() + next_pow2(2 * cap));
            }
        }
    }
    m.out_off[n] = cap_sum;
    m.total_cap = cap_sum;
    m.max_cap = max_cap;
    m.n_bases = bases;
    const uint64_t last = first + n - 1;
    m.n_words = word_off[last] + txr_packed_words(len[last]) - m.first_word;
}

static size_t meta_bytes(const BatchMeta &m)
{
    return m.word_off.size() * 8 + m.len.size() * 4 + m.out_off.size() * 8 + m.ids_small.size() * 4 + m.ids_medium.size() * 4 +
           m.ids_global.size() * 4 + m.gtable_off.size() * 8 + 64;
}

// copies the metadata arrays to the device (through `stage` when given, which must be pinned and large enough)
static int upload_batch_meta(const BatchMeta &m, BatchDev &d, cudaStream_t st, uint8_t *stage)
{
    struct Item
    {
        DevBuf *dst;
        const void *src;
        size_t bytes;
    };
    Item items[] = {{&d.word_off, m.word_off.data(), m.word_off.size() * 8},
                    {&d.len, m.len.data(), m.len.size() * 4},
                    {&d.out_off, m.out_off.data(), m.out_off.size() * 8},
                    {&d.ids_small, m.ids_small.data(), m.ids_small.size() * 4},
                    {&d.ids_medium, m.ids_medium.data(), m.ids_medium.size() * 4},
                    {&d.ids_global, m.ids_global.data(), m.ids_global.size() * 4},
                    {&d.gtable_off, m.gtable_off.data(), m.gtable_off.size() * 8}};
    size_t off = 0;
    for (auto &it : items)
    {
        if (it.bytes == 0)
            continue;
        TRY(it.dst->ensure(it.bytes));
        const void *src = it.src;
        if (stage)
        {
            memcpy(stage + off, it.src, it.bytes);
            src = stage + off;
            off += (it.bytes + 15) & ~size_t(15);
        }
        CU(cudaMemcpyAsync(it.dst->p, src, it.bytes, cudaMemcpyHostToDevice, st));
    }
    return TXR_OK;
}

// The partitioned root level pays when the root's segment does not fit L2 anyway and the 16-bit counters and 32-bit
// element indices are wide enough for the batch.  Returns log2(partitions), 0 = use the item-per-warp kernel.
constexpr uint32_t kPartCtlWords = 4 + 1025 + 1024;
static uint32_t root_partition_bits(const txr_ctx *c, const BatchMeta &m)
{
    if (!c->root_partition || !c->index.loaded || c->index.generic) // the partition key is the prototype's segment-0 slot
        return 0;
    const IxfDev &root = c->index.ixf[0];
    const uint64_t seg_bytes = (uint64_t)root.seg_len * root.tbins;
    if (root.tbins > kSmallRowBytes || m.max_cap > 65535 || m.total_cap >= (1ull << 32))
        return 0;
    if (c->root_partition == 2)
        return 3; // forced (tests): 8 partitions whatever the size
    if (seg_bytes < (48ull << 20))
        return 0;
    uint32_t bits = 1;
    while (bits < 10 && (seg_bytes >> bits) > (16ull << 20))
        ++bits;
    return bits;
}

// ---------------------------------------------------------------------------------------------------------
// launching one batch on a slot
// ---------------------------------------------------------------------------------------------------------
static int slot_reserve(txr_ctx *c, Slot &s, const BatchMeta &m)
{
    const uint32_t n = m.n_reads;
    TRY(s.hashes.ensure(std::max<uint64_t>(m.total_cap, 1) * 8));
    TRY(s.n_raw.ensure((size_t)n * 4));
    TRY(s.hash_count.ensure((size_t)n * 4));
    TRY(s.deferred.ensure((size_t)n * 4));
    TRY(s.counters.ensure(C_TOTAL * 4));
    TRY(s.h_counters.ensure(C_TOTAL * 4));
    TRY(s.h_hash_count.ensure((size_t)n * 4));
    if (c->per_read_thr)
    {
        TRY(s.thr.ensure((size_t)n * 8));
        TRY(s.h_thr.ensure((size_t)n * 8));
    }
    if (!m.ids_global.empty())
        TRY(s.gtable.ensure(m.gtable_off.back() * 8));
    if (s.queue_cap < 2 * n + 1024)
        s.queue_cap = 2 * n + 1024;
    if (s.hit_cap < 4 * n + 1024)
        s.hit_cap = 4 * n + 1024;
    const uint32_t levels = std::max<uint32_t>(c->index.depth, 1);
    TRY(s.queues.ensure((size_t)(levels + 1) * 2 * s.queue_cap * sizeof(uint2)));
    TRY(s.ixf_hist.ensure((c->index.ixf.size() + 1) * 4));
    if (root_partition_bits(c, m))
    {
        TRY(s.part_hash.ensure(std::max<uint64_t>(m.total_cap, 1) * 8));
        TRY(s.part_read.ensure(std::max<uint64_t>(m.total_cap, 1) * 4));
        TRY(s.part_counts.ensure((size_t)n * c->index.ixf[0].tbins * 2));
        TRY(s.part_ctl.ensure(kPartCtlWords * 4));
    }
    TRY(s.hit_read.ensure((size_t)s.hit_cap * 4));
    TRY(s.hit_ub.ensure((size_t)s.hit_cap * 4));
    TRY(s.hit_cnt.ensure((size_t)s.hit_cap * 4));
    TRY(s.h_hits.ensure((size_t)s.hit_cap * 12));
    return TXR_OK;
}

static int launch_hash_stage(txr_ctx *c, Slot &s, const BatchMeta &m, const BatchDev &d, const uint64_t *d_words,
                             bool dedup, cudaStream_t cs)
{
    uint32_t *cnt = s.counters.as<uint32_t>();
    HashArgs h{};
    h.words = d_words;
    h.word_off = d.word_off.as<uint64_t>();
    h.len = d.len.as<uint32_t>();
    h.out_off = d.out_off.as<uint64_t>();
    h.out = s.hashes.as<uint64_t>();
    h.n_out = c->params.use_syncmer || c->params.scaling > 1 ? s.n_raw.as<uint32_t>() : s.hash_count.as<uint32_t>();
    h.n_reads = m.n_reads;
    h.work_counter = cnt + C_HASH_WORK;
    h.overflow = cnt + C_HASH_OVERFLOW;
    h.kmer_seed = c->kmer_seed;
    h.k = c->params.kmer_size;
    h.s = c->params.syncmer_size;
    h.t = c->params.t_syncmer;
    h.window = (int)c->params.window_size - (int)c->params.kmer_size + 1;
    h.smf = c->smf_hash;
    h.ctas_per_sm = c->shape_hash;
    h.min_blocks = c->hash_regs;
    h.probe_flag = c->adaptive_small_hash ? c->probe_flag.as<uint32_t>() : nullptr;
    h.small_grid = (uint32_t)(c->adaptive_small_hash * c->sm_count);
    // the distinct set of a read (syncmer.cpp:145) is built while hashing whenever the templated kernel runs: no raw list
    // round trip through HBM for the reads of the `ids_small` class
    const bool fused = dedup && c->fuse_dedup && c->params.use_syncmer &&
                       syncmer_has_fast_kernel(c->params.kmer_size, c->params.syncmer_size, c->params.t_syncmer);
    h.fuse_dedup = fused;
    h.fuse_max_keys = std::min<uint32_t>(c->fuse_max_keys, kWarpMaxKeys);
    h.hash_count = s.hash_count.as<uint32_t>();
    h.deferred = s.deferred.as<uint32_t>();
    h.n_deferred = cnt + C_DEDUP_DEFERRED;
    h.scaling = c->params.scaling;
    h.scaling_limit = double(UINT64_MAX) / double(c->params.scaling ? c->params.scaling : 1);
    s.fused = fused;
    if (c->params.use_syncmer)
        CU(launch_syncmer(h, c->sm_count, cs));
    else if (h.window > 1)
        CU(launch_minimiser(h, c->sm_count, cs));
    else
        CU(launch_kmer(h, c->sm_count, cs));
    c->timing.hash_launches += 1;
    CU(cudaEventRecord(s.ev[2], cs));

    if (!dedup)
    {
        CU(cudaEventRecord(s.ev[3], cs));
        return TXR_OK;
    }
    DedupArgs dd{};
    dd.out_off = d.out_off.as<uint64_t>();
    dd.hashes = s.hashes.as<uint64_t>();
    dd.n_raw = s.n_raw.as<uint32_t>();
    dd.hash_count = s.hash_count.as<uint32_t>();
    dd.scaling = c->params.scaling;
    dd.scaling_limit = double(UINT64_MAX) / double(c->params.scaling ? c->params.scaling : 1);
    dd.smf = c->smf_hash;
    dd.ctas_per_sm = c->shape_dedup;
    dd.probe_flag = c->adaptive_small_dedup ? c->probe_flag.as<uint32_t>() : nullptr;
    dd.small_grid = (uint32_t)(c->adaptive_small_dedup * c->sm_count);
    if (c->params.use_syncmer)
    {
        if (!m.ids_small.empty())
        {
            if (!fused)
            {
                dd.read_ids = m.ids_small.size() == m.n_reads ? nullptr : d.ids_small.as<uint32_t>();
                dd.n_ids = (uint32_t)m.ids_small.size();
                CU(launch_dedup_warp(dd, c->sm_count, cnt + C_DEDUP_WORK, s.deferred.as<uint32_t>(), cnt + C_DEDUP_DEFERRED, cs));
                c->timing.dedup_launches += 1;
            }
            dd.read_ids = s.deferred.as<uint32_t>(); // reads with too many distinct / raw hashes for the warp table
            CU(launch_dedup_deferred(dd, c->sm_count, cnt + C_DEDUP_DEFERRED, cs));
            c->timing.dedup_launches += 1;
        }
        if (!m.ids_medium.empty())
        {
            dd.read_ids = d.ids_medium.as<uint32_t>();
            dd.n_ids = (uint32_t)m.ids_medium.size();
            CU(launch_dedup_medium(dd, cs));
            c->timing.dedup_launches += 1;
        }
        if (!m.ids_global.empty())
        {
            CU(cudaMemsetAsync(s.gtable.p, 0xff, m.gtable_off.back() * 8, cs));
            dd.read_ids = d.ids_global.as<uint32_t>();
            dd.n_ids = (uint32_t)m.ids_global.size();
            dd.gtable = s.gtable.as<uint64_t>();
            dd.gtable_off = d.gtable_off.as<uint64_t>();
            CU(launch_dedup_global(dd, cs));
            c->timing.dedup_launches += 1;
        }
    }
    else if (c->params.scaling > 1)
    {
        dd.read_ids = nullptr;
        dd.n_ids = m.n_reads;
        CU(launch_filter(dd, cs));
        c->timing.dedup_launches += 1;
    }
    CU(cudaEventRecord(s.ev[3], cs));
    return TXR_OK;
}

static int launch_query_stage(txr_ctx *c, Slot &s, const BatchMeta &m, const BatchDev &d, cudaStream_t cs)
{
    const DeviceIndex &ix = c->index;
    uint32_t *cnt = s.counters.as<uint32_t>();
    QueryArgs q{};
    q.ixf = ix.d_ixf.as<IxfDev>();
    q.bin_ub = ix.d_bin_ub.as<int32_t>();
    q.bin_child = ix.d_bin_child.as<int32_t>();
    q.bin_run_begin = ix.d_bin_run_begin.as<uint32_t>();
    q.bin_kind = ix.d_bin_kind.as<uint8_t>();
    q.hashes = s.hashes.as<uint64_t>();
    q.hash_off = d.out_off.as<uint64_t>();
    q.hash_count = s.hash_count.as<uint32_t>();
    q.thr_lut = c->d_lut.as<uint64_t>();
    q.lut_len = (uint32_t)std::min<uint64_t>(c->d_lut_len, 0xffffffffu);
    q.thr_read = c->per_read_thr ? s.thr.as<uint64_t>() : nullptr;
    q.hit_read = s.hit_read.as<uint32_t>();
    q.hit_ub = s.hit_ub.as<int32_t>();
    q.hit_cnt = s.hit_cnt.as<uint32_t>();
    q.n_hits = cnt + C_NHITS;
    q.hit_cap = s.hit_cap;
    q.next_cap = s.queue_cap;
    q.stat_bytes = reinterpret_cast<unsigned long long *>(cnt + C_STATS);
    q.stat_items = reinterpret_cast<unsigned long long *>(cnt + C_STATS + 2);
    q.stat_skipped = reinterpret_cast<unsigned long long *>(cnt + C_STATS + 4);
    q.early_exit = c->early_exit;
    q.unroll = (uint32_t)c->shape_unroll;
    q.regs32 = c->query_regs32;
    q.l2_sector64 = c->l2_sector64 == 1;
    q.generic = ix.generic;
    q.scheme = ix.scheme;
    q.smf = c->smf_query;
    q.ctas_per_sm = c->shape_query;
    uint2 *queues = s.queues.as<uint2>();
    const uint32_t levels = std::min<uint32_t>(ix.depth, C_MAX_LEVELS);
    for (uint32_t lv = 0; lv < levels; ++lv)
    {
        uint32_t *lc = cnt + C_LEVEL0 + C_PER_LEVEL * lv;
        uint32_t *nc = cnt + C_LEVEL0 + C_PER_LEVEL * (lv + 1); // next level's counters
        const bool last = lv + 1 == levels;
        // outputs of this level (a leaf level never descends; give it valid but unused targets)
        const uint32_t nl = last ? lv : lv + 1;
        q.next_small = queues + (size_t)(2 * nl) * s.queue_cap;
        q.next_large = queues + (size_t)(2 * nl + 1) * s.queue_cap;
        q.next_small_n = last ? lc + 0 : nc + 0;
        q.next_large_n = last ? lc + 1 : nc + 1;
        if (lv == 0)
        {
            q.items = nullptr;
            q.n_items_ptr = nullptr;
            q.n_items_direct = m.n_reads;
            q.items_cap = m.n_reads;
            q.cursor = lc + 2;
            if (const uint32_t bits = root_partition_bits(c, m))
            {
                RootPartArgs rp{};
                rp.root = ix.ixf[0];
                rp.hashes = q.hashes;
                rp.hash_off = q.hash_off;
                rp.hash_count = q.hash_count;
                rp.n_reads = m.n_reads;
                rp.log2_parts = bits;
                uint32_t *ctl = s.part_ctl.as<uint32_t>();
                rp.work = ctl;
                rp.hist = ctl + 4;
                rp.cursor = ctl + 4 + 1025;
                rp.part_hash = s.part_hash.as<uint64_t>();
                rp.part_read = s.part_read.as<uint32_t>();
                rp.counts = s.part_counts.as<uint32_t>();
                CU(cudaMemsetAsync(ctl, 0, kPartCtlWords * 4, cs));
                CU(cudaMemsetAsync(rp.counts, 0, (size_t)m.n_reads * ix.ixf[0].tbins * 2, cs));
                CU(launch_root_partitioned(q, rp, c->sm_count, cs));
                c->timing.query_launches += 5;
                c->timing.probe_launches += 1;
            }
            else if (ix.ixf[0].tbins <= kSmallRowBytes)
            {
                CU(launch_query_small(q, c->sm_count, cs));
                c->timing.query_launches += 1;
                c->timing.probe_launches += 1;
            }
            else
            {
                CU(launch_query_large(q, c->sm_count, ix.max_tbins, cs));
                c->timing.query_launches += 1;
                c->timing.probe_launches += 1;
            }
        }
        else
        {
            q.items_cap = s.queue_cap;
            q.n_items_direct = 0;
            q.ctas_per_sm = c->shape_level;
            // group the level's items by IXF first (L2 reuse of the child IXFs, see query_kernels.cu)
            uint2 *sorted = queues + (size_t)(2 * levels) * s.queue_cap;
            const uint2 *raw = queues + (size_t)(2 * lv) * s.queue_cap;
            if (c->sort_items)
                CU(launch_sort_items(raw, lc + 0, s.queue_cap, s.ixf_hist.as<uint32_t>(), (uint32_t)ix.ixf.size(), sorted, c->sm_count, cs));
            q.items = c->sort_items ? sorted : raw;
            q.l2_hints = c->sort_items && c->l2_hints;
            q.l2_sector64 = c->l2_sector64 != 0;
            q.n_items_ptr = lc + 0;
            q.cursor = lc + 2;
            CU(launch_query_small(q, c->sm_count, cs));
            c->timing.query_launches += c->sort_items ? 4 : 1;
            c->timing.probe_launches += 1;
            if (ix.any_large)
            {
                sorted += s.queue_cap;
                raw = queues + (size_t)(2 * lv + 1) * s.queue_cap;
                if (c->sort_items)
                    CU(launch_sort_items(raw, lc + 1, s.queue_cap, s.ixf_hist.as<uint32_t>(), (uint32_t)ix.ixf.size(), sorted, c->sm_count, cs));
                q.items = c->sort_items ? sorted : raw;
                q.n_items_ptr = lc + 1;
                q.cursor = lc + 3;
                CU(launch_query_large(q, c->sm_count, ix.max_tbins, cs));
                c->timing.query_launches += c->sort_items ? 4 : 1;
                c->timing.probe_launches += 1;
            }
        }
    }
    CU(cudaEventRecord(s.ev[4], cs));
    return TXR_OK;
}

// hixf::threshold::threshold::get for every read of the batch (taxor_search.cpp:263): the scaling factor
// hash_count / (L - k + 1) of the FracMinHash model makes the threshold a function of the read, not of hash_count
// alone, and its double/libm arithmetic stays on the host -- so this stage waits for the hash counts, evaluates
// the model per read and ships the integers back.  Only minimiser indexes (window_size > k) pay this round trip.
static int per_read_thresholds(txr_ctx *c, Slot &s, const BatchMeta &m, cudaStream_t cs)
{
    const uint32_t n = m.n_reads;
    CU(cudaMemcpyAsync(s.h_hash_count.p, s.hash_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    const uint32_t *hc = s.h_hash_count.as<uint32_t>();
    uint64_t *thr = s.h_thr.as<uint64_t>();
    const double k = (double)c->params.kmer_size;
#pragma omp parallel for schedule(static) if (n > 4096)
    for (long r = 0; r < (long)n; ++r)
        thr[r] = c->thresholder.get(hc[r], (double)hc[r] / ((double)m.len[r] - k + 1.0));
    CU(cudaMemcpyAsync(s.thr.p, thr, (size_t)n * 8, cudaMemcpyHostToDevice, cs));
    return TXR_OK;
}

// All kernels of one batch (after its reads are on the device) + the D2H of its counters.
// Default: the kernels of all slots share ONE compute stream in batch order -- copies overlap compute, kernels never
// compete with each other.  Overlap mode (TXR_OVERLAP=1, several slots): hash + dedup of batch i+1 run on a second
// stream beside the probe kernels of batch i, either on the same SMs with small grids or, with TXR_SM_SPLIT=mod:n,
// on disjoint SMs (SmFilter).
// Is the hash stage of the next batch allowed beside the probe kernels of this one?  TXR_OVERLAP: 0 never, 1 always (manual
// shapes), 2 = auto: only where it was measured to win -- syncmer indexes hashed by a templated kernel, rows of at most 512
// bytes (the wide-row kernel needs the whole register file), the default probe arithmetic, thresholds from the LUT.
static bool overlap_applies(const txr_ctx *c)
{
    if (c->n_slots <= 1 || c->overlap_mode == 0)
        return false;
    if (c->overlap_mode == 1)
        return true;
    return c->params.use_syncmer && !c->per_read_thr && !c->index.generic && !c->index.any_large &&
           syncmer_has_fast_kernel(c->params.kmer_size, c->params.syncmer_size, c->params.t_syncmer);
}

// `first` / `last`: position of the batch inside its search call.  With the overlap on, the hash + dedup kernels of a batch
// run with small grids beside the probes of the batch before it -- except for the first batch, which has nothing beside it
// and takes the whole GPU; the probes of the last batch have no hash kernel beside them and run in their stand-alone shape.
static int enqueue_kernels(txr_ctx *c, Slot &s, const BatchMeta &m, const BatchDev &bd, const uint64_t *d_words, bool run_query,
                           bool first = true, bool last = true)
{
    const bool overlap = overlap_applies(c);
    const bool split = overlap && c->sm_mod > 1 && c->sm_hash > 0 && c->sm_hash < c->sm_mod;
    const bool beside_hash = overlap && !split && !last;  // probes of this batch share the SMs with the next batch's hash stage
    const bool beside_probe = overlap && !split && !first; // hash stage of this batch shares the SMs with the previous batch's probes
    // probe kernel beside a hash kernel: one step in flight per warp (40 registers), 7 CTAs per SM launched (6 fit beside two
    // 124-register hash CTAs, the seventh whenever a hash CTA has left; profiles/r2_f_u1_sweep.txt); alone: two steps in
    // flight, 8 CTAs of 64 registers
    c->shape_query = beside_hash ? (c->query_ctas ? c->query_ctas : 7) : (!overlap && c->query_ctas ? c->query_ctas : 8);
    c->shape_unroll = beside_hash ? (c->query_unroll ? c->query_unroll : 1) : (!overlap ? c->query_unroll : 0);
    c->shape_level = c->level_ctas ? c->level_ctas : c->shape_query;
    // Adaptive grids (TXR_ADAPTIVE, default on): the hash-stage kernels of a batch that may run beside probes are launched for the
    // whole GPU, and their CTAs beyond the small share leave at once if the probes are running at that moment.  When the probes of
    // the batch before are long done -- a host-fed search bound by its H2D copies, as on a box where eight GPUs pull reads out of
    // one host -- the hash stage then takes the idle GPU instead of crawling through it with a quarter of the warps.
    const bool adaptive = beside_probe && c->adaptive && !c->hash_ctas && !c->dedup_ctas;
    c->adaptive_small_hash = adaptive ? 2 : 0;
    c->shape_hash = adaptive ? 8 : beside_probe ? (c->hash_ctas ? c->hash_ctas : 2) : (!overlap && c->hash_ctas ? c->hash_ctas : 8);
    // dedup beside the probes: 3 CTAs per SM for full-size batches (fewest probe registers taken: best device-resident step),
    // 6 for the smaller batches of a host-fed call's ramp, whose dedup must finish within the shorter probes of the batch before
    // (profiles/r2_f_u1_sweep.txt, run F4: d3 84.2 / 94.8 ms resident / end to end, d6 87.0 / 93.4)
    const int dedup_beside = m.n_reads >= c->max_batch_reads ? 3 : 6;
    c->adaptive_small_dedup = adaptive ? dedup_beside : 0;
    c->shape_dedup = adaptive ? 6 : beside_probe ? (c->dedup_ctas ? c->dedup_ctas : dedup_beside) : (!overlap && c->dedup_ctas ? c->dedup_ctas : 6);
    s.first = first;
    s.last = last;
    c->smf_hash = split ? SmFilter{c->sm_mod, 0, c->sm_hash} : SmFilter{0, 0, 0};
    c->smf_query = split ? SmFilter{c->sm_mod, c->sm_hash, c->sm_mod} : SmFilter{0, 0, 0};
    cudaStream_t cs = c->compute, hs = overlap ? c->compute_hash : c->compute;
    CU(cudaStreamWaitEvent(hs, s.ev[1], 0));
    CU(cudaMemsetAsync(s.counters.p, 0, C_TOTAL * 4, hs));
    CU(cudaEventRecord(s.ev[6], hs));
    TRY(launch_hash_stage(c, s, m, bd, d_words, true, hs));
    if (run_query && c->per_read_thr)
        TRY(per_read_thresholds(c, s, m, hs));
    CU(cudaEventRecord(s.ev[9], hs));
    CU(cudaStreamWaitEvent(cs, s.ev[9], 0));
    CU(cudaEventRecord(s.ev[8], cs));
    if (run_query)
    {
        if (overlap)
            CU(cudaMemsetAsync(c->probe_flag.p, 1, 4, cs)); // "probes running": read by the adaptive grids of the next batch's hash stage
        TRY(launch_query_stage(c, s, m, bd, cs));
        if (overlap)
            CU(cudaMemsetAsync(c->probe_flag.p, 0, 4, cs));
    }
    else
        CU(cudaEventRecord(s.ev[4], cs));
    CU(cudaEventRecord(s.ev[7], cs));
    CU(cudaStreamWaitEvent(s.stream, s.ev[7], 0));
    CU(cudaMemcpyAsync(s.h_counters.p, s.counters.p, C_TOTAL * 4, cudaMemcpyDeviceToHost, s.stream));
    s.split_used = split;
    s.ran_query = run_query;
    return TXR_OK;
}

// With an SM partition a kernel only makes progress on the SMs its filter keeps.  The block scheduler has always put
// CTAs there, but nothing in the programming model promises it, so every work-stealing cursor is checked: a cursor
// below its item count means some kernel ran without a single kept CTA.
static bool split_left_work_undone(const txr_ctx *c, const Slot &s, const BatchMeta &m)
{
    const uint32_t *hc = s.h_counters.as<uint32_t>();
    if (hc[C_HASH_WORK] < m.n_reads)
        return true;
    if (c->params.use_syncmer && !s.fused && !m.ids_small.empty() && hc[C_DEDUP_WORK] < m.ids_small.size())
        return true;
    if (!s.ran_query)
        return false;
    const uint32_t levels = std::min<uint32_t>(c->index.depth, C_MAX_LEVELS);
    for (uint32_t lv = 0; lv < levels; ++lv)
    {
        const uint32_t *lc = hc + C_LEVEL0 + C_PER_LEVEL * lv;
        const uint32_t n_small = lv == 0 ? (c->index.ixf[0].tbins <= kSmallRowBytes ? m.n_reads : 0) : std::min(lc[0], s.queue_cap);
        const uint32_t n_large = lv == 0 ? (c->index.ixf[0].tbins <= kSmallRowBytes ? 0 : m.n_reads) : std::min(lc[1], s.queue_cap);
        if (lv == 0 ? lc[2] < n_small + n_large : (lc[2] < n_small || lc[3] < n_large))
            return true;
    }
    return false;
}

// enqueue everything for one batch; `h_words` != nullptr: copy the packed reads from the host first
static int submit_batch(txr_ctx *c, Slot &s, const BatchMeta &m, const BatchDev *resident_meta,
                        const uint64_t *d_words_resident, const uint64_t *h_words, bool run_query, bool first = true, bool last = true)
{
    TRY(slot_reserve(c, s, m));
    TRY(ensure_lut(c, m.max_cap + 1));
    CU(cudaEventRecord(s.ev[0], s.stream));
    const BatchDev *bd = resident_meta;
    const uint64_t *d_words = d_words_resident;
    if (h_words)
    {
        TRY(s.words.ensure(m.n_words * 8));
        TRY(s.h_meta.ensure(meta_bytes(m) + 7 * 16));
        CU(cudaMemcpyAsync(s.words.p, h_words + m.first_word, m.n_words * 8, cudaMemcpyHostToDevice, s.stream));
        TRY(upload_batch_meta(m, s.meta, s.stream, s.h_meta.as<uint8_t>()));
        bd = &s.meta;
        d_words = s.words.as<uint64_t>();
    }
    CU(cudaEventRecord(s.ev[1], s.stream));
    TRY(enqueue_kernels(c, s, m, *bd, d_words, run_query, first, last));
    s.bm = &m;
    s.bd = bd;
    s.d_words = d_words;
    s.busy = true;
    return TXR_OK;
}

// waits for a batch, re-runs the query stage with larger queues if one overflowed, optionally fetches the hits
// and appends the batch to the result store
static int collect_batch(txr_ctx *c, Slot &s, bool fetch)
{
    const BatchMeta &m = *s.bm;
    for (int attempt = 0;; ++attempt)
    {
        CU(cudaStreamSynchronize(s.stream));
        if (s.split_used && split_left_work_undone(c, s, m))
        {
            // never observed; keeps the result exact if the block scheduler ever behaves differently
            fprintf(stderr, "taxor_b200: SM partition left work undone, disabling it and re-running the batch\n");
            c->sm_mod = 0;
            CU(cudaDeviceSynchronize());
            CU(cudaEventRecord(s.ev[1], s.stream));
            TRY(enqueue_kernels(c, s, m, *s.bd, s.d_words, s.ran_query, s.first, s.last));
            continue;
        }
        const uint32_t *hc = s.h_counters.as<uint32_t>();
        if (hc[C_HASH_OVERFLOW])
            return set_error(TXR_ERR_OVERFLOW, "hash capacity bound violated (k=%d s=%d t=%d)", c->params.kmer_size,
                             c->params.syncmer_size, c->params.t_syncmer);
        bool overflow = hc[C_NHITS] > s.hit_cap;
        uint32_t need_q = 0;
        for (uint32_t lv = 0; lv < std::min<uint32_t>(c->index.depth, C_MAX_LEVELS); ++lv)
            need_q = std::max(need_q, std::max(hc[C_LEVEL0 + C_PER_LEVEL * lv], hc[C_LEVEL0 + C_PER_LEVEL * lv + 1]));
        overflow = overflow || need_q > s.queue_cap;
        if (!overflow)
            break;
        if (attempt >= 8)
            return set_error(TXR_ERR_OVERFLOW, "hit/queue buffers still too small after %d retries", attempt);
        // grow (the overflowing counters are lower bounds only: deeper levels were truncated), then re-run
        s.hit_cap = std::max<uint32_t>(s.hit_cap, hc[C_NHITS]) * 2;
        s.queue_cap = std::max(s.queue_cap, need_q) * 2;
        TRY(slot_reserve(c, s, m));
        CU(cudaMemsetAsync(s.counters.p, 0, C_TOTAL * 4, c->compute));
        TRY(launch_query_stage(c, s, m, *s.bd, c->compute));
        CU(cudaEventRecord(s.ev[7], c->compute));
        CU(cudaStreamWaitEvent(s.stream, s.ev[7], 0));
        CU(cudaMemcpyAsync(s.h_counters.p, s.counters.p, C_TOTAL * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    const uint32_t *hc = s.h_counters.as<uint32_t>();
    const uint32_t n_hits = hc[C_NHITS];
    uint64_t stats[3];
    memcpy(stats, hc + C_STATS, 24);
    c->timing.query_bytes += stats[0];
    c->timing.query_items += stats[1];
    c->timing.skipped_hashes += stats[2];

    float ms = 0;
    cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]);
    c->timing.h2d_ms += ms;
    cudaEventElapsedTime(&ms, s.ev[6], s.ev[2]);
    c->timing.hash_ms += ms;
    cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]);
    c->timing.dedup_ms += ms;
    cudaEventElapsedTime(&ms, s.ev[8], s.ev[4]);
    c->timing.query_ms += ms;
    if (c->trace && c->trace_ev)
    {
        // TXR_TRACE=1: when each stage of the batch started and ended, in ms since the call's fork (stderr; experiments only)
        float t[8] = {0};
        const int idx[8] = {0, 1, 6, 2, 3, 8, 4, 7};
        for (int i = 0; i < 8; ++i)
            cudaEventElapsedTime(&t[i], c->trace_ev, s.ev[idx[i]]);
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c->trace_t0).count();
        fprintf(stderr, "[txr trace] reads %u..+%u  h2d %.2f-%.2f  hash %.2f-%.2f  dedup -%.2f  query %.2f-%.2f  done %.2f  (collected at host %.2f)\n",
                (unsigned)m.first_read, m.n_reads, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], host_ms);
    }

    if (fetch)
    {
        const uint32_t n = m.n_reads;
        CU(cudaEventRecord(s.ev[0], s.stream));
        CU(cudaMemcpyAsync(s.h_hash_count.p, s.hash_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
        uint32_t *hh = s.h_hits.as<uint32_t>();
        if (n_hits)
        {
            CU(cudaMemcpyAsync(hh, s.hit_read.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaMemcpyAsync(hh + s.hit_cap, s.hit_ub.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaMemcpyAsync(hh + 2 * (size_t)s.hit_cap, s.hit_cnt.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, s.stream));
        }
        CU(cudaEventRecord(s.ev[5], s.stream));
        CU(cudaStreamSynchronize(s.stream));
        cudaEventElapsedTime(&ms, s.ev[0], s.ev[5]);
        c->timing.d2h_ms += ms;

        // ---- host post-pass: group by read, DFS pre-order, threshold, 0.8*max flags ----
        ResultStore &R = c->result;
        const uint32_t *h_count = s.h_hash_count.as<uint32_t>();
        const uint32_t *h_read = hh;
        const int32_t *h_ub = reinterpret_cast<const int32_t *>(hh + s.hit_cap);
        const uint32_t *h_cnt = hh + 2 * (size_t)s.hit_cap;
        std::vector<uint32_t> begin((size_t)n + 1, 0);
        for (uint32_t i = 0; i < n_hits; ++i)
            ++begin[h_read[i] + 1];
        for (uint32_t r = 0; r < n; ++r)
            begin[r + 1] += begin[r];
        std::vector<uint32_t> order(n_hits), fill(begin.begin(), begin.end() - 1);
        for (uint32_t i = 0; i < n_hits; ++i)
            order[fill[h_read[i]]++] = i;
        const std::vector<uint32_t> &rank = c->index.dfs_rank;
        const size_t base_hits = R.user_bin.size();
        R.user_bin.resize(base_hits + n_hits);
        R.count.resize(base_hits + n_hits);
        R.keep.resize(base_hits + n_hits);
        uint64_t n_hashes = 0, hash_bytes = 0;
        const size_t base_reads = R.hash_count.size();
        R.hash_count.resize(base_reads + n);
        R.threshold.resize(base_reads + n);
        R.hit_begin.resize(base_reads + n);
        // reads are independent and write disjoint ranges: the pass of the LAST batch is what nothing overlaps
#pragma omp parallel for schedule(static) reduction(+ : n_hashes, hash_bytes) if (n > 16384)
        for (long rr = 0; rr < (long)n; ++rr)
        {
            const uint32_t r = (uint32_t)rr;
            uint32_t *o = order.data() + begin[r];
            const uint32_t cnt = begin[r + 1] - begin[r];
            if (cnt > 1)
                std::sort(o, o + cnt, [&](uint32_t x, uint32_t y) { return rank[h_ub[x]] < rank[h_ub[y]]; });
            uint64_t max_count = 0;                                       // taxor_search.cpp:275-280
            for (uint32_t i = 0; i < cnt; ++i)
                max_count = std::max<uint64_t>(max_count, h_cnt[o[i]]);
            for (uint32_t i = 0; i < cnt; ++i)
            {
                const size_t at = base_hits + begin[r] + i;
                R.user_bin[at] = h_ub[o[i]];
                R.count[at] = h_cnt[o[i]];
                // taxor_search.cpp:285: dropped iff double(count) < double(max_count) * 0.8
                R.keep[at] = !(static_cast<double>(h_cnt[o[i]]) < static_cast<double>(max_count) * 0.8);
            }
            R.hash_count[base_reads + r] = h_count[r];
            if (c->per_read_thr)
                R.threshold[base_reads + r] = s.h_thr.as<uint64_t>()[r];
            else
                R.threshold[base_reads + r] = h_count[r] < c->lut.size() ? c->lut[h_count[r]] : c->thresholder.get(h_count[r], 1.0);
            R.hit_begin[base_reads + r] = base_hits + begin[r];
            n_hashes += h_count[r];
            hash_bytes += (m.len[r] + 3) / 4;
        }
        c->timing.n_hashes += n_hashes;
        c->timing.hash_bytes += hash_bytes + 8 * n_hashes;
    }
    s.busy = false;
    return TXR_OK;
}

static void finish_result(txr_ctx *c, txr_result *out)
{
    ResultStore &R = c->result;
    R.hit_begin.push_back(R.user_bin.size());
    if (!out)
        return;
    out->n_reads = R.hash_count.size();
    out->hash_count = R.hash_count.data();
    out->threshold = R.threshold.data();
    out->hit_begin = R.hit_begin.data();
    out->user_bin = R.user_bin.data();
    out->count = R.count.data();
    out->keep = R.keep.data();
}

// split reads [0, n) into batches bounded by reads and bases.  `ramp`: when the reads come from the host, the batches
// grow geometrically from 1/16 of the limits: nothing can overlap the first copy, and a batch's copy hides behind the
// kernels of the batch before it only if it is not much larger (PCIe moves a batch in about half the time the kernels
// need for it, hence the factor 1.8).
static void plan_batches(const txr_ctx *c, const uint32_t *len, uint64_t n, std::vector<std::pair<uint64_t, uint32_t>> &out,
                         bool ramp = false)
{
    out.clear();
    double scale = ramp ? 1.0 / 16.0 : 1.0;
    uint64_t i = 0;
    while (i < n)
    {
        const uint64_t max_reads = std::max<uint64_t>((uint64_t)((double)c->max_batch_reads * scale), 1);
        const uint64_t max_bases = std::max<uint64_t>((uint64_t)((double)c->max_batch_bases * scale), 1);
        uint64_t bases = 0, j = i;
        while (j < n && j - i < max_reads && (j == i || bases + len[j] <= max_bases))
            bases += len[j++];
        out.emplace_back(i, (uint32_t)(j - i));
        i = j;
        scale = std::min(1.0, scale * 1.8);
    }
}

// make the slot streams wait for work already queued on the caller's stream ...
static int fork_streams(txr_ctx *c)
{
    if (!c->fork_ev)
    {
        CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(c->fork_ev, c->primary));
    if (c->trace)
    {
        if (!c->trace_ev)
            CU(cudaEventCreate(&c->trace_ev));
        CU(cudaEventRecord(c->trace_ev, c->primary));
        c->trace_t0 = std::chrono::steady_clock::now();
    }
    CU(cudaStreamWaitEvent(c->compute, c->fork_ev, 0));
    CU(cudaStreamWaitEvent(c->compute_hash, c->fork_ev, 0));
    for (auto &s : c->slots)
        CU(cudaStreamWaitEvent(s->stream, c->fork_ev, 0));
    return TXR_OK;
}
// ... and the caller's stream wait for everything the slots did (so events on it bracket the search)
static int join_streams(txr_ctx *c)
{
    CU(cudaEventRecord(c->join_ev, c->compute));
    CU(cudaStreamWaitEvent(c->primary, c->join_ev, 0));
    CU(cudaEventRecord(c->join_ev, c->compute_hash));
    CU(cudaStreamWaitEvent(c->primary, c->join_ev, 0));
    for (auto &s : c->slots)
    {
        CU(cudaEventRecord(c->join_ev, s->stream));
        CU(cudaStreamWaitEvent(c->primary, c->join_ev, 0));
    }
    return TXR_OK;
}

static int check_ready(txr_ctx *c)
{
    if (!c)
        return set_error(TXR_ERR_ARG, "null context");
    if (!c->index.loaded)
        return set_error(TXR_ERR_STATE, "no index uploaded");
    if (!c->have_params)
        return set_error(TXR_ERR_STATE, "txr_params_set not called");
    CU(cudaSetDevice(c->device));
    return ensure_slots(c);
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char *txr_last_error(void) { return g_last_error.c_str(); }
const char *txr_version(void) { return "taxor_b200 0.1 (sm_100a)"; }

int txr_ctx_create(int device, txr_ctx **out)
{
    if (!out)
        return set_error(TXR_ERR_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return set_error(TXR_ERR_CUDA, "CUDA device %d not available (%d devices); this library has no CPU fallback", device, n);
    CU(cudaSetDevice(device));
    auto c = std::make_unique<txr_ctx>();
    c->device = device;
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (const char *e = getenv("TXR_SORT_ITEMS"))
        c->sort_items = atoi(e) != 0;
    if (const char *e = getenv("TXR_OVERLAP"))
        c->overlap_mode = atoi(e);
    if (const char *e = getenv("TXR_SM_SPLIT")) // "mod:n": of every mod SM ids, n run hash + dedup
    {
        unsigned mod = 0, n = 0;
        if (sscanf(e, "%u:%u", &mod, &n) == 2)
        {
            c->sm_mod = mod;
            c->sm_hash = n;
        }
    }
    if (const char *e = getenv("TXR_QUERY_CTAS_PER_SM"))
        c->query_ctas = atoi(e);
    if (const char *e = getenv("TXR_LEVEL_CTAS_PER_SM"))
        c->level_ctas = atoi(e);
    if (const char *e = getenv("TXR_HASH_CTAS_PER_SM"))
        c->hash_ctas = atoi(e);
    if (const char *e = getenv("TXR_DEDUP_CTAS_PER_SM"))
        c->dedup_ctas = atoi(e);
    if (const char *e = getenv("TXR_EARLY_EXIT"))
        c->early_exit = atoi(e) != 0;
    if (const char *e = getenv("TXR_L2_HINTS"))
        c->l2_hints = atoi(e) != 0;
    if (const char *e = getenv("TXR_L2_SECTOR64"))
        c->l2_sector64 = (uint32_t)atoi(e);
    if (const char *e = getenv("TXR_QUERY_REGS"))
        c->query_regs32 = atoi(e) == 32;
    if (const char *e = getenv("TXR_TRACE"))
        c->trace = atoi(e) != 0;
    if (const char *e = getenv("TXR_ADAPTIVE"))
        c->adaptive = atoi(e) != 0;
    if (const char *e = getenv("TXR_RAMP_CUM"))
        c->ramp_cum = atof(e);
    if (const char *e = getenv("TXR_RAMP"))
        c->ramp = std::max(1.05, atof(e));
    if (const char *e = getenv("TXR_HASH_REGS"))
        c->hash_regs = atoi(e);
    if (const char *e = getenv("TXR_QUERY_UNROLL"))
        c->query_unroll = (uint32_t)atoi(e);
    if (const char *e = getenv("TXR_FUSE_DEDUP"))
        c->fuse_dedup = atoi(e) != 0;
    if (const char *e = getenv("TXR_FUSE_MAX_KEYS"))
        c->fuse_max_keys = (uint32_t)std::max(32, atoi(e));
    if (const char *e = getenv("TXR_ROOT_PARTITION"))
        c->root_partition = atoi(e);
    *out = c.release();
    return TXR_OK;
}

void txr_ctx_destroy(txr_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &s : c->slots)
    {
        for (DevBuf *b : {&s->words, &s->meta.word_off, &s->meta.len, &s->meta.out_off, &s->meta.ids_small, &s->meta.ids_medium,
                          &s->meta.ids_global, &s->meta.gtable_off, &s->hashes, &s->n_raw, &s->hash_count, &s->gtable, &s->deferred, &s->queues,
                          &s->hit_read, &s->hit_ub, &s->hit_cnt, &s->counters, &s->thr, &s->ixf_hist, &s->part_hash,
                          &s->part_read, &s->part_counts, &s->part_ctl})
            b->release();
        for (PinBuf *b : {&s->h_meta, &s->h_counters, &s->h_hash_count, &s->h_hits, &s->h_thr})
            b->release();
        for (auto &e : s->ev)
            cudaEventDestroy(e);
        cudaStreamDestroy(s->stream);
    }
    if (c->compute)
    {
        cudaStreamDestroy(c->compute);
        cudaStreamDestroy(c->compute_hash);
    }
    if (c->fork_ev)
    {
        cudaEventDestroy(c->fork_ev);
        cudaEventDestroy(c->join_ev);
    }
    c->index.release();
    c->d_lut.release();
    c->scratch_a.release();
    c->scratch_b.release();
    c->probe_flag.release();
    for (auto &b : c->binset)
        b.release();
    delete c;
}

int txr_ctx_set_stream(txr_ctx *c, void *stream)
{
    if (!c)
        return set_error(TXR_ERR_ARG, "null context");
    c->primary = static_cast<cudaStream_t>(stream);
    return TXR_OK;
}

int txr_ctx_configure(txr_ctx *c, uint64_t max_batch_reads, uint64_t max_batch_bases, int n_slots)
{
    if (!c || max_batch_reads == 0 || max_batch_bases == 0 || n_slots < 1 || n_slots > 8)
        return set_error(TXR_ERR_ARG, "bad configuration");
    if (max_batch_reads > (1u << 30))
        return set_error(TXR_ERR_ARG, "max_batch_reads too large");
    c->max_batch_reads = max_batch_reads;
    c->max_batch_bases = max_batch_bases;
    c->n_slots = n_slots;
    return TXR_OK;
}

// Fingerprint rows -> HBM arena.  The source is usually an mmap of the .hixf (pageable, maybe not even faulted in): a
// plain cudaMemcpy from it runs at ~8 GB/s, one bounce buffer at a time.  Here the rows are cut into pieces of whole rows;
// worker threads copy (or, when rows have to be padded to 64 bytes, re-stride) pieces into their own PINNED staging
// buffers and queue an async copy each on their own stream, so the page faults and memcpys of one piece overlap the DMA
// of the others.  TXR_UPLOAD_THREADS overrides the worker count (1 = the simple serial copy, for A/B measurements).
static inline uint64_t view_rows(const txr_ixf_view &x) { return x.rows ? x.rows : 3 * x.seg_len; }

static int upload_fingerprints(txr_ctx *c, const txr_hixf_view *v, DeviceIndex &ix)
{
    struct Piece
    {
        uint64_t ixf, row0, rows;
    };
    constexpr uint64_t kStage = 4ull << 20; // pinned staging is allocated per upload: 2 x 4 MB per worker keeps that under ~30 ms
    const bool bin_major = v->scheme && v->scheme->layout == TXR_IXF_LAYOUT_BIN_MAJOR;
    std::vector<Piece> pieces;
    uint64_t total = 0;
    for (uint64_t i = 0; i < v->n_ixf; ++i)
    {
        const uint64_t rows = view_rows(v->ixf[i]), stride = ix.ixf[i].tbins;
        const uint64_t per = std::max<uint64_t>(1, kStage / stride);
        for (uint64_t r0 = 0; r0 < rows; r0 += per)
            pieces.push_back(Piece{i, r0, std::min(per, rows - r0)});
        total += rows * stride;
    }
    int n_workers = (int)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 12);
    if (const char *e = getenv("TXR_UPLOAD_THREADS"))
        n_workers = std::max(1, atoi(e));
    n_workers = (int)std::min<uint64_t>((uint64_t)n_workers, std::max<uint64_t>(1, total / kStage));
    if (bin_major)
        n_workers = std::max(n_workers, 1);
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    std::string first_error;
    std::mutex err_m;
    auto work = [&](int) {
        if (cudaSetDevice(c->device) != cudaSuccess)
        {
            failed = 1;
            return;
        }
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        uint8_t *buf[2] = {nullptr, nullptr};
        auto fail = [&](const char *what, cudaError_t e) {
            std::lock_guard<std::mutex> l(err_m);
            if (first_error.empty())
                first_error = std::string(what) + ": " + cudaGetErrorString(e);
            failed = 1;
        };
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b)
        {
            e = cudaHostAlloc((void **)&buf[b], kStage, cudaHostAllocDefault);
            if (e == cudaSuccess)
                e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
        }
        if (e != cudaSuccess)
            fail("upload staging", e);
        int cur = 0;
        bool used[2] = {false, false};
        while (!failed)
        {
            const size_t p = next.fetch_add(1);
            if (p >= pieces.size())
                break;
            const Piece &pc = pieces[p];
            const txr_ixf_view &x = v->ixf[pc.ixf];
            const uint64_t stride = ix.ixf[pc.ixf].tbins;
            if (used[cur] && (e = cudaEventSynchronize(ev[cur])) != cudaSuccess)
            {
                fail("upload wait", e);
                break;
            }
            const uint8_t *src = x.fp + pc.row0 * x.tbins;
#ifdef MADV_POPULATE_READ
            if (!bin_major)
            {
                // the source is typically a fresh mmap of the .hixf: map the piece's pages in one go instead of taking a
                // minor fault every 4 KB inside the memcpy (a hint -- errors, e.g. from older kernels, are ignored)
                const uintptr_t lo = reinterpret_cast<uintptr_t>(src) & ~uintptr_t(4095);
                const uintptr_t hi = reinterpret_cast<uintptr_t>(src) + pc.rows * x.tbins;
                madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_POPULATE_READ);
            }
#endif
            if (bin_major)
            {
                // fp[bin * rows + slot] -> rows of `stride` bytes (HBM is always slot-major: one probe = three rows)
                const uint64_t all_rows = view_rows(x);
                memset(buf[cur], 0, pc.rows * stride);
                for (uint64_t b = 0; b < x.bins; ++b)
                {
                    const uint8_t *col = x.fp + b * all_rows + pc.row0;
                    for (uint64_t r = 0; r < pc.rows; ++r)
                        buf[cur][r * stride + b] = col[r];
                }
            }
            else if (stride == x.tbins)
                memcpy(buf[cur], src, pc.rows * stride);
            else
                for (uint64_t r = 0; r < pc.rows; ++r)
                {
                    memcpy(buf[cur] + r * stride, src + r * x.tbins, x.tbins);
                    memset(buf[cur] + r * stride + x.tbins, 0, stride - x.tbins);
                }
            uint8_t *dst = const_cast<uint8_t *>(ix.ixf[pc.ixf].fp) + pc.row0 * stride;
            e = cudaMemcpyAsync(dst, buf[cur], pc.rows * stride, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess)
                e = cudaEventRecord(ev[cur], st);
            if (e != cudaSuccess)
            {
                fail("upload copy", e);
                break;
            }
            used[cur] = true;
            cur ^= 1;
        }
        if (st)
        {
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess)
                fail("upload sync", e);
            cudaStreamDestroy(st);
        }
        for (int b = 0; b < 2; ++b)
        {
            if (ev[b])
                cudaEventDestroy(ev[b]);
            if (buf[b])
                cudaFreeHost(buf[b]);
        }
    };
    if (n_workers <= 1 && !bin_major)
    {
        // the plain path: pageable copies, row-strided when padding is needed
        for (uint64_t i = 0; i < v->n_ixf; ++i)
        {
            const txr_ixf_view &x = v->ixf[i];
            uint8_t *dst = const_cast<uint8_t *>(ix.ixf[i].fp);
            const uint64_t rows = view_rows(x), dev_tbins = ix.ixf[i].tbins;
            if (dev_tbins == x.tbins)
                CU(cudaMemcpy(dst, x.fp, rows * x.tbins, cudaMemcpyHostToDevice));
            else
            {
                CU(cudaMemset(dst, 0, rows * dev_tbins));
                CU(cudaMemcpy2D(dst, dev_tbins, x.fp, x.tbins, x.tbins, rows, cudaMemcpyHostToDevice));
            }
        }
        return TXR_OK;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_workers; ++t)
        th.emplace_back(work, t);
    for (auto &t : th)
        t.join();
    if (failed)
        return set_error(TXR_ERR_CUDA, "index upload failed: %s", first_error.c_str());
    return TXR_OK;
}

// Allocates the per-slot device and pinned buffers a search of up to n_reads reads / n_bases bases per call would allocate
// on its first call, so that the first call does not pay for it (a driver with several GPUs calls this from each GPU's own
// thread while the reads are still being parsed: the allocations of eight contexts otherwise land in the timed search phase).
int txr_ctx_reserve(txr_ctx *c, uint64_t n_reads, uint64_t n_bases)
{
    if (!c || !c->have_params || n_reads == 0)
        return set_error(TXR_ERR_STATE, "txr_params_set not called");
    CU(cudaSetDevice(c->device));
    TRY(ensure_slots(c));
    const uint64_t n = std::min<uint64_t>(n_reads, c->max_batch_reads);
    const uint64_t bases = std::min<uint64_t>(std::max<uint64_t>(n_bases, n), c->max_batch_bases);
    BatchMeta m;
    m.n_reads = (uint32_t)n;
    m.n_bases = bases;
    const uint64_t mean_len = (bases + n - 1) / n;
    const uint64_t cap = hash_capacity(c->params, mean_len + 64);
    m.total_cap = cap * n;
    m.max_cap = cap;
    m.n_words = bases / 32 + 2 * n + 64;
    m.gtable_off.assign(1, 0);
    TRY(ensure_lut(c, cap + 1));
    for (auto &sp : c->slots)
    {
        Slot &s = *sp;
        TRY(slot_reserve(c, s, m));
        TRY(s.words.ensure(m.n_words * 8));
        TRY(s.h_meta.ensure((size_t)n * 32 + 4096));
        TRY(s.meta.word_off.ensure((size_t)n * 8));
        TRY(s.meta.len.ensure((size_t)n * 4));
        TRY(s.meta.out_off.ensure((size_t)(n + 1) * 8));
    }
    return TXR_OK;
}

int txr_index_upload(txr_ctx *c, const txr_hixf_view *v)
{
    if (!c || !v || v->n_ixf == 0)
        return set_error(TXR_ERR_ARG, "null or empty index");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    DeviceIndex &ix = c->index;
    ix.release();
    const uint64_t n = v->n_ixf;
    IxfScheme sch = ixf_default_scheme();
    if (v->scheme)
    {
        sch = IxfScheme{v->scheme->slots, v->scheme->mix, v->scheme->fingerprint, v->scheme->rot1, v->scheme->rot2};
        if (sch.rot1 == 0 && sch.rot2 == 0)
        {
            sch.rot1 = 21;
            sch.rot2 = 42;
        }
        if (!ixf_scheme_valid(sch) || v->scheme->layout > TXR_IXF_LAYOUT_BIN_MAJOR)
            return set_error(TXR_ERR_ARG, "unknown interleaved-XOR-filter scheme (slots %u mix %u fingerprint %u rot %u/%u layout %u)",
                             v->scheme->slots, v->scheme->mix, v->scheme->fingerprint, v->scheme->rot1, v->scheme->rot2, v->scheme->layout);
    }
    ix.scheme = sch;
    ix.generic = !ixf_scheme_is_default(sch);
    // validate + size the arena (rows padded to a multiple of 64 bytes, each IXF 256-byte aligned)
    std::vector<uint64_t> arena_off(n);
    uint64_t arena_bytes = 0, total_bins = v->bin_off[n];
    ix.ixf.resize(n);
    ix.max_tbins = 0;
    ix.any_large = false;
    ix.fp_bytes = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        const txr_ixf_view &x = v->ixf[i];
        uint64_t count_len = 0;
        if (x.bins == 0 || x.tbins < x.bins || !x.fp || !ixf_geometry_ok(sch, x.seg_len, view_rows(x), count_len))
            return set_error(TXR_ERR_ARG, "IXF %llu: inconsistent geometry (bins %llu, row width %llu, %llu slots per segment, %llu rows)",
                             (unsigned long long)i, (unsigned long long)x.bins, (unsigned long long)x.tbins,
                             (unsigned long long)x.seg_len, (unsigned long long)view_rows(x));
        if (v->bin_off[i + 1] - v->bin_off[i] != x.bins)
            return set_error(TXR_ERR_ARG, "IXF %llu: bin metadata size != bins", (unsigned long long)i);
        const uint64_t dev_tbins = (x.tbins + 63) / 64 * 64;
        if (dev_tbins > query_large_max_tbins()) // the counters of one (read, IXF) must fit a CTA's shared memory
            return set_error(TXR_ERR_UNSUPPORTED, "IXF %llu: %llu technical bins not supported (limit %u)", (unsigned long long)i,
                             (unsigned long long)x.tbins, query_large_max_tbins());
        arena_off[i] = arena_bytes;
        arena_bytes += (view_rows(x) * dev_tbins + 255) / 256 * 256;
        ix.ixf[i] = IxfDev{nullptr, x.seed, (uint32_t)x.seg_len, (uint32_t)dev_tbins, (uint32_t)x.bins, (uint32_t)v->bin_off[i], 1u,
                           (uint32_t)count_len};
        ix.max_tbins = std::max<uint32_t>(ix.max_tbins, (uint32_t)dev_tbins);
        ix.any_large = ix.any_large || dev_tbins > kSmallRowBytes;
        ix.fp_bytes += view_rows(x) * dev_tbins;
    }
    TRY(ix.arena.ensure_exact(arena_bytes));
    for (uint64_t i = 0; i < n; ++i)
        ix.ixf[i].fp = ix.arena.as<uint8_t>() + arena_off[i];
    TRY(upload_fingerprints(c, v, ix));
    // per-bin metadata: kind, user bin, child, first bin of the split run (hixf.hpp:313-338)
    std::vector<int32_t> ub(total_bins), child(total_bins);
    std::vector<uint32_t> run_begin(total_bins);
    std::vector<uint8_t> kind(total_bins);
    for (uint64_t i = 0; i < n; ++i)
    {
        const uint64_t o = v->bin_off[i], nb = v->ixf[i].bins;
        uint32_t run_start = 0;
        for (uint64_t b = 0; b < nb; ++b)
        {
            const int64_t cur = v->bin_to_user_bin[o + b];
            if (cur < 0)
            {
                const int64_t nx = v->next_ixf_id[o + b];
                if (nx < 0 || (uint64_t)nx >= n || (uint64_t)nx == i)
                    return set_error(TXR_ERR_ARG, "IXF %llu bin %llu: merged bin without a child IXF", (unsigned long long)i,
                                     (unsigned long long)b);
                kind[o + b] = kBinMerged;
                child[o + b] = (int32_t)nx;
                ub[o + b] = -1;
                run_start = (uint32_t)b + 1;
            }
            else
            {
                if ((uint64_t)cur >= v->n_user_bins)
                    return set_error(TXR_ERR_ARG, "user bin id %lld out of range", (long long)cur);
                const bool end = b + 1 == nb || cur != v->bin_to_user_bin[o + b + 1];
                kind[o + b] = end ? kBinRunEnd : kBinMid;
                ub[o + b] = (int32_t)cur;
                child[o + b] = -1;
                run_begin[o + b] = run_start;
                if (end)
                {
                    ix.ixf[i].max_run = std::max<uint32_t>(ix.ixf[i].max_run, (uint32_t)b + 1 - run_start);
                    run_start = (uint32_t)b + 1;
                }
            }
        }
    }
    // DFS pre-order rank of every user bin + tree depth (iterative DFS from IXF 0, bins ascending)
    ix.dfs_rank.assign(v->n_user_bins, 0xffffffffu);
    ix.n_user_bins = v->n_user_bins;
    {
        struct Frame
        {
            uint64_t ixf, bin;
            uint32_t level;
        };
        std::vector<Frame> stack{{0, 0, 1}};
        std::vector<uint8_t> seen(n, 0);
        seen[0] = 1;
        uint32_t rank = 0, depth = 1;
        while (!stack.empty())
        {
            Frame &f = stack.back();
            if (f.bin == v->ixf[f.ixf].bins)
            {
                stack.pop_back();
                continue;
            }
            const uint64_t at = v->bin_off[f.ixf] + f.bin++;
            if (kind[at] == kBinMerged)
            {
                const uint64_t ch = (uint64_t)child[at];
                if (seen[ch])
                    return set_error(TXR_ERR_ARG, "IXF %llu is reachable twice: not a tree", (unsigned long long)ch);
                seen[ch] = 1;
                const uint32_t lvl = f.level + 1;
                depth = std::max(depth, lvl);
                stack.push_back(Frame{ch, 0, lvl});
            }
            else if (kind[at] == kBinRunEnd)
            {
                if (ix.dfs_rank[ub[at]] == 0xffffffffu)
                    ix.dfs_rank[ub[at]] = rank++;
            }
        }
        ix.depth = depth;
        if (depth > C_MAX_LEVELS)
            return set_error(TXR_ERR_UNSUPPORTED, "HIXF deeper than %d levels", (int)C_MAX_LEVELS);
    }
    TRY(ix.d_ixf.ensure(n * sizeof(IxfDev)));
    TRY(ix.d_bin_ub.ensure(total_bins * 4));
    TRY(ix.d_bin_child.ensure(total_bins * 4));
    TRY(ix.d_bin_run_begin.ensure(total_bins * 4));
    TRY(ix.d_bin_kind.ensure(total_bins));
    CU(cudaMemcpy(ix.d_ixf.p, ix.ixf.data(), n * sizeof(IxfDev), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ix.d_bin_ub.p, ub.data(), total_bins * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ix.d_bin_child.p, child.data(), total_bins * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ix.d_bin_run_begin.p, run_begin.data(), total_bins * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ix.d_bin_kind.p, kind.data(), total_bins, cudaMemcpyHostToDevice));
    ix.loaded = true;
    return TXR_OK;
}

// Replicates the index of `src` (same or another GPU) into `dst` without touching the host copy again: the arena
// goes device to device (NVLink when the GPUs are peers), the small metadata arrays likewise.  This is how a multi-GPU
// driver fills its contexts: one upload over PCIe, then a doubling tree of clones (taxor_main.cpp).
int txr_index_clone(txr_ctx *dst, txr_ctx *src)
{
    if (!dst || !src || dst == src)
        return set_error(TXR_ERR_ARG, "null or identical contexts");
    if (!src->index.loaded)
        return set_error(TXR_ERR_STATE, "source context has no index");
    CU(cudaSetDevice(src->device));
    CU(cudaDeviceSynchronize());
    CU(cudaSetDevice(dst->device));
    CU(cudaDeviceSynchronize());
    if (dst->device != src->device)
    {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
        if (can)
        {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return set_error(TXR_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    const DeviceIndex &a = src->index;
    DeviceIndex &b = dst->index;
    b.release();
    b.ixf = a.ixf;
    b.dfs_rank = a.dfs_rank;
    b.depth = a.depth;
    b.max_tbins = a.max_tbins;
    b.any_large = a.any_large;
    b.n_user_bins = a.n_user_bins;
    b.fp_bytes = a.fp_bytes;
    b.scheme = a.scheme;
    b.generic = a.generic;
    struct Pair
    {
        const DevBuf *from;
        DevBuf *to;
    };
    const Pair bufs[] = {{&a.arena, &b.arena}, {&a.d_bin_ub, &b.d_bin_ub}, {&a.d_bin_child, &b.d_bin_child},
                         {&a.d_bin_run_begin, &b.d_bin_run_begin}, {&a.d_bin_kind, &b.d_bin_kind}};
    for (const Pair &p : bufs)
    {
        if (!p.from->cap)
            continue;
        TRY(p.to->ensure_exact(p.from->cap));
        CU(cudaMemcpyPeer(p.to->p, dst->device, p.from->p, src->device, p.from->cap));
    }
    const uint8_t *base_a = a.arena.as<uint8_t>();
    for (auto &d : b.ixf)
        d.fp = b.arena.as<uint8_t>() + (d.fp - base_a);
    TRY(b.d_ixf.ensure(b.ixf.size() * sizeof(IxfDev)));
    CU(cudaMemcpy(b.d_ixf.p, b.ixf.data(), b.ixf.size() * sizeof(IxfDev), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    b.loaded = true;
    return TXR_OK;
}

int txr_params_set(txr_ctx *c, const txr_params *p)
{
    if (!c || !p)
        return set_error(TXR_ERR_ARG, "null argument");
    if (p->kmer_size < 1 || p->kmer_size > 32)
        return set_error(TXR_ERR_ARG, "kmer_size %d out of range", p->kmer_size);
    if (p->use_syncmer)
    {
        const int wn = p->kmer_size - p->syncmer_size + 1;
        if (p->syncmer_size < 1 || p->syncmer_size >= p->kmer_size || p->t_syncmer < 1 || p->t_syncmer > wn)
            return set_error(TXR_ERR_ARG, "syncmer parameters k=%d s=%d t=%d out of range", p->kmer_size, p->syncmer_size, p->t_syncmer);
    }
    else if (p->window_size < p->kmer_size)
        return set_error(TXR_ERR_ARG, "window_size %u smaller than kmer_size %d", p->window_size, p->kmer_size);
    else if (p->window_size - p->kmer_size + 1 > (uint32_t)kMaxMinimiserValues)
        return set_error(TXR_ERR_UNSUPPORTED, "minimiser windows of more than %d k-mers are not supported (taxor build limits --window-size to 96)",
                         kMaxMinimiserValues);
    CU(cudaSetDevice(c->device));
    c->params = *p;
    if (c->params.scaling == 0)
        c->params.scaling = 1;
    c->thresholder = Thresholder(p->window_size, p->kmer_size, p->percentage, p->error_rate, p->use_syncmer != 0);
    c->kmer_seed = 0x8F3F73B5CF1C9ADEULL >> (64u - 2u * p->kmer_size); // src/hixf/build/adjust_seed.hpp:40-44
    c->per_read_thr = c->thresholder.kind() == ThresholdKind::fracminhash;
    c->lut.clear();
    c->d_lut_len = 0;
    c->have_params = true;
    return ensure_lut(c, 4096);
}

int txr_threshold_get(txr_ctx *c, uint64_t hash_count, double scaling_factor, uint64_t *out)
{
    if (!c || !out || !c->have_params)
        return set_error(TXR_ERR_STATE, "no parameters");
    *out = c->thresholder.get(hash_count, scaling_factor);
    return TXR_OK;
}

int txr_threshold_eval(const txr_params *p, uint64_t hash_count, double scaling_factor, uint64_t *out)
{
    if (!p || !out)
        return set_error(TXR_ERR_ARG, "null argument");
    *out = Thresholder(p->window_size, p->kmer_size, p->percentage, p->error_rate, p->use_syncmer != 0).get(hash_count, scaling_factor);
    return TXR_OK;
}

// ---- packing ----
uint64_t txr_packed_words(uint64_t n_bases) { return (n_bases + 31) / 32 + 1; }

struct Dna4Table
{
    int8_t tab[256];
    Dna4Table()
    {
        memset(tab, -1, sizeof tab);
        const char *groups[4] = {"ANRWMDHV", "CYSB", "GK", "TU"}; // seqan3::dna4 char_to_rank
        for (int r = 0; r < 4; ++r)
            for (const char *q = groups[r]; *q; ++q)
            {
                tab[(unsigned char)*q] = (int8_t)r;
                tab[(unsigned char)(*q + 32)] = (int8_t)r;
            }
    }
};
static const int8_t *dna4_table()
{
    static const Dna4Table t; // C++11 magic static: txr_pack_2bit is first called from many pack threads at once
    return t.tab;
}

int txr_pack_2bit(const char *ascii, uint64_t len, uint64_t *dst)
{
    const int8_t *tab = dna4_table();
    const uint64_t nw = txr_packed_words(len);
    uint64_t i = 0;
    for (uint64_t w = 0; w + 1 < nw; ++w)
    {
        if (len - i >= 32) // whole words of plain ACGT/acgt take the SIMD lane; it stops at the first word with anything else
        {
            const size_t done = txr::pack_plain_words(ascii + i, (size_t)((len - i) / 32), dst + w);
            if (done)
            {
                w += done - 1;
                i += 32 * (uint64_t)done;
                continue;
            }
        }
        uint64_t x = 0;
        const uint64_t end = std::min<uint64_t>(len, i + 32);
        int sh = 62;
        for (; i < end; ++i, sh -= 2)
        {
            const int8_t r = tab[(unsigned char)ascii[i]];
            if (r < 0)
                return set_error(TXR_ERR_FORMAT, "illegal nucleotide character 0x%02x at position %llu", (unsigned char)ascii[i],
                                 (unsigned long long)i);
            x |= (uint64_t)r << sh;
        }
        dst[w] = x;
    }
    dst[nw - 1] = 0;
    return TXR_OK;
}

int txr_pack_codes(const uint8_t *codes, uint64_t len, uint64_t *dst)
{
    const uint64_t nw = txr_packed_words(len);
    uint64_t i = 0;
    for (uint64_t w = 0; w + 1 < nw; ++w)
    {
        uint64_t x = 0;
        const uint64_t end = std::min<uint64_t>(len, i + 32);
        int sh = 62;
        for (; i < end; ++i, sh -= 2)
        {
            if (codes[i] > 3)
                return set_error(TXR_ERR_FORMAT, "base code %d at position %llu", codes[i], (unsigned long long)i);
            x |= (uint64_t)codes[i] << sh;
        }
        dst[w] = x;
    }
    dst[nw - 1] = 0;
    return TXR_OK;
}

int txr_unpack_codes(const uint64_t *words, uint64_t len, uint8_t *codes)
{
    for (uint64_t i = 0; i < len; ++i)
        codes[i] = (uint8_t)((words[i >> 5] >> (62 - 2 * (i & 31))) & 3);
    return TXR_OK;
}

void *txr_host_alloc(size_t bytes)
{
    void *p = nullptr;
    // TXR_HOST_WC=1: write-combined pinned memory (the CPU only ever streams packed reads INTO these buffers; DMA reads of
    // write-combined memory skip the cache snoop).  An experiment knob for the many-GPU end-to-end curve; CPU reads are slow.
    static const bool wc = getenv("TXR_HOST_WC") && atoi(getenv("TXR_HOST_WC")) != 0;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess)
    {
        set_error(TXR_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void *txr_ctx_host_alloc(txr_ctx *c, size_t bytes)
{
    if (!c)
    {
        set_error(TXR_ERR_ARG, "null context");
        return nullptr;
    }
    if (cudaSetDevice(c->device) != cudaSuccess)
    {
        set_error(TXR_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        return nullptr;
    }
    return txr_host_alloc(bytes);
}
void txr_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

// ---- search ----
static int validate_reads(const uint64_t *word_off, const uint32_t *len, uint64_t n)
{
    for (uint64_t i = 0; i + 1 < n; ++i)
        if (word_off[i + 1] < word_off[i] + txr_packed_words(len[i]))
            return set_error(TXR_ERR_ARG, "reads must be packed in ascending, non-overlapping order (read %llu)", (unsigned long long)i);
    return TXR_OK;
}

// A search that fails half way (overflow, allocation failure) must not leave batches in flight: their slots would still
// be `busy` with Slot::bm pointing into the failed call's stack, and the next call on the context would collect them.
static int fail_clean(txr_ctx *c, int rc)
{
    if (rc != TXR_OK && c)
    {
        const std::string keep = g_last_error;
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (auto &s : c->slots)
        {
            s->busy = false;
            s->bm = nullptr;
            s->bd = nullptr;
            s->d_words = nullptr;
        }
        c->result.clear();
        g_last_error = keep;
    }
    return rc;
}

static int search_host_impl(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                            txr_result *out, std::vector<BatchMeta> &metas);

int txr_search(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
               txr_result *out)
{
    std::vector<BatchMeta> metas; // outlives every batch in flight: fail_clean() drains the device before it goes
    return fail_clean(c, search_host_impl(c, words, word_off, len, n_reads, out, metas));
}

static int search_host_impl(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                            txr_result *out, std::vector<BatchMeta> &metas)
{
    TRY(check_ready(c));
    if (!words || !word_off || !len || !out)
        return set_error(TXR_ERR_ARG, "null argument");
    c->result.clear();
    c->timing = txr_timing{};
    const auto t0 = std::chrono::steady_clock::now();
    TRY(fork_streams(c));
    metas.resize(c->n_slots);
    const size_t S = (size_t)c->n_slots;
    // Batches are planned and validated one at a time, right before they are submitted: two passes over a million reads in
    // front of the first copy were 2 ms in which the GPU had nothing to do (the same rule as plan_batches, ramp included).
    double scale = c->n_slots > 1 ? 1.0 / 16.0 : 1.0;
    uint64_t next_read = 0;
    for (size_t b = 0; next_read < n_reads; ++b)
    {
        if (b >= S - 1)
        {
            Slot &done = *c->slots[(b - (S - 1)) % S];
            if (done.busy)
                TRY(collect_batch(c, done, true));
        }
        uint64_t max_reads = std::max<uint64_t>((uint64_t)((double)c->max_batch_reads * scale), 1);
        const uint64_t max_bases = std::max<uint64_t>((uint64_t)((double)c->max_batch_bases * scale), 1);
        // With the hash stage of batch b+1 running beside the probes of batch b, a batch must not be much larger than what was
        // submitted before it: its copy ends late and its (slowed) hash stage would outlast the probes it hides behind
        // (profiles/r2_trace_e2e_timeline.txt).  Cap: a fraction of all reads submitted so far, never below 1/4 of a full batch.
        if (c->ramp_cum > 0 && overlap_applies(c))
            max_reads = std::min<uint64_t>(max_reads, std::max<uint64_t>({(uint64_t)(c->ramp_cum * (double)next_read), c->max_batch_reads / 4, (uint64_t)1}));
        uint64_t bases = 0, j = next_read;
        while (j < n_reads && j - next_read < max_reads && (j == next_read || bases + len[j] <= max_bases))
            bases += len[j++];
        scale = std::min(1.0, scale * c->ramp);
        TRY(validate_reads(word_off + next_read, len + next_read, std::min<uint64_t>(j - next_read + 1, n_reads - next_read)));
        Slot &s = *c->slots[b % S];
        if (s.busy)
            TRY(collect_batch(c, s, true));
        BatchMeta &m = metas[b % S];
        build_batch_meta(c, word_off, len, next_read, (uint32_t)(j - next_read), m);
        TRY(submit_batch(c, s, m, nullptr, nullptr, words, true, b == 0, j >= n_reads));
        next_read = j;
    }
    // the batches still in flight, oldest first (slots are used round robin)
    for (size_t k = 0; k < S; ++k)
    {
        Slot *oldest = nullptr;
        for (auto &sl : c->slots)
            if (sl->busy && (!oldest || sl->bm->first_read < oldest->bm->first_read))
                oldest = sl.get();
        if (!oldest)
            break;
        TRY(collect_batch(c, *oldest, true));
    }
    TRY(join_streams(c));
    finish_result(c, out);
    c->timing.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return TXR_OK;
}

int txr_reads_upload(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                     txr_reads **out)
{
    TRY(check_ready(c));
    if (!words || !word_off || !len || !out || n_reads == 0)
        return set_error(TXR_ERR_ARG, "null or empty argument");
    TRY(validate_reads(word_off, len, n_reads));
    auto r = std::make_unique<txr_reads>();
    r->n_reads = n_reads;
    const uint64_t first_word = word_off[0];
    const uint64_t total_words = word_off[n_reads - 1] + txr_packed_words(len[n_reads - 1]) - first_word;
    TRY(r->words.ensure(total_words * 8));
    CU(cudaMemcpy(r->words.p, words + first_word, total_words * 8, cudaMemcpyHostToDevice));
    std::vector<std::pair<uint64_t, uint32_t>> plan;
    plan_batches(c, len, n_reads, plan);
    r->batches.resize(plan.size());
    for (size_t b = 0; b < plan.size(); ++b)
    {
        BatchMeta &m = r->batches[b];
        build_batch_meta(c, word_off, len, plan[b].first, plan[b].second, m);
        m.first_word -= first_word; // now relative to the resident word array
        r->dev.push_back(std::make_unique<BatchDev>());
        TRY(upload_batch_meta(m, *r->dev.back(), nullptr, nullptr));
    }
    CU(cudaDeviceSynchronize());
    *out = r.release();
    return TXR_OK;
}

void txr_reads_free(txr_ctx *c, txr_reads *r)
{
    if (!r)
        return;
    if (c)
        cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    r->words.release();
    for (auto &d : r->dev)
        for (DevBuf *b : {&d->word_off, &d->len, &d->out_off, &d->ids_small, &d->ids_medium, &d->ids_global, &d->gtable_off})
            b->release();
    delete r;
}

static int search_resident_impl(txr_ctx *c, txr_reads *r, int fetch, txr_result *out);
int txr_search_resident(txr_ctx *c, txr_reads *r, int fetch, txr_result *out)
{
    return fail_clean(c, search_resident_impl(c, r, fetch, out));
}
static int search_resident_impl(txr_ctx *c, txr_reads *r, int fetch, txr_result *out)
{
    TRY(check_ready(c));
    if (!r || (fetch && !out))
        return set_error(TXR_ERR_ARG, "null argument");
    c->result.clear();
    c->timing = txr_timing{};
    const auto t0 = std::chrono::steady_clock::now();
    TRY(fork_streams(c));
    const size_t S = (size_t)c->n_slots;
    for (size_t b = 0; b < r->batches.size(); ++b)
    {
        Slot &s = *c->slots[b % S];
        if (s.busy)
            TRY(collect_batch(c, s, fetch != 0));
        const BatchMeta &m = r->batches[b];
        TRY(submit_batch(c, s, m, r->dev[b].get(), r->words.as<uint64_t>() + m.first_word, nullptr, true, b == 0, b + 1 == r->batches.size()));
    }
    // collect in submission order
    for (size_t b = (r->batches.size() > S ? r->batches.size() - S : 0); b < r->batches.size(); ++b)
    {
        Slot &s = *c->slots[b % S];
        if (s.busy)
            TRY(collect_batch(c, s, fetch != 0));
    }
    TRY(join_streams(c));
    if (fetch)
        finish_result(c, out);
    c->timing.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return TXR_OK;
}

int txr_get_timing(txr_ctx *c, txr_timing *out)
{
    if (!c || !out)
        return set_error(TXR_ERR_ARG, "null argument");
    *out = c->timing;
    return TXR_OK;
}

// ---- parity entry points ----
int txr_hash_batch(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                   int dedup, const uint64_t **hash_off, const uint64_t **hashes)
{
    if (!c || !c->have_params)
        return set_error(TXR_ERR_STATE, "txr_params_set not called");
    CU(cudaSetDevice(c->device));
    TRY(ensure_slots(c));
    if (!words || !word_off || !len || !hash_off || !hashes)
        return set_error(TXR_ERR_ARG, "null argument");
    TRY(validate_reads(word_off, len, n_reads));
    c->hb_off.assign(1, 0);
    c->hb_hashes.clear();
    c->shape_hash = c->shape_dedup = 0; // this entry point runs the hash stage alone: full grids, every SM
    c->adaptive_small_hash = c->adaptive_small_dedup = 0;
    c->smf_hash = SmFilter{0, 0, 0};
    std::vector<std::pair<uint64_t, uint32_t>> plan;
    plan_batches(c, len, n_reads, plan);
    Slot &s = *c->slots[0];
    BatchMeta m;
    std::vector<uint64_t> h_hashes;
    std::vector<uint32_t> h_cnt;
    for (auto &pb : plan)
    {
        build_batch_meta(c, word_off, len, pb.first, pb.second, m);
        TRY(slot_reserve(c, s, m));
        TRY(s.words.ensure(m.n_words * 8));
        CU(cudaMemcpyAsync(s.words.p, words + m.first_word, m.n_words * 8, cudaMemcpyHostToDevice, s.stream));
        TRY(upload_batch_meta(m, s.meta, s.stream, nullptr));
        CU(cudaMemsetAsync(s.counters.p, 0, C_TOTAL * 4, s.stream));
        TRY(launch_hash_stage(c, s, m, s.meta, s.words.as<uint64_t>(), dedup != 0, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        uint32_t ovf = 0;
        CU(cudaMemcpy(&ovf, s.counters.as<uint32_t>() + C_HASH_OVERFLOW, 4, cudaMemcpyDeviceToHost));
        if (ovf)
            return set_error(TXR_ERR_OVERFLOW, "hash capacity bound violated");
        h_hashes.resize(std::max<uint64_t>(m.total_cap, 1));
        h_cnt.resize(m.n_reads);
        CU(cudaMemcpy(h_hashes.data(), s.hashes.p, m.total_cap * 8, cudaMemcpyDeviceToHost));
        const bool counted = dedup && (c->params.use_syncmer || c->params.scaling > 1);
        const bool raw_in_nraw = c->params.use_syncmer || c->params.scaling > 1;
        const void *src = counted || !raw_in_nraw ? s.hash_count.p : s.n_raw.p;
        CU(cudaMemcpy(h_cnt.data(), src, (size_t)m.n_reads * 4, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < m.n_reads; ++i)
        {
            c->hb_hashes.insert(c->hb_hashes.end(), h_hashes.begin() + m.out_off[i], h_hashes.begin() + m.out_off[i] + h_cnt[i]);
            c->hb_off.push_back(c->hb_hashes.size());
        }
    }
    *hash_off = c->hb_off.data();
    *hashes = c->hb_hashes.data();
    return TXR_OK;
}

// ---- build side: per-user-bin distinct hashes ----
namespace
{
inline uint64_t host_mer(const uint64_t *w, uint64_t pos, int n)
{
    const uint64_t wi = pos >> 5;
    const int off = (int)(pos & 31) * 2;
    uint64_t x = w[wi];
    if (off)
        x = (x << off) | (w[wi + 1] >> (64 - off));
    return x >> (64 - 2 * n);
}
inline uint64_t host_revcomp(uint64_t fwd, int n)
{
    uint64_t y = ~fwd, r = 0;
    for (int i = 0; i < n; ++i, y >>= 2)
        r = (r << 2) | (y & 3);
    return r;
}
// Is the scan state at k-mer window j independent of everything before it?  Syncmers: the k-s+1 canonical s-mers of the
// window have ONE minimum (syncmer.cpp:113-141 then tracks exactly that position, whatever happened earlier, and a scan
// that starts here picks the same one).  Minimisers: the W values starting at j have one minimum.  k-mers: stateless.
bool state_is_history_free(const txr_params &p, uint64_t kmer_seed, const uint64_t *w, uint64_t j)
{
    int n_vals, mer;
    if (p.use_syncmer)
    {
        n_vals = p.kmer_size - p.syncmer_size + 1;
        mer = p.syncmer_size;
    }
    else
    {
        n_vals = (int)p.window_size - p.kmer_size + 1;
        mer = p.kmer_size;
        if (n_vals <= 1)
            return true;
    }
    uint64_t best = ~0ULL;
    int count = 0;
    for (int i = 0; i < n_vals; ++i)
    {
        const uint64_t f = host_mer(w, j + i, mer), r = host_revcomp(f, mer);
        const uint64_t v = p.use_syncmer ? std::min(f, r) : std::min(f ^ kmer_seed, r ^ kmer_seed);
        if (v < best)
        {
            best = v;
            count = 1;
        }
        else if (v == best)
            ++count;
    }
    return count == 1;
}
// target k-mer windows per segment (TXR_SEGMENT_WINDOWS overrides it: the tests cut every few thousand windows)
uint64_t segment_windows()
{
    const char *e = getenv("TXR_SEGMENT_WINDOWS");
    const long long v = e ? atoll(e) : 0;
    return v >= 64 ? (uint64_t)v : (1u << 18);
}
} // namespace

// Cut points (k-mer window indices, multiples of 32, ascending, first = 0) of one sequence: a cut is only placed at a
// window whose scan state is history free, so the pieces [cut_i, cut_{i+1}) can be hashed as independent reads.
static void plan_cuts(const txr_params &p, uint64_t kmer_seed, const uint64_t *w, uint64_t L, uint64_t target, std::vector<uint64_t> &cuts)
{
    const int k = p.kmer_size;
    const int span = p.use_syncmer ? k : std::max<int>(k, (int)p.window_size); // bases a window's state test reads
    const uint64_t n_win = L >= (uint64_t)k ? L - k + 1 : 0;
    cuts.assign(1, 0);
    uint64_t start = 0;
    while (n_win > start + target + target / 2)
    {
        uint64_t cand = (start + target + 31) & ~31ull, cut = 0;
        for (int tries = 0; tries < 256 && cand + span <= L && cand + target / 4 < n_win; ++tries, cand += 32)
            if (state_is_history_free(p, kmer_seed, w, cand))
            {
                cut = cand;
                break;
            }
        if (!cut)
            break; // no history-free window nearby (a long repeat): the rest of the sequence stays one piece
        cuts.push_back(cut);
        start = cut;
    }
}

int txr_plan_segments(const txr_params *p, const uint64_t *words, uint64_t len, uint64_t target_windows, uint64_t *cuts,
                      uint64_t cap, uint64_t *n_cuts)
{
    if (!p || !words || !n_cuts || (!cuts && cap) || target_windows < 64)
        return set_error(TXR_ERR_ARG, "bad argument");
    if (p->kmer_size < 1 || p->kmer_size > 32 || (p->use_syncmer && (p->syncmer_size < 1 || p->syncmer_size >= p->kmer_size)) ||
        (!p->use_syncmer && (p->window_size < p->kmer_size || p->window_size - p->kmer_size + 1 > (uint32_t)kMaxMinimiserValues)))
        return set_error(TXR_ERR_ARG, "parameters out of range");
    std::vector<uint64_t> v;
    plan_cuts(*p, 0x8F3F73B5CF1C9ADEULL >> (64u - 2u * p->kmer_size), words, len, target_windows, v);
    *n_cuts = v.size();
    for (size_t i = 0; i < v.size() && i < cap; ++i)
        cuts[i] = v[i];
    return v.size() <= cap ? TXR_OK : set_error(TXR_ERR_OVERFLOW, "%zu cuts, room for %llu", v.size(), (unsigned long long)cap);
}

int txr_hash_user_bins(txr_ctx *c, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_seqs,
                       const uint32_t *seq_bin, uint64_t n_bins, txr_bin_hashes *out)
{
    if (!c || !c->have_params)
        return set_error(TXR_ERR_STATE, "txr_params_set not called");
    if (!words || !word_off || !len || !seq_bin || !out || n_bins == 0 || n_bins > 0xffffffffull)
        return set_error(TXR_ERR_ARG, "null or empty argument");
    CU(cudaSetDevice(c->device));
    TRY(ensure_slots(c));
    for (uint64_t i = 0; i < n_seqs; ++i)
    {
        if (seq_bin[i] >= n_bins || (i && seq_bin[i] < seq_bin[i - 1]))
            return set_error(TXR_ERR_ARG, "seq_bin must be non-decreasing and below n_bins (sequence %llu)", (unsigned long long)i);
        if (i + 1 < n_seqs && word_off[i + 1] < word_off[i] + txr_packed_words(len[i]))
            return set_error(TXR_ERR_ARG, "sequences must be packed in ascending, non-overlapping order (sequence %llu)", (unsigned long long)i);
    }
    const txr_params &p = c->params;
    const int k = p.kmer_size;
    // 1. segments: cut long sequences at history-free windows that start on a word boundary
    std::vector<uint64_t> seg_off;
    std::vector<uint32_t> seg_len, seg_bin;
    const int span = p.use_syncmer ? k : std::max<int>(k, (int)p.window_size);
    const uint64_t target = segment_windows();
    std::vector<uint64_t> cuts;
    for (uint64_t q = 0; q < n_seqs; ++q)
    {
        const uint64_t L = len[q];
        plan_cuts(p, c->kmer_seed, words + word_off[q], L, target, cuts);
        for (size_t i = 0; i < cuts.size(); ++i)
        {
            const bool last = i + 1 == cuts.size();
            seg_off.push_back(word_off[q] + cuts[i] / 32);
            seg_len.push_back((uint32_t)(last ? L - cuts[i] : cuts[i + 1] - cuts[i] + span - 1));
            seg_bin.push_back(seq_bin[q]);
        }
    }
    const uint64_t n_seg = seg_off.size();
    c->ub_off.assign(n_bins + 1, 0);
    c->ub_hashes.clear();
    c->shape_hash = c->shape_dedup = 0;
    c->adaptive_small_hash = c->adaptive_small_dedup = 0;
    c->smf_hash = SmFilter{0, 0, 0};
    Slot &s = *c->slots[0];
    cudaStream_t st = s.stream;
    DevBuf &d_seg_bin = c->binset[0], &d_tables = c->binset[1], &d_table_off = c->binset[2], &d_flags = c->binset[3],
           &d_out = c->binset[4], &d_out_off = c->binset[5], &d_count = c->binset[6], &d_work = c->binset[7];
    auto release_all = [&] {
        for (auto &b : c->binset)
            b.release();
    };
    std::vector<uint64_t> bin_total(n_bins, 0);
    // 2. batches of whole user bins
    uint64_t a = 0;
    while (a < n_seg)
    {
        uint64_t b = a, bases = 0;
        while (b < n_seg && (b == a || seg_bin[b] == seg_bin[b - 1] || bases + seg_len[b] <= c->max_batch_bases))
            bases += seg_len[b++]; // a bin is never split over two batches
        const uint32_t n = (uint32_t)(b - a);
        const uint32_t bin_lo = seg_bin[a], bin_hi = seg_bin[b - 1], nb = bin_hi - bin_lo + 1;
        BatchMeta m;
        build_batch_meta(c, seg_off.data(), seg_len.data(), a, n, m);
        // segments overlap by k-1 bases, so their word ranges do too: the batch's words run to the end of the last one
        uint64_t last_word = 0;
        for (uint64_t i = a; i < b; ++i)
            last_word = std::max(last_word, seg_off[i] + txr_packed_words(seg_len[i]));
        m.n_words = last_word - m.first_word;
        int rc = slot_reserve(c, s, m);
        if (rc == TXR_OK)
            rc = s.words.ensure(m.n_words * 8);
        if (rc != TXR_OK)
        {
            release_all();
            return rc;
        }
        CU(cudaMemcpyAsync(s.words.p, words + m.first_word, m.n_words * 8, cudaMemcpyHostToDevice, st));
        TRY(upload_batch_meta(m, s.meta, st, nullptr));
        CU(cudaMemsetAsync(s.counters.p, 0, C_TOTAL * 4, st));
        TRY(launch_hash_stage(c, s, m, s.meta, s.words.as<uint64_t>(), false, st));
        // raw counts -> table sizes
        std::vector<uint32_t> n_raw(n);
        const bool raw_in_nraw = p.use_syncmer || p.scaling > 1;
        CU(cudaMemcpyAsync(n_raw.data(), raw_in_nraw ? s.n_raw.p : s.hash_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        uint32_t ovf = 0;
        CU(cudaMemcpyAsync(&ovf, s.counters.as<uint32_t>() + C_HASH_OVERFLOW, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ovf)
        {
            release_all();
            return set_error(TXR_ERR_OVERFLOW, "hash capacity bound violated");
        }
        std::vector<uint64_t> raw_per_bin(nb, 0), table_off(nb + 1, 0), out_off(nb + 1, 0);
        std::vector<uint32_t> local_bin(n);
        for (uint32_t i = 0; i < n; ++i)
        {
            local_bin[i] = seg_bin[a + i] - bin_lo;
            raw_per_bin[local_bin[i]] += n_raw[i];
        }
        for (uint32_t q = 0; q < nb; ++q)
        {
            table_off[q + 1] = table_off[q] + std::max<uint64_t>(32, next_pow2(2 * raw_per_bin[q]));
            out_off[q + 1] = out_off[q] + raw_per_bin[q] + 1;
            if (table_off[q + 1] - table_off[q] > (1ull << 32))
            {
                release_all();
                return set_error(TXR_ERR_UNSUPPORTED, "user bin %u: more than 2^31 raw hashes", bin_lo + q);
            }
        }
        rc = d_seg_bin.ensure((size_t)n * 4);
        rc = rc ? rc : d_tables.ensure(table_off[nb] * 8);
        rc = rc ? rc : d_table_off.ensure((nb + 1) * 8);
        rc = rc ? rc : d_flags.ensure((size_t)nb * 4);
        rc = rc ? rc : d_out.ensure(out_off[nb] * 8);
        rc = rc ? rc : d_out_off.ensure((nb + 1) * 8);
        rc = rc ? rc : d_count.ensure((size_t)nb * 4);
        rc = rc ? rc : d_work.ensure(4);
        if (rc != TXR_OK)
        {
            release_all();
            return rc;
        }
        CU(cudaMemcpyAsync(d_seg_bin.p, local_bin.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_table_off.p, table_off.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_out_off.p, out_off.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(d_tables.p, 0xff, table_off[nb] * 8, st));
        CU(cudaMemsetAsync(d_flags.p, 0, (size_t)nb * 4, st));
        CU(cudaMemsetAsync(d_count.p, 0, (size_t)nb * 4, st));
        CU(cudaMemsetAsync(d_work.p, 0, 4, st));
        BinSetArgs ba{};
        ba.hashes = s.hashes.as<uint64_t>();
        ba.out_off = s.meta.out_off.as<uint64_t>();
        ba.n_raw = raw_in_nraw ? s.n_raw.as<uint32_t>() : s.hash_count.as<uint32_t>();
        ba.seg_bin = d_seg_bin.as<uint32_t>();
        ba.n_segments = n;
        ba.tables = d_tables.as<uint64_t>();
        ba.table_off = d_table_off.as<uint64_t>();
        ba.n_bins = nb;
        ba.bin_has_empty_key = d_flags.as<uint32_t>();
        ba.out = d_out.as<uint64_t>();
        ba.out_bin_off = d_out_off.as<uint64_t>();
        ba.bin_count = d_count.as<uint32_t>();
        ba.work = d_work.as<uint32_t>();
        ba.scaling = p.scaling;
        ba.scaling_limit = double(UINT64_MAX) / double(p.scaling ? p.scaling : 1);
        CU(launch_binset(ba, c->sm_count, st));
        std::vector<uint32_t> counts(nb);
        CU(cudaMemcpyAsync(counts.data(), d_count.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // bins ascend with the segments and a bin never spans two batches: appending keeps the result in bin order
        uint64_t at = c->ub_hashes.size(), add = 0;
        for (uint32_t q = 0; q < nb; ++q)
            add += counts[q];
        c->ub_hashes.resize(at + add);
        for (uint32_t q = 0; q < nb; ++q)
        {
            bin_total[bin_lo + q] = counts[q];
            if (counts[q])
                CU(cudaMemcpyAsync(c->ub_hashes.data() + at, d_out.as<uint64_t>() + out_off[q], (size_t)counts[q] * 8, cudaMemcpyDeviceToHost, st));
            at += counts[q];
        }
        CU(cudaStreamSynchronize(st));
        a = b;
    }
    for (uint64_t q = 0; q < n_bins; ++q)
        c->ub_off[q + 1] = c->ub_off[q] + bin_total[q];
    release_all();
    out->n_bins = n_bins;
    out->bin_off = c->ub_off.data();
    out->hashes = c->ub_hashes.data();
    out->n_segments = n_seg;
    return TXR_OK;
}

int txr_ixf_bulk_count(txr_ctx *c, uint64_t ixf_idx, const uint64_t *values, uint64_t n, uint32_t *counts)
{
    if (!c || !c->index.loaded)
        return set_error(TXR_ERR_STATE, "no index uploaded");
    if (ixf_idx >= c->index.ixf.size() || (!values && n) || !counts || n > 0xffffffffu)
        return set_error(TXR_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    TRY(ensure_slots(c));
    const IxfDev &d = c->index.ixf[ixf_idx];
    TRY(c->scratch_a.ensure(std::max<uint64_t>(n, 1) * 8));
    TRY(c->scratch_b.ensure((size_t)d.tbins * 4));
    cudaStream_t st = c->slots[0]->stream;
    if (n)
        CU(cudaMemcpyAsync(c->scratch_a.p, values, n * 8, cudaMemcpyHostToDevice, st));
    CU(launch_bulk_count(d, c->index.scheme, c->scratch_a.as<uint64_t>(), (uint32_t)n, c->scratch_b.as<uint32_t>(), st));
    CU(cudaMemcpyAsync(counts, c->scratch_b.p, (size_t)d.bins * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return TXR_OK;
}

} // extern "C"
