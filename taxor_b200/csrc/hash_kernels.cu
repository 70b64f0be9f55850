// hash_kernels.cu -- kernel #1 of the `taxor search` hot path for sm_100a.
//
//   syncmer_kernel<K,S,T>   open canonical syncmers -> wyhash (replaces hashing::seq_to_syncmers,
//                           src/hashing/syncmer.cpp:80-165, before the set insert)
//   syncmer_generic_kernel  same for any (k,s,t): one thread per read, sequential restatement
//   kmer_kernel             canonical k-mers XOR seed (replaces seq | minimiser_hash with window == k,
//                           src/main/taxor_search.cpp:210-212,240-256)
//   dedup_warp_kernel /     per-read distinct set (the ankerl::unordered_dense::set of syncmer.cpp:145) and the
//   dedup_smem_kernel /     FracMin scaling filter (taxor_search.cpp:223-233, 244-251); syncmer_kernel can also build the
//   dedup_global_kernel /   set itself while hashing (HashArgs::fuse_dedup, measured slower, off by default)
//   filter_kernel
//   binset_*_kernel         build side: per-user-bin distinct sets (compute_hashes.cpp:76-142)
// The hash-stage kernels have adaptive grids (yield_to_probes): launched for the whole GPU, their CTAs beyond a small share
// leave while probe kernels of the previous batch are running (engine.cu: overlap).
//
// Design (not a translation): reads are 2-bit packed MSB-first, so the forward code of any s-mer / k-mer is a
// bit field of the packed stream and its reverse complement is a bit field of the bit-reversed, complemented
// stream.  One warp owns one read and walks it in tiles of 1024 windows; each lane owns 32 consecutive windows
// whose 42 canonical s-mers live in registers.  A k-mer is an open syncmer iff the s-mer at offset t-1 is the
// minimum of its k-s+1 s-mers, i.e. iff it is a STRICT minimum of the t-1 s-mers to its left and the k-s+1-t to
// its right: two fixed-width sliding minima computed by doubling.  Whenever that s-mer only TIES the
// neighbourhood minimum, the reference's history-dependent tie rule decides (leftmost minimum in the first
// window, rightmost on a rescan, no replacement on equal arrival; syncmer.cpp:116-140); such tiles are replayed
// sequentially by lane 0 with the exact rule, anchored at the nearest window whose minimum is unique.
#include "device_types.cuh"
#include "ixf_arith.cuh"

#include <algorithm>

namespace txr
{
constexpr int kHashWarps = 4; // warps per CTA of the syncmer kernel

namespace
{
__device__ __forceinline__ bool scaling_keep(uint64_t h, uint32_t scaling, double limit)
{
    // taxor_search.cpp:227-228: double(wyhash(h)) <= double(UINT64_MAX) / double(scaling)
    return scaling <= 1 || __ull2double_rn(wyhash_u64(h)) <= limit;
}

__device__ __forceinline__ uint64_t pair_swap(uint64_t y)
{
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
}
// reverse complement of a 64-base block (2-bit groups reversed, each complemented)
__device__ __forceinline__ uint64_t revcomp_block(uint64_t x) { return pair_swap(__brevll(~x)); }

// forward code of the n-mer starting at base `pos` of a packed read (n <= 32)
__device__ __forceinline__ uint64_t mer_at(const uint64_t *__restrict__ w, uint64_t pos, int n)
{
    const uint64_t wi = pos >> 5;
    const int off = (int)(pos & 31) * 2;
    uint64_t x = w[wi];
    if (off)
        x = (x << off) | (w[wi + 1] >> (64 - off));
    return x >> (64 - 2 * n);
}
__device__ __forceinline__ uint64_t revcomp_mer(uint64_t fwd, int n)
{
    return pair_swap(__brevll(~fwd)) >> (64 - 2 * n);
}
__device__ __forceinline__ uint64_t canon_mer_at(const uint64_t *__restrict__ w, uint64_t pos, int n)
{
    const uint64_t f = mer_at(w, pos, n);
    const uint64_t r = revcomp_mer(f, n);
    return f < r ? f : r;
}

// static bit-field extraction from a 128-bit value held as four 32-bit registers (x[3] most significant)
template <int START, int LEN>
__device__ __forceinline__ uint32_t field128(const uint32_t (&x)[4])
{
    constexpr int sh = 128 - 2 * (START + LEN);
    static_assert(sh >= 0 && 2 * LEN <= 32, "field out of range");
    constexpr int wi = sh >> 5, b = sh & 31;
    constexpr uint32_t mask = (2 * LEN == 32) ? 0xffffffffu : ((1u << (2 * LEN)) - 1u);
    const uint32_t lo = x[wi];
    const uint32_t hi = (wi + 1 < 4) ? x[wi + 1 < 4 ? wi + 1 : 3] : 0u;
    return __funnelshift_r(lo, hi, b) & mask;
}

template <int Q, int QN, int S>
struct SmerFill
{
    __device__ __forceinline__ static void run(const uint32_t (&f)[4], const uint32_t (&r)[4], uint32_t (&v)[QN])
    {
        const uint32_t a = field128<Q, S>(f);
        const uint32_t b = field128<64 - Q - S, S>(r);
        v[Q] = min(a, b);
        SmerFill<Q + 1, QN, S>::run(f, r, v);
    }
};
template <int QN, int S>
struct SmerFill<QN, QN, S>
{
    __device__ __forceinline__ static void run(const uint32_t (&)[4], const uint32_t (&)[4], uint32_t (&)[QN]) {}
};


// min(v[start .. start+N-1]) from the doubling tables (all indices are compile-time after unrolling)
template <int N, int QN>
__device__ __forceinline__ uint32_t range_min(const uint32_t (&v)[QN], const uint32_t (&m2)[QN], const uint32_t (&m4)[QN],
                                              const uint32_t (&m8)[QN], const uint32_t (&m16)[QN], int start)
{
    if constexpr (N == 0)
        return 0xffffffffu;
    else if constexpr (N == 1)
        return v[start];
    else if constexpr (N < 4)
        return min(m2[start], m2[start + N - 2]);
    else if constexpr (N < 8)
        return min(m4[start], m4[start + N - 4]);
    else if constexpr (N < 16)
        return min(m8[start], m8[start + N - 8]);
    else
        return min(m16[start], m16[start + N - 16]);
}

// exact sequential state of the reference scan --------------------------------------------------------------
struct ScanState
{
    uint64_t min_val;
    uint64_t min_pos; // position (s-mer start) of the current window minimum
};

// minimum of window j (s-mers j .. j+wn-1): value, leftmost and rightmost position
__device__ void window_min(const uint64_t *w, uint64_t j, int wn, int s, uint64_t &mv, uint64_t &lm, uint64_t &rm)
{
    mv = ~0ULL;
    lm = rm = j;
    for (int q = 0; q < wn; ++q)
    {
        const uint64_t v = canon_mer_at(w, j + q, s);
        if (v < mv)
        {
            mv = v;
            lm = rm = j + q;
        }
        else if (v == mv)
            rm = j + q;
    }
}

// state after window j, given the state after window j-1 (syncmer.cpp:125-140); v_new = s-mer j+wn-1
__device__ __forceinline__ void scan_step(const uint64_t *w, uint64_t j, int wn, int s, uint64_t v_new, ScanState &st)
{
    if (st.min_pos == j - 1) // the minimum left the window: rescan, rightmost minimum wins
    {
        uint64_t mv, lm, rm;
        window_min(w, j, wn, s, mv, lm, rm);
        st.min_val = mv;
        st.min_pos = rm;
    }
    else if (v_new < st.min_val) // strictly smaller arrival
    {
        st.min_val = v_new;
        st.min_pos = j + wn - 1;
    }
}

// exact state after window j0 of a read, derived from the nearest anchor (a window with a unique minimum,
// or window 0 where the leftmost minimum is taken; syncmer.cpp:116-123)
__device__ ScanState scan_state_at(const uint64_t *w, uint64_t j0, int wn, int s)
{
    uint64_t a = j0;
    uint64_t mv, lm, rm;
    while (true)
    {
        window_min(w, a, wn, s, mv, lm, rm);
        if (lm == rm || a == 0)
            break;
        --a;
    }
    ScanState st{mv, lm};
    for (uint64_t j = a + 1; j <= j0; ++j)
        scan_step(w, j, wn, s, canon_mer_at(w, j + wn - 1, s), st);
    return st;
}
} // namespace

// -----------------------------------------------------------------------------------------------------------
// fast syncmer kernel: one warp per read, 2S <= 32
// -----------------------------------------------------------------------------------------------------------
namespace
{
// 32 stream bits starting at base START of the 128-bit value x (x[3] most significant), left-aligned
template <int START>
__device__ __forceinline__ uint32_t top32(const uint32_t (&x)[4])
{
    constexpr int lo_bit = 128 - 2 * START - 32; // bit index of the lowest of the 32 bits (may be negative)
    static_assert(128 - 2 * START > 0, "field starts beyond the value");
    if constexpr (lo_bit >= 0)
    {
        constexpr int wi = lo_bit >> 5, b = lo_bit & 31;
        return __funnelshift_r(x[wi], wi + 1 < 4 ? x[wi + 1 < 4 ? wi + 1 : 3] : 0u, b);
    }
    else
        return x[0] << (-lo_bit);
}

// canonical s-mers Q..QN-1 of a lane: min(forward field, reverse-complement field), right-aligned
template <int Q, int QN, int S>
struct CanonFill
{
    __device__ __forceinline__ static void run(const uint32_t (&f)[4], const uint32_t (&r)[4], uint32_t (&v)[QN])
    {
        // both fields are taken left-aligned (low bits carry stream garbage); the order of the top 2S bits decides
        // the minimum and the shift drops the garbage of whichever won
        v[Q] = min(top32<Q>(f), top32<64 - Q - S>(r)) >> (32 - 2 * S);
        CanonFill<Q + 1, QN, S>::run(f, r, v);
    }
};
template <int QN, int S>
struct CanonFill<QN, QN, S>
{
    __device__ __forceinline__ static void run(const uint32_t (&)[4], const uint32_t (&)[4], uint32_t (&)[QN]) {}
};

// Exact decision for ONE window j whose candidate s-mer ties the window minimum: replay the reference's state
// machine (syncmer.cpp:116-140) from the nearest anchor (a window with a unique minimum, or window 0).
// sv holds the canonical s-mers of the current tile (index q - tile); earlier positions come from the packed read.
// Returns false through `give_up` when the tied run is longer than the local budget.
__device__ bool resolve_tie_local(const uint64_t *w, const uint32_t *sv, uint64_t tile, uint64_t j, int wn, int s, int t,
                                  bool &give_up)
{
    auto val = [&](uint64_t q) -> uint32_t { return q >= tile ? sv[q - tile] : (uint32_t)canon_mer_at(w, q, s); };
    auto wmin = [&](uint64_t a, uint32_t &mv, uint64_t &lm, uint64_t &rm)
    {
        mv = 0xffffffffu;
        lm = rm = a;
        for (int q = 0; q < wn; ++q)
        {
            const uint32_t x = val(a + q);
            if (x < mv)
            {
                mv = x;
                lm = rm = a + q;
            }
            else if (x == mv)
                rm = a + q;
        }
    };
    uint64_t a = j, lm, rm;
    uint32_t mv;
    for (int steps = 0;; ++steps)
    {
        wmin(a, mv, lm, rm);
        if (lm == rm || a == 0)
            break;
        if (steps >= 40)
        {
            give_up = true;
            return false;
        }
        --a;
    }
    uint32_t st_val = mv;
    uint64_t st_pos = lm; // unique minimum, or the leftmost one in window 0 (syncmer.cpp:116-123)
    for (uint64_t jj = a + 1; jj <= j; ++jj)
    {
        if (st_pos == jj - 1) // popped: rescan, rightmost minimum (syncmer.cpp:128-136)
        {
            wmin(jj, mv, lm, rm);
            st_val = mv;
            st_pos = rm;
        }
        else
        {
            const uint32_t x = val(jj + wn - 1);
            if (x < st_val) // strictly smaller arrival (syncmer.cpp:137-140)
            {
                st_val = x;
                st_pos = jj + wn - 1;
            }
        }
    }
    return st_pos == j + t - 1;
}
} // namespace

// MINB: CTAs per SM the register allocation has to allow (4: 124 registers, no spills; 5: 102 registers, experiment knob
// TXR_HASH_REGS=5 -- more warps for an issue-bound kernel against a few spills)
template <int K, int S, int T, int MINB>
__global__ void __launch_bounds__(32 * kHashWarps, MINB) syncmer_kernel(HashArgs a)
{
    if (!sm_filter_keep(a.smf) || yield_to_probes(a.probe_flag, a.small_grid))
        return;
    constexpr int WN = K - S + 1;    // s-mers per k-mer window
    constexpr int QN = 32 + WN - 1;  // s-mers a lane needs for its 32 windows
    constexpr int NL = T - 1;        // neighbourhood left of the candidate s-mer
    constexpr int NR = WN - T;       // ... and right of it
    constexpr bool kSignTrick = 2 * S <= 30; // values < 2^31: comparisons through the sign of a difference
    static_assert(32 + K - 1 <= 64, "two packed words per lane");
    static_assert(T >= 1 && T <= WN, "t out of range");

    // per warp: canonical s-mers of the tile (tie handling only), later reused for the compacted selected windows
    __shared__ uint32_t s_v[kHashWarps][kTileWindows + 32];
    __shared__ uint32_t s_sel[kHashWarps][32];      // per-lane selection masks written by the sequential replay
    // fused distinct set: kWarpSlots u32 per warp (dynamic: with it the CTA exceeds the 48 KB static limit).  A slot is
    // 21 tag bits of the key | 11 bits: 0 empty, 1..kWarpMaxKeys = 1 + position of the key in the read's output list,
    // kPendingBase + n = claimed by key n of the group of keys being inserted (staged in shared memory until it settles).
    extern __shared__ uint32_t s_tab_dyn[];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    volatile uint32_t *tab = s_tab_dyn + (a.fuse_dedup ? wib * kWarpSlots : 0);
    constexpr uint32_t kPendingBase = 1920; // + the key's number inside its group of 128
    static_assert(kWarpMaxKeys < kPendingBase && kPendingBase + 127 < 2048, "slot payload ranges overlap");
    while (true)
    {
        // a CTA beyond the small share re-checks before every read: it leaves as soon as a probe kernel starts, so that the probe
        // CTAs find registers (a persistent full grid that only looked once would hold the SMs until the whole batch is hashed)
        uint32_t r = 0xffffffffu;
        if (lane == 0 && !yield_to_probes(a.probe_flag, a.small_grid))
            r = atomicAdd(a.work_counter, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_reads)
            break;
        const uint32_t L = a.len[r];
        const uint64_t W = L >= (uint32_t)K ? (uint64_t)L - K + 1 : 0;
        const uint64_t *__restrict__ w = a.words + a.word_off[r];
        const uint64_t nw = ((uint64_t)L + 31) / 32 + 1;
        uint64_t *__restrict__ out = a.out + a.out_off[r];
        const uint64_t cap = a.out_off[r + 1] - a.out_off[r];
        uint64_t cursor = 0;
        bool carry_valid = false;
        ScanState carry{0, 0};
        // set_mode: the hashes go through the warp's table and only first occurrences are written (distinct list);
        // it ends (for the rest of the read) when the table would pass its load limit -- the tail is then written raw
        // behind the distinct prefix and a CTA-per-read kernel finishes the read (same distinct set).
        bool set_mode = a.fuse_dedup && cap <= kFuseMaxCap;
        const bool fused_read = set_mode;
        uint32_t emitted = 0; // raw emissions (capacity check)
        if (set_mode)
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(s_tab_dyn + wib * kWarpSlots);
#pragma unroll
            for (int i = 0; i < kWarpSlots / 128; ++i)
                t4[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
        }

        uint64_t hi_next = (uint64_t)lane < nw && W ? w[lane] : 0;
        for (uint64_t tile = 0; tile < W; tile += kTileWindows)
        {
            // this tile's word was fetched while the previous tile was processed; start the next fetch now
            const uint64_t hi = hi_next;
            const uint64_t g_next = ((tile + kTileWindows) >> 5) + lane;
            hi_next = g_next < nw ? w[g_next] : 0;
            uint64_t lo = __shfl_down_sync(0xffffffffu, hi, 1);
            const uint64_t lo31 = __shfl_sync(0xffffffffu, hi_next, 0); // word 32 of this tile = word 0 of the next
            if (lane == 31)
                lo = lo31;
            const uint64_t rhi = revcomp_block(lo), rlo = revcomp_block(hi);
            const uint32_t f[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
            const uint32_t rc[4] = {(uint32_t)rlo, (uint32_t)(rlo >> 32), (uint32_t)rhi, (uint32_t)(rhi >> 32)};

            uint32_t v[QN];
            CanonFill<0, QN, S>::run(f, rc, v);

            // sliding minima by doubling: mN[q] = min(v[q .. q+N-1]) (clamped at the end of the lane's range)
            uint32_t m2[QN], m4[QN], m8[QN], m16[QN];
#pragma unroll
            for (int q = 0; q < QN; ++q)
                m2[q] = q + 1 < QN ? min(v[q], v[q + 1]) : v[q];
#pragma unroll
            for (int q = 0; q < QN; ++q)
                m4[q] = q + 2 < QN ? min(m2[q], m2[q + 2]) : m2[q];
#pragma unroll
            for (int q = 0; q < QN; ++q)
                m8[q] = q + 4 < QN ? min(m4[q], m4[q + 4]) : m4[q];
#pragma unroll
            for (int q = 0; q < QN; ++q)
                m16[q] = q + 8 < QN ? min(m8[q], m8[q + 8]) : m8[q];

            // candidate s-mer of window i is v[i+T-1]; selected iff it is a strict minimum of its NL left and NR
            // right neighbours; an equality is a tie that the exact rule has to decide
            uint32_t sel = 0;
            bool any_tie = false;
#pragma unroll
            for (int i = 31; i >= 0; --i)
            {
                const uint32_t other = min(range_min<NL>(v, m2, m4, m8, m16, i), range_min<NR>(v, m2, m4, m8, m16, i + T));
                const uint32_t c = v[i + T - 1];
                if constexpr (kSignTrick)
                    sel = __funnelshift_l(c - other, sel, 1); // shifts in the sign bit of (c - other)
                else
                    sel = (sel << 1) | (c < other ? 1u : 0u);
                any_tie |= c == other;
            }
            const uint64_t j0 = tile + 32ull * lane;
            const uint32_t valid = j0 >= W ? 0u : (W - j0 >= 32 ? 0xffffffffu : ((1u << (uint32_t)(W - j0)) - 1u));
            sel &= valid;

            bool replayed = false;
            if (__any_sync(0xffffffffu, any_tie && valid != 0))
            {
                uint32_t tie = 0;
#pragma unroll
                for (int i = 31; i >= 0; --i)
                {
                    const uint32_t other = min(range_min<NL>(v, m2, m4, m8, m16, i), range_min<NR>(v, m2, m4, m8, m16, i + T));
                    tie = (tie << 1) | (v[i + T - 1] == other ? 1u : 0u);
                }
                tie &= valid;
                if (__any_sync(0xffffffffu, tie != 0))
                {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        s_v[wib][32 * lane + i] = v[i];
                    if (lane == 31)
                    {
#pragma unroll
                        for (int i = 32; i < QN; ++i)
                            s_v[wib][32 * lane + i] = v[i];
                    }
                    __syncwarp();
                    // (1) every lane settles its own tied windows with a short local replay
                    bool give_up = false;
                    uint32_t m = tie;
                    while (m && !give_up)
                    {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        if (resolve_tie_local(w, s_v[wib], tile, j0 + b, WN, S, T, give_up))
                            sel |= 1u << b;
                    }
                    // (2) long tied runs (low-complexity sequence): exact replay of the whole tile by lane 0
                    if (__any_sync(0xffffffffu, give_up))
                    {
                        replayed = true;
                        s_sel[wib][lane] = 0;
                        __syncwarp();
                        if (lane == 0)
                        {
                            const uint64_t jend = min(W, tile + (uint64_t)kTileWindows);
                            ScanState st;
                            uint64_t j = tile;
                            if (tile == 0)
                            {
                                uint64_t mv, lm, rm;
                                window_min(w, 0, WN, S, mv, lm, rm);
                                st = ScanState{mv, lm};
                                if (st.min_pos == (uint64_t)(T - 1))
                                    s_sel[wib][0] |= 1u;
                                j = 1;
                            }
                            else
                                st = carry_valid ? carry : scan_state_at(w, tile - 1, WN, S);
                            for (; j < jend; ++j)
                            {
                                const uint32_t lj = (uint32_t)(j - tile);
                                if (st.min_pos == j - 1)
                                {
                                    uint32_t mv = 0xffffffffu, mp = 0;
                                    for (int q = WN - 1; q >= 0; --q) // rightmost minimum
                                    {
                                        const uint32_t x = s_v[wib][lj + q];
                                        if (x < mv)
                                        {
                                            mv = x;
                                            mp = q;
                                        }
                                    }
                                    st.min_val = mv;
                                    st.min_pos = j + mp;
                                }
                                else
                                {
                                    const uint32_t x = s_v[wib][lj + WN - 1];
                                    if (x < st.min_val)
                                    {
                                        st.min_val = x;
                                        st.min_pos = j + WN - 1;
                                    }
                                }
                                if (st.min_pos == j + T - 1)
                                    s_sel[wib][lj >> 5] |= 1u << (lj & 31);
                            }
                            carry = st;
                        }
                        __syncwarp();
                        sel = s_sel[wib][lane];
                        carry.min_val = __shfl_sync(0xffffffffu, carry.min_val, 0);
                        carry.min_pos = __shfl_sync(0xffffffffu, carry.min_pos, 0);
                    }
                    __syncwarp();
                }
            }
            carry_valid = replayed;

            // compact the selected windows of the tile, then hash them with all lanes busy
            const uint32_t n = __popc(sel);
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d)
                    incl += t2;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_v[wib]);
            uint32_t at = incl - n;
            uint32_t m = sel;
            while (m)
            {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                s_pos[at++] = (uint16_t)(32 * lane + b);
            }
            __syncwarp();
            emitted += total;
            // Keys are handled kGroup = 32 * KB at a time (a typical tile of a random read selects ~93 windows: one group): all
            // KB hashes of a lane are computed first -- their packed-word loads overlap -- and the claim rounds then work on
            // KB independent table slots per lane.  (A version that inserted 32 keys at a time exposed one load latency and two
            // or three warp-synchronous rounds per chunk and lost 12 ms per 1 M reads against the separate dedup kernel.)
            constexpr int KB = 4;
            volatile uint64_t *s_keys = reinterpret_cast<volatile uint64_t *>(s_v[wib] + 512); // 128 keys; s_pos ends below byte 2048
            for (uint32_t it0 = 0; it0 < total; it0 += 32 * KB)
            {
                uint64_t h[KB];
                bool live[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j)
                {
                    const uint32_t it = it0 + 32 * j + lane;
                    live[j] = it < total;
                    h[j] = live[j] ? wyhash_u64(canon_mer_at(w, tile + s_pos[live[j] ? it : 0], K)) : 0;
                }
                const uint32_t group = min((uint32_t)(32 * KB), total - it0);
                if (set_mode && cursor + group > a.fuse_max_keys)
                    set_mode = false; // warp-uniform: the rest of the read is written raw
                if (!set_mode)
                {
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        if (live[j] && cursor + 32 * j + lane < cap)
                            out[cursor + 32 * j + lane] = h[j];
                    cursor += group;
                    continue;
                }
                // ---- the ankerl::set insert of syncmer.cpp:145 for one group of keys ----
                // FracMin first (taxor_search.cpp:223-233 filters the set; filtering before the insert gives the same set)
                bool pending[KB], first[KB];
                uint32_t slot[KB], mine[KB];
                __syncwarp(); // the previous group's readers of s_keys are done
#pragma unroll
                for (int j = 0; j < KB; ++j)
                {
                    s_keys[32 * j + lane] = h[j];
                    pending[j] = live[j] && scaling_keep(h[j], a.scaling, a.scaling_limit);
                    first[j] = false;
                    slot[j] = (uint32_t)(h[j] ^ (h[j] >> 29)) & (kWarpSlots - 1);
                    mine[j] = ((uint32_t)(h[j] >> 43) << 11) | (kPendingBase + 32u * j + (uint32_t)lane);
                }
                __syncwarp();
                while (true)
                {
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        any |= pending[j];
                    if (!__any_sync(0xffffffffu, any))
                        break;
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        if (pending[j] && tab[slot[j]] == 0)
                            tab[slot[j]] = mine[j]; // racy on purpose: lanes of this round that share the slot, one store lands
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                    {
                        if (!pending[j])
                            continue;
                        const uint32_t now = tab[slot[j]];
                        const uint32_t low = now & 2047u;
                        if (now == mine[j])
                        {
                            first[j] = true;
                            pending[j] = false;
                        }
                        // an equal tag: compare the keys -- a key of this group (staged in shared memory) or an earlier one
                        // (already in the output list)
                        else if (low != 0 && (now & ~2047u) == (mine[j] & ~2047u) &&
                                 (low >= kPendingBase ? s_keys[low - kPendingBase] : __ldcg(out + (low - 1))) == h[j])
                            pending[j] = false; // an earlier (or concurrent) occurrence owns this key
                        else
                            slot[j] = (slot[j] + 1) & (kWarpSlots - 1);
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < KB; ++j)
                {
                    const uint32_t bal = __ballot_sync(0xffffffffu, first[j]);
                    if (first[j])
                    {
                        const uint32_t idx = (uint32_t)cursor + __popc(bal & ((1u << lane) - 1u));
                        tab[slot[j]] = (mine[j] & ~2047u) | (idx + 1); // settle the claim: position in the output list
                        if (idx < cap)                                 // always, unless the capacity bound is violated (reported below)
                            out[idx] = h[j];
                    }
                    cursor += __popc(bal);
                }
                __syncwarp();
            }
            __syncwarp();
        }
        if (lane == 0)
        {
            if (emitted > cap)
                *a.overflow = 1u;
            if (fused_read && set_mode)
                a.hash_count[r] = (uint32_t)cursor;
            else
            {
                a.n_out[r] = (uint32_t)min(cursor, cap);
                if (fused_read)
                    a.deferred[atomicAdd(a.n_deferred, 1u)] = r;
            }
        }
    }
}

// -----------------------------------------------------------------------------------------------------------
// generic syncmer kernel: any (k <= 32, s < k, t); one thread per read, sequential scan with the reference's
// state machine (syncmer.cpp:97-146).  Fallback only -- correctness first.
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) syncmer_generic_kernel(HashArgs a)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0)
        atomicAdd(a.work_counter, a.n_reads); // not work-stealing, but the engine checks this counter after a batch
    if (r >= a.n_reads)
        return;
    const int K = a.k, S = a.s, T = a.t, WN = K - S + 1;
    const uint32_t L = a.len[r];
    const uint64_t W = L >= (uint32_t)K ? (uint64_t)L - K + 1 : 0;
    const uint64_t *w = a.words + a.word_off[r];
    uint64_t *out = a.out + a.out_off[r];
    const uint64_t cap = a.out_off[r + 1] - a.out_off[r];
    uint64_t cursor = 0;
    ScanState st{0, 0};
    for (uint64_t j = 0; j < W; ++j)
    {
        if (j == 0)
        {
            uint64_t mv, lm, rm;
            window_min(w, 0, WN, S, mv, lm, rm);
            st = ScanState{mv, lm};
        }
        else
            scan_step(w, j, WN, S, canon_mer_at(w, j + WN - 1, S), st);
        if (st.min_pos == j + T - 1)
        {
            if (cursor < cap)
                out[cursor] = wyhash_u64(canon_mer_at(w, j, K));
            ++cursor;
        }
    }
    a.n_out[r] = (uint32_t)min(cursor, cap);
    if (cursor > cap)
        *a.overflow = 1u;
}

// -----------------------------------------------------------------------------------------------------------
// canonical k-mer kernel (window == k): one warp per read; window j of a tile is handled by lane j % 32 so
// that the 8-byte stores of a warp are contiguous.
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kmer_kernel(HashArgs a)
{
    if (!sm_filter_keep(a.smf))
        return;
    const int lane = threadIdx.x & 31;
    const int K = a.k;
    while (true)
    {
        uint32_t r = 0;
        if (lane == 0)
            r = atomicAdd(a.work_counter, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_reads)
            break;
        const uint32_t L = a.len[r];
        const uint64_t W = L >= (uint32_t)K ? (uint64_t)L - K + 1 : 0;
        const uint64_t *__restrict__ w = a.words + a.word_off[r];
        const uint64_t nw = ((uint64_t)L + 31) / 32 + 1;
        uint64_t *__restrict__ out = a.out + a.out_off[r];
        const uint64_t cap = a.out_off[r + 1] - a.out_off[r];
        for (uint64_t tile = 0; tile < W; tile += kTileWindows)
        {
            const uint64_t g = (tile >> 5) + lane;
            const uint64_t mine = g < nw ? w[g] : 0;
            const uint64_t extra = (lane == 0 && g + 32 < nw) ? w[g + 32] : 0; // word 32 of the tile
#pragma unroll 4
            for (int i = 0; i < 32; ++i)
            {
                const uint64_t hi = __shfl_sync(0xffffffffu, mine, i);
                uint64_t lo = __shfl_sync(0xffffffffu, mine, (i + 1) & 31);
                const uint64_t ex = __shfl_sync(0xffffffffu, extra, 0);
                if (i == 31)
                    lo = ex;
                const uint64_t j = tile + 32ull * i + lane;
                const int off = 2 * lane;
                const uint64_t x = off ? (hi << off) | (lo >> (64 - off)) : hi;
                const uint64_t fw = x >> (64 - 2 * K);
                const uint64_t rv = revcomp_mer(fw, K);
                const uint64_t fs = fw ^ a.kmer_seed, rs = rv ^ a.kmer_seed;
                if (j < W && j < cap)
                    out[j] = fs < rs ? fs : rs;
            }
        }
        if (lane == 0)
        {
            a.n_out[r] = (uint32_t)min(W, cap);
            if (W > cap)
                *a.overflow = 1u;
        }
    }
}

// -----------------------------------------------------------------------------------------------------------
// minimiser kernel (window_size > k): seq | seqan3::views::minimiser_hash(ungapped k, window_size, seed)
// (taxor_search.cpp:210-212,242).  Upstream SeqAn3 3.3.0 semantics of views::minimiser over the canonical k-mer
// values v[q] = min(fwd ^ seed, rc ^ seed): W = window_size - k + 1 values per window (clamped to the number of
// values of a short read); the tracked minimiser is the RIGHTMOST minimum of the first window; on every shift it is
// re-chosen (rightmost minimum of the new window) and reported again when it was the value that left, replaced and
// reported when the arriving value is STRICTLY smaller, kept silently otherwise.  Duplicates are kept.
//
// One warp per read, tiles of kTileWindows windows.  The values a tile needs are staged in shared memory once
// (coalesced extraction from the packed words); lane l then runs the state machine over its 32 consecutive windows.
// The state a lane starts from is history dependent only where the preceding window holds its minimum more than
// once (repeats, palindromic k-mer pairs): a window with a unique minimum pins the state whatever came before.  Lane 0
// starts from the state carried across tiles; a lane whose preceding window is tied waits for its left neighbour
// (sequential hand-over, homopolymers and tandem repeats only).  Reported positions are buffered per lane and written
// in window order after a warp prefix sum.
// -----------------------------------------------------------------------------------------------------------
constexpr int kMinWarps = 4;
constexpr int kMinVals = kTileWindows + kMaxMinimiserValues - 1;       // values a tile can need
constexpr int kMinValsPadded = kMinVals + kMinVals / 32 + 1;           // one pad slot per 32: lanes 32 windows apart hit different banks
namespace
{
__device__ __forceinline__ int mpad(int e) { return e + (e >> 5); }

struct MinState
{
    int pos;       // tile-relative index of the tracked minimiser (may be -1: it just left the window)
};

// rightmost / leftmost minimum of the W values starting at tile-relative index `first`
__device__ __forceinline__ void window_extrema(const uint64_t *sv, int first, int W, int &leftmost, int &rightmost)
{
    uint64_t best = sv[mpad(first)];
    leftmost = rightmost = first;
    for (int j = first + 1; j < first + W; ++j)
    {
        const uint64_t x = sv[mpad(j)];
        if (x < best)
        {
            best = x;
            leftmost = rightmost = j;
        }
        else if (x == best)
            rightmost = j;
    }
}

// runs windows [a, b) (tile-relative) from state `pos`; records the reported positions; returns the final state
__device__ __forceinline__ int minimiser_run(const uint64_t *sv, int a, int b, int W, int pos, uint16_t *ev, int &n_ev)
{
    for (int i = a; i < b; ++i)
    {
        if (pos < i)
        {
            int lm, rm;
            window_extrema(sv, i, W, lm, rm);
            pos = rm;
            ev[n_ev++] = (uint16_t)pos;
        }
        else
        {
            const int arriving = i + W - 1;
            if (sv[mpad(arriving)] < sv[mpad(pos)])
            {
                pos = arriving;
                ev[n_ev++] = (uint16_t)pos;
            }
        }
    }
    return pos;
}
} // namespace

__global__ void __launch_bounds__(32 * kMinWarps) minimiser_kernel(HashArgs a)
{
    if (!sm_filter_keep(a.smf))
        return;
    __shared__ uint64_t s_val[kMinWarps][kMinValsPadded];
    __shared__ uint16_t s_ev[kMinWarps][32][34]; // 34: lanes land on different banks
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint64_t *sv = s_val[wib];
    uint16_t *ev = s_ev[wib][lane];
    const int K = a.k;
    while (true)
    {
        uint32_t r = 0;
        if (lane == 0)
            r = atomicAdd(a.work_counter, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_reads)
            break;
        const uint32_t L = a.len[r];
        const uint64_t n_val = L >= (uint32_t)K ? (uint64_t)L - K + 1 : 0;
        const uint64_t *__restrict__ w = a.words + a.word_off[r];
        uint64_t *__restrict__ out = a.out + a.out_off[r];
        const uint64_t cap = a.out_off[r + 1] - a.out_off[r];
        const int W = (int)min((uint64_t)a.window, n_val);           // clamped window (iterator constructor)
        const uint64_t n_win = n_val ? n_val - W + 1 : 0;
        uint64_t cursor = 0;
        int carried = 0;                                              // tracked position relative to the NEXT tile
        for (uint64_t tile = 0; tile < n_win; tile += kTileWindows)
        {
            const int tw = (int)min((uint64_t)kTileWindows, n_win - tile);
            const int nv = tw + W - 1;
            __syncwarp();
            for (int e = lane; e < nv; e += 32)
            {
                const uint64_t fw = mer_at(w, tile + e, K);
                const uint64_t rv = revcomp_mer(fw, K);
                const uint64_t fs = fw ^ a.kmer_seed, rs = rv ^ a.kmer_seed;
                sv[mpad(e)] = fs < rs ? fs : rs;
            }
            __syncwarp();
            const int wa = lane * 32, wb = min(wa + 32, tw);
            const bool active = wa < tw;
            int n_ev = 0, pos = 0;
            bool resolved = true;
            if (active)
            {
                int start = wa;
                if (tile == 0 && lane == 0)
                {
                    int lm, rm;
                    window_extrema(sv, 0, W, lm, rm);                 // first window: rightmost minimum, reported
                    pos = rm;
                    ev[n_ev++] = (uint16_t)pos;
                    start = 1;
                }
                else if (lane == 0)
                    pos = carried;
                else
                {
                    int lm, rm;
                    window_extrema(sv, wa - 1, W, lm, rm);
                    pos = rm;
                    resolved = lm == rm;                              // unique minimum: the state is that position
                }
                if (resolved)
                    pos = minimiser_run(sv, start, wb, W, pos, ev, n_ev);
            }
            // sequential hand-over for lanes whose preceding window is tied (ascending, so the left neighbour is done)
            uint32_t pending = __ballot_sync(0xffffffffu, active && !resolved);
            while (pending)
            {
                const int l = __ffs(pending) - 1;
                const int from_left = __shfl_sync(0xffffffffu, pos, l - 1);
                if (lane == l)
                    pos = minimiser_run(sv, wa, wb, W, from_left, ev, n_ev);
                pending &= pending - 1;
            }
            const int last_lane = (tw - 1) >> 5;
            carried = __shfl_sync(0xffffffffu, pos, last_lane) - tw;  // may become -1: left the window, rescan next
            // ordered output
            int incl = n_ev;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d)
                    incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const uint64_t base = cursor + (uint64_t)(incl - n_ev);
            for (int e = 0; e < n_ev; ++e)
                if (base + e < cap)
                    out[base + e] = sv[mpad(ev[e])];
            cursor += (uint64_t)total;
        }
        if (lane == 0)
        {
            a.n_out[r] = (uint32_t)min(cursor, cap);
            if (cursor > cap)
                *a.overflow = 1u;
        }
    }
}

// -----------------------------------------------------------------------------------------------------------
// per-read distinct set + FracMin scaling filter
// -----------------------------------------------------------------------------------------------------------
namespace
{
template <typename table_ptr_t>
__device__ __forceinline__ void table_insert(table_ptr_t tab, uint32_t mask, uint64_t h)
{
    uint32_t slot = (uint32_t)(h ^ (h >> 32)) & mask;
    while (true)
    {
        const unsigned long long old = atomicCAS((unsigned long long *)&tab[slot], (unsigned long long)kEmptyKey, (unsigned long long)h);
        if (old == kEmptyKey || old == h)
            return;
        slot = (slot + 1) & mask;
    }
}

// compacts the non-empty slots of tab[0..slots) that pass the scaling filter into dst; returns the count.
// Block-wide; every thread must call it.
__device__ uint32_t table_compact(const uint64_t *tab, uint32_t slots, uint64_t *dst, bool extra_empty_key,
                                  uint32_t scaling, double limit, uint32_t *s_scan)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    uint32_t base = 0;
    for (uint32_t c0 = 0; c0 < slots; c0 += nt)
    {
        const uint32_t i = c0 + tid;
        uint64_t key = kEmptyKey;
        if (i < slots)
            key = tab[i];
        const bool keep = key != kEmptyKey && scaling_keep(key, scaling, limit);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            s_scan[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, all = 0;
        for (int q = 0; q < (nt >> 5); ++q)
        {
            const uint32_t c = s_scan[q];
            if (q < wid)
                before += c;
            all += c;
        }
        if (keep)
            dst[base + before + __popc(bal & ((1u << lane) - 1u))] = key;
        base += all;
        __syncthreads();
    }
    if (extra_empty_key && scaling_keep(kEmptyKey, scaling, limit))
    {
        if (tid == 0)
            dst[base] = kEmptyKey;
        ++base;
    }
    return base;
}
} // namespace

// One WARP per read.  The table (kWarpSlots x u32 in shared memory) stores 1 + the INDEX of the first occurrence of a
// key in the read's raw list; the 64-bit keys stay where kernel #1 wrote them.  Shared-memory atomicCAS turned out to
// be the bottleneck of an atomic version (~4 cycles per lane and SM on B200), so slots are claimed WITHOUT atomics,
// in warp-synchronous rounds: every pending lane looks at its slot, writes its index if the slot is empty (lanes of
// the same round may collide: one store lands), and after a __syncwarp() reads the slot back -- seeing its own
// index means the lane owns the first occurrence, any other index is compared by key (duplicate) or probed past.
// Pass 2 compacts the list in place in first-occurrence order (the order an ankerl set iterates in) and applies
// the FracMin filter.  Reads with more raw hashes than the table takes at load factor 3/4 are appended to
// `deferred` for the CTA-per-read kernel.
constexpr int kDedupWarps = 4; // 32 KB of static shared memory per CTA
__global__ void __launch_bounds__(32 * kDedupWarps) dedup_warp_kernel(DedupArgs a, uint32_t *work_counter, uint32_t *deferred,
                                                                      uint32_t *n_deferred)
{
    if (!sm_filter_keep(a.smf) || yield_to_probes(a.probe_flag, a.small_grid))
        return;
    __shared__ uint32_t s_tab[kDedupWarps][kWarpSlots];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    volatile uint32_t *tab = s_tab[wib];
    while (true)
    {
        uint32_t id = 0xffffffffu;
        if (lane == 0 && !yield_to_probes(a.probe_flag, a.small_grid)) // re-checked per read, as in the syncmer kernel
            id = atomicAdd(work_counter, 1u);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= a.n_ids)
            break;
        const uint32_t r = a.read_ids ? a.read_ids[id] : id;
        const uint32_t n = a.n_raw[r];
        if (n > kWarpMaxKeys)
        {
            if (lane == 0)
                deferred[atomicAdd(n_deferred, 1u)] = r;
            continue;
        }
        uint4 *t4 = reinterpret_cast<uint4 *>(s_tab[wib]);
#pragma unroll
        for (int i = 0; i < kWarpSlots / 128; ++i)
            t4[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        uint64_t *p = a.hashes + a.out_off[r];
        // Both passes stream the list in stages of kStage chunks whose loads are all issued before the first use:
        // the raw list comes from DRAM (a batch's lists exceed L2), and one exposed load latency per CHUNK was what
        // bounded earlier versions of this kernel (57 dependent round trips per read).
        constexpr int kStage = 16;
        // pass 1: bit c of `firsts` = this lane's key of chunk c is a first occurrence
        uint64_t firsts = 0; // kWarpMaxKeys / 32 = 48 chunks at most
        for (uint32_t s0 = 0; s0 < n; s0 += 32 * kStage)
        {
            uint64_t keys[kStage];
#pragma unroll
            for (int u = 0; u < kStage; ++u)
            {
                const uint32_t i = s0 + 32 * u + lane;
                keys[u] = i < n ? p[i] : 0;
            }
#pragma unroll
            for (int u = 0; u < kStage; ++u)
            {
                const uint32_t i = s0 + 32 * u + lane;
                if (s0 + 32 * u >= n)
                    break;
                const uint64_t key = keys[u];
                bool pending = i < n;
                uint32_t slot = (uint32_t)(key ^ (key >> 29)) & (kWarpSlots - 1);
                // slot word = 21 tag bits of the key | 11 bits (index + 1): a different tag proves a different key
                // without touching global memory
                const uint32_t mine = ((uint32_t)(key >> 43) << 11) | (i + 1);
                while (__any_sync(0xffffffffu, pending))
                {
                    if (pending && tab[slot] == 0)
                        tab[slot] = mine; // racy on purpose: lanes of this round that share the slot, one store lands
                    __syncwarp();
                    if (pending)
                    {
                        const uint32_t now = tab[slot];
                        if (now == mine)
                        {
                            firsts |= 1ull << ((s0 >> 5) + u);
                            pending = false;
                        }
                        else if ((now >> 11) == (mine >> 11) && p[(now & 2047u) - 1] == key)
                            pending = false; // an earlier (or concurrent) occurrence owns this key
                        else
                            slot = (slot + 1) & (kWarpSlots - 1);
                    }
                }
            }
        }
        __syncwarp();
        // Most reads have no repeated k-mer at all: every key was a first occurrence, the list is already the distinct set in
        // place, and the second pass (another read and a write of the whole list) has nothing to do.
        {
            const uint32_t mine_n = n > (uint32_t)lane ? (n - (uint32_t)lane + 31u) / 32u : 0u; // keys this lane looked at
            const bool clean = (uint32_t)__popcll(firsts) == mine_n;
            if (a.scaling <= 1 && __all_sync(0xffffffffu, clean))
            {
                if (lane == 0)
                    a.hash_count[r] = n;
                __syncwarp();
                continue;
            }
        }
        // pass 2: in-place compaction in first-occurrence order (a stage is loaded completely before it is written)
        uint32_t base = 0;
        for (uint32_t s0 = 0; s0 < n; s0 += 32 * kStage)
        {
            uint64_t keys[kStage];
#pragma unroll
            for (int u = 0; u < kStage; ++u)
            {
                const uint32_t i = s0 + 32 * u + lane;
                keys[u] = i < n ? p[i] : 0;
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kStage; ++u)
            {
                const uint32_t i = s0 + 32 * u + lane;
                const bool keep = i < n && ((firsts >> ((s0 >> 5) + u)) & 1) && scaling_keep(keys[u], a.scaling, a.scaling_limit);
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep)
                    p[base + __popc(bal & ((1u << lane) - 1u))] = keys[u];
                base += __popc(bal);
            }
        }
        if (lane == 0)
            a.hash_count[r] = base;
        __syncwarp();
    }
}

// one CTA per read, table in shared memory (SLOTS a power of two >= 2 * capacity of the class)
template <int SLOTS>
__global__ void dedup_smem_kernel(DedupArgs a)
{
    extern __shared__ uint64_t s_tab[];
    __shared__ uint32_t s_scan[32];
    __shared__ int s_saw_empty;
    const uint32_t id = blockIdx.x;
    if (id >= a.n_ids)
        return;
    const uint32_t r = a.read_ids ? a.read_ids[id] : id;
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x)
        s_tab[i] = kEmptyKey;
    if (threadIdx.x == 0)
        s_saw_empty = 0;
    __syncthreads();
    uint64_t *p = a.hashes + a.out_off[r];
    const uint32_t n = a.n_raw[r];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        const uint64_t h = p[i];
        if (h == kEmptyKey)
            s_saw_empty = 1;
        else
            table_insert(s_tab, SLOTS - 1, h);
    }
    __syncthreads();
    const uint32_t cnt = table_compact(s_tab, SLOTS, p, s_saw_empty != 0, a.scaling, a.scaling_limit, s_scan);
    if (threadIdx.x == 0)
        a.hash_count[r] = cnt;
}

// one CTA per read, table in global memory (pre-filled with kEmptyKey by the host)
__global__ void dedup_global_kernel(DedupArgs a)
{
    __shared__ uint32_t s_scan[32];
    __shared__ int s_saw_empty;
    const uint32_t id = blockIdx.x;
    if (id >= a.n_ids)
        return;
    const uint32_t r = a.read_ids ? a.read_ids[id] : id;
    uint64_t *tab = a.gtable + a.gtable_off[id];
    const uint32_t slots = (uint32_t)(a.gtable_off[id + 1] - a.gtable_off[id]);
    if (threadIdx.x == 0)
        s_saw_empty = 0;
    __syncthreads();
    uint64_t *p = a.hashes + a.out_off[r];
    const uint32_t n = a.n_raw[r];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        const uint64_t h = p[i];
        if (h == kEmptyKey)
            s_saw_empty = 1;
        else
            table_insert(tab, slots - 1, h);
    }
    __threadfence();
    __syncthreads();
    const uint32_t cnt = table_compact(tab, slots, p, s_saw_empty != 0, a.scaling, a.scaling_limit, s_scan);
    if (threadIdx.x == 0)
        a.hash_count[r] = cnt;
}

// k-mer mode with scaling > 1: order-preserving filter, one CTA per read (no dedup: duplicates are kept,
// taxor_search.cpp:242-255)
__global__ void filter_kernel(DedupArgs a)
{
    __shared__ uint32_t s_scan[32];
    const uint32_t r = blockIdx.x;
    if (r >= a.n_ids)
        return;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    uint64_t *p = a.hashes + a.out_off[r];
    const uint32_t n = a.n_raw[r];
    uint32_t base = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += nt)
    {
        const uint32_t i = c0 + tid;
        uint64_t key = 0;
        bool keep = false;
        if (i < n)
        {
            key = p[i];
            keep = scaling_keep(key, a.scaling, a.scaling_limit);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            s_scan[wid] = __popc(bal);
        __syncthreads(); // all reads of this chunk are done before anyone writes below it
        uint32_t before = 0, all = 0;
        for (int q = 0; q < (nt >> 5); ++q)
        {
            const uint32_t c = s_scan[q];
            if (q < wid)
                before += c;
            all += c;
        }
        if (keep)
            p[base + before + __popc(bal & ((1u << lane) - 1u))] = key;
        base += all;
        __syncthreads();
    }
    if (tid == 0)
        a.hash_count[r] = base;
}

// -----------------------------------------------------------------------------------------------------------
// build side: the std::set / ankerl set that compute_hashes fills per user bin
// (src/hixf/build/compute_hashes.cpp:76-142), for genomes cut into segments that kernel #1 hashed independently.
// A user bin's raw hashes are spread over many segments, so the set is a global-memory open-addressing table per bin
// that all CTAs insert into (atomicCAS); a second pass compacts the tables.  The FracMin scaling filter is applied
// on insertion.  Order inside a bin is arbitrary: the consumer is XOR-filter construction, a set operation.
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) binset_insert_kernel(BinSetArgs a)
{
    const int lane = threadIdx.x & 31;
    while (true)
    {
        uint32_t sg = 0;
        if (lane == 0)
            sg = atomicAdd(a.work, 1u);
        sg = __shfl_sync(0xffffffffu, sg, 0);
        if (sg >= a.n_segments)
            break;
        const uint32_t b = a.seg_bin[sg];
        uint64_t *tab = a.tables + a.table_off[b];
        const uint32_t mask = (uint32_t)(a.table_off[b + 1] - a.table_off[b]) - 1u;
        const uint64_t *hp = a.hashes + a.out_off[sg];
        const uint32_t n = a.n_raw[sg];
        for (uint32_t i = lane; i < n; i += 32)
        {
            const uint64_t h = hp[i];
            if (!scaling_keep(h, a.scaling, a.scaling_limit))
                continue;
            if (h == kEmptyKey)
                a.bin_has_empty_key[b] = 1u;
            else
                table_insert(tab, mask, h);
        }
    }
}

__global__ void __launch_bounds__(256) binset_compact_kernel(BinSetArgs a)
{
    const uint64_t total = a.table_off[a.n_bins];
    const int lane = threadIdx.x & 31;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base < total; base += (uint64_t)gridDim.x * blockDim.x)
    {
        // a warp looks at 32 consecutive slots; tables are powers of two >= 32 slots, so they belong to one bin
        uint32_t lo = 0, hi = a.n_bins;
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.table_off[mid] <= base)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t b = lo;
        const uint64_t key = a.tables[base + lane];
        const bool keep = key != kEmptyKey;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal)
        {
            uint32_t at = 0;
            if (lane == 0)
                at = atomicAdd(&a.bin_count[b], (uint32_t)__popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (keep)
                a.out[a.out_bin_off[b] + at + __popc(bal & ((1u << lane) - 1u))] = key;
        }
    }
}

// the sentinel key itself, if a bin saw it (one thread per bin)
__global__ void binset_sentinel_kernel(BinSetArgs a)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < a.n_bins && a.bin_has_empty_key[b])
        a.out[a.out_bin_off[b] + atomicAdd(&a.bin_count[b], 1u)] = kEmptyKey;
}

// -----------------------------------------------------------------------------------------------------------
// host-side launchers
// -----------------------------------------------------------------------------------------------------------
template <int K, int S, int T>
static bool try_launch_syncmer(const HashArgs &a, int k, int s, int t, int grid, cudaStream_t st)
{
    if (k != K || s != S || t != T)
        return false;
    const size_t smem = a.fuse_dedup ? (size_t)kHashWarps * kWarpSlots * 4 : 0;
    if (a.min_blocks >= 5 && !smem)
    {
        syncmer_kernel<K, S, T, 5><<<grid, 32 * kHashWarps, 0, st>>>(a);
        return true;
    }
    if (smem) // static + dynamic shared memory exceed 48 KB; the attribute is per device, so it is set on every launch
        cudaFuncSetAttribute(syncmer_kernel<K, S, T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    syncmer_kernel<K, S, T, 4><<<grid, 32 * kHashWarps, smem, st>>>(a);
    return true;
}

// CTAs per SM of the persistent hash / dedup grids come with the arguments (HashArgs / DedupArgs::ctas_per_sm): the
// defaults fill the SM; the engine lowers them when these ALU-bound kernels share SMs with the probe kernels ("overlap").
static int clamp_ctas(int v, int dflt) { return v <= 0 ? dflt : v > 16 ? 16 : v; }

// the (k, s, t) triples with a templated kernel (the only ones that can build the distinct set while hashing)
bool syncmer_has_fast_kernel(int k, int s, int t)
{
    static const int list[][3] = {{22, 12, 5}, {20, 10, 5}, {24, 12, 6}, {26, 14, 6}, {28, 14, 7}, {30, 16, 7}, {18, 10, 4}, {16, 8, 4}};
    for (auto &e : list)
        if (e[0] == k && e[1] == s && e[2] == t)
            return true;
    return false;
}

cudaError_t launch_syncmer(const HashArgs &a, int sm_count, cudaStream_t st)
{
    const int grid = sm_count * clamp_ctas(a.ctas_per_sm, 8);
    bool ok = try_launch_syncmer<22, 12, 5>(a, a.k, a.s, a.t, grid, st)      // Taxor's published indexes
              || try_launch_syncmer<20, 10, 5>(a, a.k, a.s, a.t, grid, st)   // taxor build defaults
              || try_launch_syncmer<24, 12, 6>(a, a.k, a.s, a.t, grid, st)
              || try_launch_syncmer<26, 14, 6>(a, a.k, a.s, a.t, grid, st)
              || try_launch_syncmer<28, 14, 7>(a, a.k, a.s, a.t, grid, st)
              || try_launch_syncmer<30, 16, 7>(a, a.k, a.s, a.t, grid, st)
              || try_launch_syncmer<18, 10, 4>(a, a.k, a.s, a.t, grid, st)
              || try_launch_syncmer<16, 8, 4>(a, a.k, a.s, a.t, grid, st);
    if (!ok)
        syncmer_generic_kernel<<<(a.n_reads + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_syncmer_generic(const HashArgs &a, cudaStream_t st)
{
    syncmer_generic_kernel<<<(a.n_reads + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_kmer(const HashArgs &a, int sm_count, cudaStream_t st)
{
    kmer_kernel<<<sm_count * 4, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_minimiser(const HashArgs &a, int sm_count, cudaStream_t st)
{
    minimiser_kernel<<<sm_count * 4, 32 * kMinWarps, 0, st>>>(a);
    return cudaGetLastError();
}

// capacity <= 2048: warp-per-read kernel; reads it defers (more than 2/3 * 2048 raw hashes) go to `deferred`
cudaError_t launch_dedup_warp(const DedupArgs &a, int sm_count, uint32_t *work_counter, uint32_t *deferred, uint32_t *n_deferred,
                              cudaStream_t st)
{
    if (a.n_ids == 0)
        return cudaSuccess;
    const int ctas = clamp_ctas(a.ctas_per_sm, 6);
    const unsigned grid = a.smf.mod ? (unsigned)(sm_count * ctas)
                                     : (unsigned)std::min<uint64_t>((uint64_t)sm_count * ctas, ((uint64_t)a.n_ids + kDedupWarps - 1) / kDedupWarps);
    dedup_warp_kernel<<<grid, 32 * kDedupWarps, 0, st>>>(a, work_counter, deferred, n_deferred);
    return cudaGetLastError();
}

// CTA-per-read kernel over an explicit list whose length is only known on the device (the deferred reads)
__global__ void dedup_deferred_kernel(DedupArgs a, const uint32_t *n_deferred)
{
    extern __shared__ uint64_t s_tab[];
    __shared__ uint32_t s_scan[32];
    __shared__ int s_saw_empty;
    const uint32_t n_ids = *n_deferred;
    for (uint32_t id = blockIdx.x; id < n_ids; id += gridDim.x)
    {
        const uint32_t r = a.read_ids[id];
        for (int i = threadIdx.x; i < 4096; i += blockDim.x)
            s_tab[i] = kEmptyKey;
        if (threadIdx.x == 0)
            s_saw_empty = 0;
        __syncthreads();
        uint64_t *p = a.hashes + a.out_off[r];
        const uint32_t n = a.n_raw[r];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint64_t h = p[i];
            if (h == kEmptyKey)
                s_saw_empty = 1;
            else
                table_insert(s_tab, 4095u, h);
        }
        __syncthreads();
        const uint32_t cnt = table_compact(s_tab, 4096, p, s_saw_empty != 0, a.scaling, a.scaling_limit, s_scan);
        if (threadIdx.x == 0)
            a.hash_count[r] = cnt;
        __syncthreads();
    }
}

cudaError_t launch_dedup_deferred(const DedupArgs &a, int sm_count, const uint32_t *n_deferred, cudaStream_t st)
{
    dedup_deferred_kernel<<<sm_count * 2, 128, 4096 * 8, st>>>(a, n_deferred);
    return cudaGetLastError();
}

cudaError_t launch_dedup_medium(const DedupArgs &a, cudaStream_t st) // capacity <= 8192
{
    if (a.n_ids == 0)
        return cudaSuccess;
    cudaFuncSetAttribute(dedup_smem_kernel<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    dedup_smem_kernel<16384><<<a.n_ids, 256, 16384 * 8, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_dedup_global(const DedupArgs &a, cudaStream_t st)
{
    if (a.n_ids == 0)
        return cudaSuccess;
    dedup_global_kernel<<<a.n_ids, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_filter(const DedupArgs &a, cudaStream_t st)
{
    if (a.n_ids == 0)
        return cudaSuccess;
    filter_kernel<<<a.n_ids, 128, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_binset(const BinSetArgs &a, int sm_count, cudaStream_t st)
{
    binset_insert_kernel<<<sm_count * 8, 256, 0, st>>>(a);
    binset_compact_kernel<<<sm_count * 8, 256, 0, st>>>(a);
    binset_sentinel_kernel<<<(a.n_bins + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}
} // namespace txr
