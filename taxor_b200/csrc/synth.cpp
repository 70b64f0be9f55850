// synth.cpp -- CPU tooling (libtaxor_tools.so): synthetic genomes and reads, a hierarchical layout and the
// XOR-filter construction that turn per-user-bin hash sets into a valid HIXF.  NOT part of the search path:
// index construction stays on the CPU in the reference as well (src/hixf/build/*, out of scope per SURVEY 8);
// this is the generator for tests and benchmarks.  It obeys the structural invariants of
// IXF 0 is the root; next_ixf_id[i][b] == i for non-merged bins; merged bins carry -1 as user bin
// and store the union of their subtree (construct_ixf.cpp:83-98); a split user bin occupies consecutive bins
// with even chunks of its hashes (hierarchical_build.cpp:91-110); all bins of an IXF share one capacity and one
// seed, and a failed peel re-seeds the whole IXF (construct_ixf.cpp:101-108).
// The filter arithmetic comes from ixf_arith.cuh (parity unpinned, see there); the peeling itself is the
// standard 3-wise XOR-filter construction (Graf & Lemire), as in src/main/xorfilter.hpp:140-334.
#include "ixf_arith.cuh"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <numeric>
#include <parallel/algorithm>
#include <queue>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
inline uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
inline uint64_t packed_words(uint64_t n) { return (n + 31) / 32 + 1; }
inline unsigned base_at(const uint64_t *w, uint64_t i) { return (unsigned)((w[i >> 5] >> (62 - 2 * (i & 31))) & 3); }

struct BuiltIxf
{
    uint64_t seed{}, bins{}, tbins{}, seg_len{}, rows{}, count_len{}, capacity{};
    std::vector<uint8_t> data;
    std::vector<int64_t> next, ub;
};

struct Hixf
{
    std::deque<BuiltIxf> ixf; // a deque: references stay valid while other threads append
    std::mutex ixf_mutex;
    // flattened accessors
    std::vector<uint64_t> seed, bins, tbins, seg_len, bin_off, rows, capacity;
    txr::IxfScheme scheme{txr::kIxfSlotsXor3, txr::kIxfMixAddSeed, txr::kIxfFpFold32, 21u, 42u};
    std::vector<const uint8_t *> data;
    std::vector<int64_t> next, ub;
    uint64_t n_user_bins{};
    uint64_t reseeds{};
};

// peel one bin; fp (3*seg_len bytes) receives the fingerprints.  false: not peelable with this seed
bool peel_bin(const uint64_t *keys, size_t n, uint64_t seed, const txr::IxfScheme &sch, uint32_t seg_len, uint32_t count_len, size_t slots,
              uint8_t *fp, std::vector<uint32_t> &cnt, std::vector<uint64_t> &xr, std::vector<uint32_t> &queue,
              std::vector<uint64_t> &stack_h, std::vector<uint32_t> &stack_s)
{
    std::memset(fp, 0, slots);
    if (n == 0)
        return true;
    cnt.assign(slots, 0);
    xr.assign(slots, 0);
    for (size_t i = 0; i < n; ++i)
    {
        const uint64_t h = txr::ixf_mix_g(keys[i], seed, sch);
        uint32_t p[3];
        txr::ixf_slots_g(h, seg_len, count_len, sch, p[0], p[1], p[2]);
        for (int j = 0; j < 3; ++j)
        {
            ++cnt[p[j]];
            xr[p[j]] ^= h;
        }
    }
    queue.clear();
    for (size_t s = 0; s < slots; ++s)
        if (cnt[s] == 1)
            queue.push_back((uint32_t)s);
    stack_h.clear();
    stack_s.clear();
    while (!queue.empty())
    {
        const uint32_t s = queue.back();
        queue.pop_back();
        if (cnt[s] != 1)
            continue;
        const uint64_t h = xr[s];
        stack_h.push_back(h);
        stack_s.push_back(s);
        uint32_t p[3];
        txr::ixf_slots_g(h, seg_len, count_len, sch, p[0], p[1], p[2]);
        for (int j = 0; j < 3; ++j)
        {
            --cnt[p[j]];
            xr[p[j]] ^= h;
            if (cnt[p[j]] == 1)
                queue.push_back(p[j]);
        }
    }
    if (stack_h.size() != n)
        return false;
    for (size_t i = n; i-- > 0;)
    {
        const uint64_t h = stack_h[i];
        uint32_t p[3];
        txr::ixf_slots_g(h, seg_len, count_len, sch, p[0], p[1], p[2]);
        uint8_t f = (uint8_t)txr::ixf_fingerprint_g(h, sch);
        for (int j = 0; j < 3; ++j)
            if (p[j] != stack_s[i])
                f ^= fp[p[j]];
        fp[stack_s[i]] = f;
    }
    return true;
}

struct BinSpec
{
    int64_t ub{-1};             // user bin, or -1 for a merged bin
    uint32_t split_idx{0}, split_n{1};
    std::vector<uint32_t> group; // merged: the user bins below
};

struct Builder
{
    const uint64_t *const *ub_hashes;
    const uint64_t *ub_n;
    uint32_t t_max;
    uint64_t seed_state;
    Hixf *out;
    uint32_t t_max_lower{0}; // != 0: bin budget of the IXFs below the root (a wide root over narrow lower levels)

    // returns the index of the IXF built for `ubs` (sorted by size, descending); `all` receives the sorted
    // distinct union of the subtree when want_union
    size_t build(const std::vector<uint32_t> &ubs, bool want_union, std::vector<uint64_t> &all, int depth = 0)
    {
        size_t my;
        {
            std::lock_guard<std::mutex> g(out->ixf_mutex);
            my = out->ixf.size();
            out->ixf.emplace_back();
        }
        const uint32_t t_max = depth > 0 && t_max_lower ? t_max_lower : this->t_max;
        const size_t n = ubs.size();
        std::vector<BinSpec> spec;
        if (n <= t_max)
        {
            // every user bin gets a bin; spare bins split the heaviest ones (never below 64 hashes per part)
            std::vector<uint32_t> parts(n, 1);
            using Item = std::pair<double, uint32_t>;
            std::priority_queue<Item> pq;
            for (uint32_t i = 0; i < n; ++i)
                pq.emplace((double)ub_n[ubs[i]], i);
            for (size_t spare = t_max - n; spare > 0 && !pq.empty(); --spare)
            {
                auto [load, i] = pq.top();
                pq.pop();
                if (load / 2 < 64)
                    break;
                ++parts[i];
                pq.emplace((double)ub_n[ubs[i]] / parts[i], i);
            }
            for (uint32_t i = 0; i < n; ++i)
                for (uint32_t q = 0; q < parts[i]; ++q)
                {
                    BinSpec b;
                    b.ub = ubs[i];
                    b.split_idx = q;
                    b.split_n = parts[i];
                    spec.push_back(std::move(b));
                }
        }
        else
        {
            uint64_t total = 0;
            for (auto u : ubs)
                total += ub_n[u];
            // user bins at least as heavy as an average bin keep a bin of their own, the rest is packed into
            // merged bins of roughly equal weight (contiguous in the size-sorted order)
            size_t singles = 0;
            uint64_t rest = total;
            while (singles < n && singles + 1 < t_max && (double)ub_n[ubs[singles]] >= (double)rest / (double)(t_max - singles))
                rest -= ub_n[ubs[singles++]];
            for (size_t i = 0; i < singles; ++i)
            {
                BinSpec b;
                b.ub = ubs[i];
                spec.push_back(std::move(b));
            }
            size_t groups = t_max - singles, i = singles;
            uint64_t remaining = rest;
            for (size_t g = 0; g < groups && i < n; ++g)
            {
                const size_t groups_left = groups - g;
                const uint64_t target = remaining / groups_left;
                BinSpec b;
                uint64_t acc = 0;
                // leave at least one user bin for every remaining group
                while (i < n && (b.group.empty() || (acc + ub_n[ubs[i]] / 2 <= target && n - i > groups_left - 1) || groups_left == 1))
                {
                    acc += ub_n[ubs[i]];
                    b.group.push_back(ubs[i++]);
                }
                remaining -= acc;
                if (b.group.size() == 1)
                {
                    b.ub = b.group[0];
                    b.group.clear();
                }
                spec.push_back(std::move(b));
            }
        }
        const size_t bins = spec.size();
        // key sets per technical bin
        std::vector<std::vector<uint64_t>> owned(bins);
        std::vector<const uint64_t *> kptr(bins);
        std::vector<size_t> kn(bins);
        std::vector<int64_t> next(bins), ubv(bins);
        // Children of a wide IXF (hundreds of merged bins) are built by different threads: each subtree's peels and
        // sorts are small, so parallelism has to come from the subtrees.  IXF numbers then depend on the schedule, as in
        // the reference (atomic counter, build_data.hpp:34-37); narrow IXFs keep the serial, reproducible order.
        size_t n_merged = 0;
        for (size_t b = 0; b < bins; ++b)
            n_merged += spec[b].ub < 0;
        auto fill_bin = [&](size_t b)
        {
            if (spec[b].ub >= 0)
            {
                const uint64_t u = (uint64_t)spec[b].ub, cnt = ub_n[u];
                const uint64_t lo = cnt * spec[b].split_idx / spec[b].split_n, hi = cnt * (spec[b].split_idx + 1) / spec[b].split_n;
                kptr[b] = ub_hashes[u] + lo;
                kn[b] = hi - lo;
                next[b] = (int64_t)my;
                ubv[b] = (int64_t)u;
            }
            else
            {
                const size_t child = build(spec[b].group, true, owned[b], depth + 1);
                kptr[b] = owned[b].data();
                kn[b] = owned[b].size();
                next[b] = (int64_t)child;
                ubv[b] = -1;
            }
        };
#ifdef _OPENMP
        const bool par_children = n_merged >= 128 && !omp_in_parallel() && omp_get_max_threads() > 1;
#else
        const bool par_children = false;
#endif
        if (par_children)
        {
#pragma omp parallel for schedule(dynamic, 1)
            for (long b = 0; b < (long)bins; ++b)
                fill_bin((size_t)b);
        }
        else
            for (size_t b = 0; b < bins; ++b)
                fill_bin(b);
        BuiltIxf &x = out->ixf[my];
        size_t max_n = 0;
        for (size_t b = 0; b < bins; ++b)
            max_n = std::max(max_n, kn[b]);
        x.bins = bins;
        x.tbins = (bins + 63) / 64 * 64;
        // capacity: the reference sizes an IXF from the layout's (HyperLogLog, i.e. approximate) max_bin_hashes
        // (construct_ixf.cpp:58).  1.23*n slots is marginal for peeling when all bins of an IXF must succeed with
        // ONE seed, so the generator adds 6 % headroom; otherwise a balanced 64-bin IXF needs ~1000 re-seeds.
        // With thousands of small bins (wide test indexes) every one of them has to peel under the same seed, and a small
        // 3-wise XOR system fails far more often than a large one: more headroom per doubling of the bin count beyond 64.
        size_t wide_extra = 0;
        for (size_t bb = bins; bb > 64 && max_n < (1u << 16); bb >>= 1)
            wide_extra += max_n / 24 + 16;
        const txr::IxfScheme &sch = out->scheme;
        // binary fuse filters come with their own size factor (>= 1.125): no extra headroom beyond the wide-index one
        x.capacity = sch.slots == txr::kIxfSlotsFuse3 ? max_n + wide_extra + (max_n < 4096 ? max_n / 8 + 16 : 0) : max_n + max_n / 16 + 32 + wide_extra;
        const txr::IxfGeometry geo = txr::ixf_geometry_for(sch, x.capacity);
        x.seg_len = geo.seg_len;
        x.rows = geo.rows;
        x.count_len = geo.count_len;
        x.seed = 13572355802537770549ULL; // default seed of the prototype (xorfilter.hpp:153)
        x.next = next;
        x.ub = ubv;
        const size_t slots = x.rows;
        x.data.assign(slots * x.tbins, 0);
        std::vector<uint8_t> cols(slots * bins); // bin-major scratch: cols[b * slots + slot]
        while (true)
        {
            int failed = 0;
#pragma omp parallel
            {
                std::vector<uint32_t> cnt, queue, stack_s;
                std::vector<uint64_t> xr, stack_h;
#pragma omp for schedule(dynamic, 1)
                for (long b = 0; b < (long)bins; ++b)
                {
                    if (failed)
                        continue;
                    if (!peel_bin(kptr[b], kn[b], x.seed, sch, (uint32_t)x.seg_len, (uint32_t)x.count_len, slots,
                                  cols.data() + (size_t)b * slots, cnt, xr, queue, stack_h, stack_s))
                    {
#pragma omp atomic write
                        failed = 1;
                    }
                }
            }
            if (!failed)
                break;
            // construct_ixf.cpp:101-108: clear the whole IXF and draw a new seed
            {
                std::lock_guard<std::mutex> g(out->ixf_mutex);
                x.seed = splitmix64(seed_state);
                ++out->reseeds;
            }
        }
        // interleave: data[slot * tbins + bin]; every thread owns a block of rows, so no cache line is shared
        {
            const size_t block = 2048;
#pragma omp parallel for schedule(static)
            for (long s0 = 0; s0 < (long)slots; s0 += (long)block)
            {
                const size_t s1 = std::min(slots, (size_t)s0 + block);
                for (size_t b = 0; b < bins; ++b)
                {
                    const uint8_t *src = cols.data() + b * slots;
                    uint8_t *dst = x.data.data() + b;
                    for (size_t sl = (size_t)s0; sl < s1; ++sl)
                        dst[sl * x.tbins] = src[sl];
                }
            }
        }
        if (want_union)
        {
            size_t total = 0;
            for (size_t b = 0; b < bins; ++b)
                total += kn[b];
            all.clear();
            all.reserve(total);
            for (size_t b = 0; b < bins; ++b)
                all.insert(all.end(), kptr[b], kptr[b] + kn[b]);
            if (all.size() > (1u << 20))
                __gnu_parallel::sort(all.begin(), all.end());
            else
                std::sort(all.begin(), all.end());
            all.erase(std::unique(all.begin(), all.end()), all.end());
        }
        return my;
    }
};
} // namespace

extern "C" {

uint64_t txs_packed_words(uint64_t n_bases) { return packed_words(n_bases); }

// uniform i.i.d. genome, 2-bit packed in the library's read layout (MSB-first, one zero pad word)
void txs_genome(uint64_t seed, uint64_t len, uint64_t *words)
{
    uint64_t st = seed ^ 0x5851f42d4c957f2dULL; // scramble: consecutive seeds must not give shifted copies
    st = splitmix64(st) ^ (seed << 32);
    const uint64_t full = len / 32;
    for (uint64_t w = 0; w < full; ++w)
        words[w] = splitmix64(st);
    const uint64_t nw = packed_words(len);
    if (full + 1 < nw)
    {
        const unsigned rem = (unsigned)(len % 32);
        words[full] = splitmix64(st) & ~((1ULL << (64 - 2 * rem)) - 1);
    }
    words[nw - 1] = 0;
}

// ONT-like reads: genome, start and strand uniform; per-base errors at rate `err` split evenly into
// substitutions (to a different base), insertions (uniform base) and deletions.  Read i depends only on
// (seed, i).  word_off[i] must leave txs_packed_words(read_len[i]) words per read.
int txs_reads(const uint64_t *const *genome_words, const uint64_t *genome_len, uint64_t n_genomes, uint64_t n_reads,
              const uint32_t *read_len, double err, uint64_t seed, uint64_t *words, const uint64_t *word_off,
              uint32_t *out_genome, int threads)
{
#ifdef _OPENMP
    if (threads > 0)
        omp_set_num_threads(threads);
#endif
    int bad = 0;
    const uint64_t sub_t = (uint64_t)(err / 3.0 * 18446744073709551615.0);
#pragma omp parallel for schedule(dynamic, 256)
    for (long long r = 0; r < (long long)n_reads; ++r)
    {
        uint64_t st = seed ^ (0xd1b54a32d192ed03ULL * (uint64_t)(r + 1));
        st = splitmix64(st) ^ (uint64_t)r; // decorrelate neighbouring reads
        const uint32_t L = read_len[r];
        const uint64_t g = splitmix64(st) % n_genomes;
        const uint64_t G = genome_len[g];
        const uint64_t span = (uint64_t)L + L / 8 + 64; // source window incl. room for deletions
        if (G < span)
        {
            bad = 1;
            continue;
        }
        const uint64_t start = splitmix64(st) % (G - span + 1);
        const bool rev = splitmix64(st) & 1;
        const uint64_t *gw = genome_words[g];
        // built in a thread-local buffer and copied out once: the destination may be write-combined pinned memory
        // (TXR_HOST_WC=1), where a read-modify-write per base would be an uncached read per base
        uint64_t *out_words = words + word_off[r];
        const uint64_t nw = packed_words(L);
        thread_local std::vector<uint64_t> local;
        local.assign(nw, 0);
        uint64_t *dst = local.data();
        uint64_t src = 0; // offset inside the source window (in read orientation)
        uint32_t produced = 0;
        while (produced < L && src < span)
        {
            const uint64_t e = splitmix64(st);
            unsigned b;
            if (e < sub_t) // substitution
            {
                const unsigned o = rev ? 3 - base_at(gw, start + span - 1 - src) : base_at(gw, start + src);
                b = (o + 1 + (unsigned)(splitmix64(st) % 3)) & 3;
                ++src;
            }
            else if (e < 2 * sub_t) // insertion
                b = (unsigned)(splitmix64(st) & 3);
            else if (e < 3 * sub_t) // deletion
            {
                ++src;
                continue;
            }
            else
            {
                b = rev ? 3 - base_at(gw, start + span - 1 - src) : base_at(gw, start + src);
                ++src;
            }
            dst[produced >> 5] |= (uint64_t)b << (62 - 2 * (produced & 31));
            ++produced;
        }
        std::memcpy(out_words, dst, nw * 8);
        if (produced < L)
            bad = 1;
        if (out_genome)
            out_genome[r] = (uint32_t)g;
    }
    return bad ? -1 : 0;
}

// sorts and de-duplicates n_ub key arrays in place (parallel over arrays); counts[i] is updated
void txs_sort_unique_many(uint64_t *const *keys, uint64_t *counts, uint64_t n_ub, int threads)
{
#ifdef _OPENMP
    if (threads > 0)
        omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < (long)n_ub; ++i)
    {
        std::sort(keys[i], keys[i] + counts[i]);
        counts[i] = (uint64_t)(std::unique(keys[i], keys[i] + counts[i]) - keys[i]);
    }
}

void *txs_hixf_build3(const uint64_t *const *ub_hashes, const uint64_t *ub_n, uint64_t n_ub, uint32_t t_max, uint32_t t_max_lower,
                      uint64_t seed, int threads, const uint32_t *scheme5);
void *txs_hixf_build2(const uint64_t *const *ub_hashes, const uint64_t *ub_n, uint64_t n_ub, uint32_t t_max, uint32_t t_max_lower,
                      uint64_t seed, int threads)
{
    return txs_hixf_build3(ub_hashes, ub_n, n_ub, t_max, t_max_lower, seed, threads, nullptr);
}
void *txs_hixf_build(const uint64_t *const *ub_hashes, const uint64_t *ub_n, uint64_t n_ub, uint32_t t_max, uint64_t seed,
                     int threads)
{
    return txs_hixf_build2(ub_hashes, ub_n, n_ub, t_max, 0, seed, threads);
}

// t_max_lower != 0: the IXFs below the root get that bin budget instead of t_max (GTDB-shaped test/bench indexes: a
// 4096-bin root over narrow lower levels, several levels deep without needing millions of user bins)
// scheme5 (may be null = the prototype's arithmetic): {slots, mix, fingerprint, rot1, rot2} of txr_ixf_scheme / txr::IxfScheme
void *txs_hixf_build3(const uint64_t *const *ub_hashes, const uint64_t *ub_n, uint64_t n_ub, uint32_t t_max, uint32_t t_max_lower,
                      uint64_t seed, int threads, const uint32_t *scheme5)
{
#ifdef _OPENMP
    if (threads > 0)
        omp_set_num_threads(threads);
#endif
    if (n_ub == 0 || t_max < 2 || t_max_lower == 1)
        return nullptr;
    auto h = std::make_unique<Hixf>();
    h->n_user_bins = n_ub;
    if (scheme5)
    {
        h->scheme = txr::IxfScheme{scheme5[0], scheme5[1], scheme5[2], scheme5[3], scheme5[4]};
        if (h->scheme.rot1 == 0 && h->scheme.rot2 == 0)
        {
            h->scheme.rot1 = 21;
            h->scheme.rot2 = 42;
        }
        if (!txr::ixf_scheme_valid(h->scheme))
            return nullptr;
    }
    std::vector<uint32_t> ubs(n_ub);
    std::iota(ubs.begin(), ubs.end(), 0u);
    std::stable_sort(ubs.begin(), ubs.end(), [&](uint32_t a, uint32_t b) { return ub_n[a] > ub_n[b]; });
    Builder bld{ub_hashes, ub_n, t_max, seed, h.get(), t_max_lower};
    std::vector<uint64_t> unused;
    bld.build(ubs, false, unused);
    h->bin_off.push_back(0);
    for (auto &x : h->ixf)
    {
        h->seed.push_back(x.seed);
        h->bins.push_back(x.bins);
        h->tbins.push_back(x.tbins);
        h->seg_len.push_back(x.seg_len);
        h->rows.push_back(x.rows);
        h->capacity.push_back(x.capacity);
        h->data.push_back(x.data.data());
        h->next.insert(h->next.end(), x.next.begin(), x.next.end());
        h->ub.insert(h->ub.end(), x.ub.begin(), x.ub.end());
        h->bin_off.push_back(h->next.size());
    }
    return h.release();
}

void txs_hixf_free(void *p) { delete static_cast<Hixf *>(p); }
uint64_t txs_hixf_n_ixf(void *p) { return static_cast<Hixf *>(p)->ixf.size(); }
uint64_t txs_hixf_reseeds(void *p) { return static_cast<Hixf *>(p)->reseeds; }
const uint64_t *txs_hixf_rows(void *p) { return static_cast<Hixf *>(p)->rows.data(); }
const uint64_t *txs_hixf_capacity(void *p) { return static_cast<Hixf *>(p)->capacity.data(); } // the max_elems an IXF was sized for
void txs_hixf_arrays(void *p, const uint64_t **seed, const uint64_t **bins, const uint64_t **tbins, const uint64_t **seg_len,
                     const uint8_t *const **data, const uint64_t **bin_off, const int64_t **next, const int64_t **ub)
{
    Hixf *h = static_cast<Hixf *>(p);
    *seed = h->seed.data();
    *bins = h->bins.data();
    *tbins = h->tbins.data();
    *seg_len = h->seg_len.data();
    *data = h->data.data();
    *bin_off = h->bin_off.data();
    *next = h->next.data();
    *ub = h->ub.data();
}

} // extern "C"
