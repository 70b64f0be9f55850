/*
 * taxor_oracle.cpp -- CPU ORACLE (restatement) of the `taxor search` hot path.  TEST INFRASTRUCTURE ONLY.
 * See taxor_oracle.h for the parity status ("parity unpinned" for the three third-party pieces).
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 * Nothing in taxor_b200/ may include, link or call this file.
 */
#include "taxor_oracle.h"
#include "ixf_ref.h"
#if defined(__AVX512BW__) || defined(__AVX2__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * scalar pieces
 * ---------------------------------------------------------------------------------------------- */

/* ankerl::unordered_dense v3.0.1 (pinned at src/hashing/CMakeLists.txt.in:9), detail::wyhash::hash(uint64_t):
 * mix(x, 0x9E3779B97F4A7C15) with mix(a,b) = lo64(a*b) ^ hi64(a*b).  Call site src/hashing/syncmer.cpp:73-77. */
extern "C" uint64_t orc_wyhash_u64(uint64_t x)
{
    __uint128_t r = (__uint128_t)x * (__uint128_t)UINT64_C(0x9E3779B97F4A7C15);
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}

/* src/hixf/build/adjust_seed.hpp:40-44 */
extern "C" uint64_t orc_adjust_seed(uint8_t kmer_size)
{
    return UINT64_C(0x8F3F73B5CF1C9ADE) >> (64u - 2u * kmer_size);
}

/* seqan3::dna4 char_to_rank as used by dna4_traits (src/hixf/build/dna4_traits.hpp:15-18):
 * every IUPAC symbol is collapsed at parse time, case-insensitively; anything else is illegal. */
extern "C" int orc_dna4_rank(unsigned char c)
{
    switch (c >= 'a' && c <= 'z' ? c - 32 : c)
    {
    case 'A': case 'N': case 'R': case 'W': case 'M': case 'D': case 'H': case 'V': return 0;
    case 'C': case 'Y': case 'S': case 'B': return 1;
    case 'G': case 'K': return 2;
    case 'T': case 'U': return 3;
    default: return -1;
    }
}

/* src/main/taxor_build.cpp:510: ceil((k - s + 1) / 2) with INTEGER division inside the ceil */
extern "C" int orc_t_syncmer(int k, int s)
{
    return (int)std::ceil((double)((k - s + 1) / 2));
}

/* src/main/taxor_search.cpp:223-233 / 244-251 */
extern "C" int orc_scaling_keep(uint64_t h, uint16_t scaling)
{
    uint64_t v = orc_wyhash_u64(h);
    return double(v) <= double(UINT64_MAX) / double(scaling);
}

/* ------------------------------------------------------------------------------------------------
 * hashing
 * ---------------------------------------------------------------------------------------------- */

namespace
{
/* The scan of src/hashing/syncmer.cpp:80-155, with the std::deque replaced by a ring of the last
 * k-s+1 canonical s-mers.  `emit(hash)` is called for every selected k-mer, in scan order. */
template <typename emit_t>
void syncmer_scan(const uint8_t *codes, int64_t len, size_t k, size_t s, size_t t, emit_t &&emit)
{
    const uint64_t kmask = (k >= 32) ? ~UINT64_C(0) : ((UINT64_C(1) << 2 * k) - 1); /* :86 */
    const uint64_t smask = (UINT64_C(1) << 2 * s) - 1;                             /* :87 */
    const uint64_t kshift = (k - 1) * 2, sshift = (s - 1) * 2;                     /* :88-89 */
    const size_t w = k - s + 1;                                                    /* queue capacity */
    std::vector<uint64_t> q(w + 1);                                                /* ring: logical index 0 = oldest */
    size_t q_head = 0, q_size = 0;
    auto q_at = [&](size_t j) -> uint64_t { return q[(q_head + j) % (w + 1)]; };
    uint64_t min_val = UINT64_MAX;                                                 /* :91 */
    size_t min_pos = (size_t)-1;                                                   /* :92 */
    size_t l = 0;
    uint64_t xk0 = 0, xk1 = 0, xs0 = 0, xs1 = 0;
    for (int64_t ii = 0; ii < len; ++ii)
    {
        const size_t i = (size_t)ii;
        const int c = codes[ii];
        if (c < 4)
        {
            xk0 = (xk0 << 2 | (uint64_t)c) & kmask;                                /* :101 */
            xk1 = xk1 >> 2 | (uint64_t)(3 - c) << kshift;                          /* :102 */
            xs0 = (xs0 << 2 | (uint64_t)c) & smask;                                /* :103 */
            xs1 = xs1 >> 2 | (uint64_t)(3 - c) << sshift;                          /* :104 */
            if (++l < s)                                                           /* :105 */
                continue;
            const uint64_t hash_s = std::min(xs0, xs1);                            /* :109 */
            q[(q_head + q_size) % (w + 1)] = hash_s;                               /* push_back :111 */
            ++q_size;
            if (q_size < w)                                                        /* :113 */
                continue;
            if (q_size == w)                                                       /* first k-mer: leftmost min :116-123 */
            {
                for (size_t j = 0; j < q_size; ++j)
                    if (q_at(j) < min_val)
                    {
                        min_val = q_at(j);
                        min_pos = i - k + j + 1;
                    }
            }
            else
            {
                q_head = (q_head + 1) % (w + 1);                                   /* pop_front :127 */
                --q_size;
                if (min_pos == i - k)                                              /* min left: rightmost min :128-136 */
                {
                    min_val = UINT64_MAX;
                    min_pos = i - s + 1;
                    for (int64_t j = (int64_t)q_size - 1; j >= 0; --j)
                        if (q_at((size_t)j) < min_val)
                        {
                            min_val = q_at((size_t)j);
                            min_pos = i - k + (size_t)j + 1;
                        }
                }
                else if (hash_s < min_val)                                         /* strictly smaller arrival :137-140 */
                {
                    min_val = hash_s;
                    min_pos = i - s + 1;
                }
            }
            if (min_pos == i - k + t)                                              /* :142 */
                emit(orc_wyhash_u64(std::min(xk0, xk1)));                          /* :144-145 */
        }
        else                                                                       /* "N": restart :147-153 */
        {
            min_val = UINT64_MAX;
            min_pos = (size_t)-1;
            l = 0;
            xs0 = xs1 = xk0 = xk1 = 0;
            q_head = q_size = 0;
        }
    }
}
} // namespace

extern "C" int64_t orc_syncmer_hashes(const uint8_t *codes, int64_t len, int k, int s, int t,
                                      uint64_t *out, int64_t cap)
{
    /* ankerl::unordered_dense::set<size_t> (syncmer.cpp:145,157-165): distinct, iterates in insertion order */
    std::unordered_set<uint64_t> seen;
    int64_t n = 0;
    syncmer_scan(codes, len, (size_t)k, (size_t)s, (size_t)t,
                 [&](uint64_t h)
                 {
                     if (seen.insert(h).second)
                     {
                         if (n < cap)
                             out[n] = h;
                         ++n;
                     }
                 });
    return n <= cap ? n : -n;
}

extern "C" int64_t orc_syncmer_hashes_raw(const uint8_t *codes, int64_t len, int k, int s, int t,
                                          uint64_t *out, int64_t cap)
{
    int64_t n = 0;
    syncmer_scan(codes, len, (size_t)k, (size_t)s, (size_t)t,
                 [&](uint64_t h)
                 {
                     if (n < cap)
                         out[n] = h;
                     ++n;
                 });
    return n <= cap ? n : -n;
}

/* seq | seqan3::views::minimiser_hash(ungapped k, window_size == k, seed) (taxor_search.cpp:210-212,242):
 * UPSTREAM SeqAn3 3.3.0 semantics (fork unverified -> parity unpinned): forward code = sum rank*4^(k-1-i),
 * reverse strand likewise on the reverse complement, both XOR seed, value = min of the two, one value per
 * k-mer start, duplicates kept, in position order. */
extern "C" int64_t orc_kmer_hashes(const uint8_t *codes, int64_t len, int k, uint64_t seed,
                                   uint64_t *out, int64_t cap)
{
    const uint64_t kmask = (k >= 32) ? ~UINT64_C(0) : ((UINT64_C(1) << 2 * k) - 1);
    const uint64_t kshift = (uint64_t)(k - 1) * 2;
    uint64_t f = 0, r = 0;
    int64_t n = 0;
    for (int64_t i = 0; i < len; ++i)
    {
        const uint64_t c = codes[i] & 3;
        f = (f << 2 | c) & kmask;
        r = r >> 2 | (3 - c) << kshift;
        if (i + 1 < k)
            continue;
        if (n < cap)
            out[n] = std::min(f ^ seed, r ^ seed);
        ++n;
    }
    return n <= cap ? n : -n;
}

/* seq | seqan3::views::minimiser_hash(ungapped k, window_size w > k, seed) (taxor_search.cpp:210-212,242;
 * build side compute_hashes.cpp:120-124).  UPSTREAM SeqAn3 3.3.0 semantics of views::minimiser over the
 * per-position values min(fwd ^ seed, rc ^ seed) (fork unverified -> parity unpinned):
 *   - W = w - k + 1 values per window; a range with fewer values than W is ONE window (the iterator
 *     constructor clamps window_size to the range size); no values -> no output;
 *   - first window: the RIGHTMOST minimum (min_element with std::less_equal) is reported;
 *   - every shift: if the tracked minimiser was the value that left the window, the window is rescanned
 *     (rightmost minimum again) and the result is reported EVEN IF ITS VALUE IS UNCHANGED; else a new value
 *     STRICTLY smaller than the tracked one becomes the minimiser and is reported; an equal one is ignored.
 * Output = the reported values in order, duplicates kept (taxor_search.cpp:242-255 pushes every one). */
extern "C" int64_t orc_minimiser_hashes(const uint8_t *codes, int64_t len, int k, int w, uint64_t seed,
                                        uint64_t *out, int64_t cap)
{
    const int64_t n = len >= k ? len - k + 1 : 0;
    if (n == 0)
        return 0;
    std::vector<uint64_t> v((size_t)n);
    if (orc_kmer_hashes(codes, len, k, seed, v.data(), n) != n)
        return 0;
    const int64_t W = std::min<int64_t>(std::max(w - k + 1, 1), n);
    int64_t cnt = 0;
    auto report = [&](uint64_t x)
    {
        if (cnt < cap)
            out[cnt] = x;
        ++cnt;
    };
    auto rightmost_min = [&](int64_t first)
    {
        int64_t pos = first;
        for (int64_t j = first + 1; j < first + W; ++j)
            if (v[(size_t)j] <= v[(size_t)pos])
                pos = j;
        return pos;
    };
    int64_t pos = rightmost_min(0);
    report(v[(size_t)pos]);
    for (int64_t first = 1; first + W <= n; ++first)
    {
        const int64_t arriving = first + W - 1;
        if (pos < first)
        {
            pos = rightmost_min(first);
            report(v[(size_t)pos]);
        }
        else if (v[(size_t)arriving] < v[(size_t)pos])
        {
            pos = arriving;
            report(v[(size_t)pos]);
        }
    }
    return cnt <= cap ? cnt : -cnt;
}

/* ------------------------------------------------------------------------------------------------
 * thresholds  (src/hixf/search/)
 * ---------------------------------------------------------------------------------------------- */

namespace
{
/* syncmer_model.hpp:14-36 */
const double matching_ratios[21][10] = {
    {0.552077, 0.195989, 0.151428, 0.118475, 0.0946177, 0.0797244, 0.0604658, 0.0480255, 0.0367569, 0.0252911},
    {0.552385, 0.207533, 0.161204, 0.127368, 0.103704, 0.0881939, 0.0689396, 0.0556991, 0.044185, 0.0298818},
    {0.552239, 0.220393, 0.17382, 0.139866, 0.113736, 0.0966358, 0.0783558, 0.0639223, 0.0523452, 0.0389549},
    {0.552682, 0.236329, 0.188152, 0.152267, 0.126191, 0.106106, 0.0876917, 0.0730642, 0.0621864, 0.0489249},
    {0.553172, 0.254091, 0.202686, 0.165344, 0.137087, 0.116649, 0.098822, 0.0831266, 0.0703342, 0.0582562},
    {0.553716, 0.271183, 0.219848, 0.181959, 0.152163, 0.130048, 0.110622, 0.0942414, 0.0810792, 0.0688187},
    {0.554532, 0.292154, 0.240059, 0.199738, 0.168952, 0.144956, 0.122726, 0.105878, 0.0940805, 0.0777557},
    {0.557957, 0.313553, 0.260912, 0.220014, 0.186567, 0.16101, 0.137399, 0.119867, 0.10453, 0.0900014},
    {0.563925, 0.338316, 0.283689, 0.2401, 0.206963, 0.179541, 0.155347, 0.135128, 0.121575, 0.104741},
    {0.568519, 0.364594, 0.310373, 0.267578, 0.231083, 0.20088, 0.174376, 0.153111, 0.139339, 0.120042},
    {0.579726, 0.395595, 0.338947, 0.295287, 0.258713, 0.22876, 0.200759, 0.175309, 0.161306, 0.139616},
    {0.599258, 0.430241, 0.371291, 0.325596, 0.289651, 0.257329, 0.228011, 0.201799, 0.186956, 0.164794},
    {0.611572, 0.468953, 0.410482, 0.363923, 0.325828, 0.293046, 0.26167, 0.235216, 0.216716, 0.192162},
    {0.624341, 0.510411, 0.452122, 0.407016, 0.370022, 0.334601, 0.303413, 0.275232, 0.254563, 0.227871},
    {0.655724, 0.555245, 0.498564, 0.453201, 0.416285, 0.381883, 0.352291, 0.322556, 0.299739, 0.271481},
    {0.694872, 0.608367, 0.552085, 0.509395, 0.471692, 0.437803, 0.405938, 0.377117, 0.354352, 0.325132},
    {0.742071, 0.669034, 0.613738, 0.57366, 0.539215, 0.50832, 0.476855, 0.449152, 0.42683, 0.397277},
    {0.795543, 0.733694, 0.68341, 0.647737, 0.617382, 0.588448, 0.56083, 0.533714, 0.514757, 0.486399},
    {0.853121, 0.802585, 0.763169, 0.733734, 0.708902, 0.684331, 0.660171, 0.637633, 0.621567, 0.596993},
    {0.918163, 0.882314, 0.854479, 0.835831, 0.819643, 0.804269, 0.788526, 0.771895, 0.763059, 0.742114},
    {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0}};

/* gaussian_inverse.cpp:12-23 */
double rational_approximation(double t)
{
    const double c[] = {2.515517, 0.802853, 0.010328};
    const double d[] = {1.432788, 0.189269, 0.001308};
    return t - ((c[2] * t + c[1]) * t + c[0]) / (((d[2] * t + d[1]) * t + d[0]) * t + 1.0);
}
/* kmer_model.cpp:28-32 */
double expected_nmut_kmer(double r, size_t kmer_size, size_t kmer_count)
{
    double q = 1.0 - pow(1.0 - r, kmer_size);
    return kmer_count * q;
}
/* kmer_model.cpp:34-41 (same expression as in calculate_nmut_kmer_CI :12-16) */
double variance_nmut_kmer(double r, size_t kmer_size, size_t kmer_count)
{
    double q = 1.0 - pow(1.0 - r, kmer_size);
    double varN = (double)kmer_count * (1.0 - q) * (q * (2.0 * (double)kmer_size + (2.0 / r) - 1.0) - 2.0 * (double)kmer_size)
                + (double)kmer_size * ((double)kmer_size - 1.0) * pow((1.0 - q), 2.0)
                + (2.0 * (1.0 - q) / (pow(r, 2.0))) * ((1.0 + ((double)kmer_size - 1.0) * (1.0 - q)) * r - q);
    return varN;
}
/* kmer_model.cpp:43-46 */
double expected_nmut_kmer_squared(double r, size_t kmer_size, size_t kmer_count)
{
    return pow(expected_nmut_kmer(r, kmer_size, kmer_count), 2) + variance_nmut_kmer(r, kmer_size, kmer_count);
}
/* fracminhash_model.cpp:9-21 */
double expected_containment_index(double r, size_t kmer_size) { return pow((1.0 - r), kmer_size); }
double variance_containment_index(double r, size_t kmer_size, size_t kmer_count, double scaling_factor)
{
    double term3 = variance_nmut_kmer(r, kmer_size, kmer_count) / pow(kmer_count, 2);
    double term2 = kmer_count * expected_nmut_kmer(r, kmer_size, kmer_count) - expected_nmut_kmer_squared(r, kmer_size, kmer_count);
    double denominator = scaling_factor * pow(kmer_count, 3) * pow(1.0 - pow(1.0 - scaling_factor, kmer_count), 2);
    double term1 = (1.0 - scaling_factor) / denominator;
    return term1 * term2 + term3;
}
} // namespace

/* gaussian_inverse.cpp:28-52 (the throw for p outside (0,1) becomes NaN here; never reached: p = 0.975) */
extern "C" double orc_normal_cdf_inverse(double p)
{
    if (p <= 0.0 || p >= 1.0)
        return NAN;
    if (p < 0.5)
        return -rational_approximation(sqrt(-2.0 * log(p)));
    return rational_approximation(sqrt(-2.0 * log(1.0 - p)));
}

/* kmer_model.cpp:10-26 */
extern "C" void orc_kmer_ci(double r, uint64_t kmer_size, uint64_t kmer_count, double confidence,
                            uint64_t *low, uint64_t *high)
{
    double q = 1.0 - pow(1.0 - r, kmer_size);
    double varN = variance_nmut_kmer(r, kmer_size, kmer_count);
    double alpha = 1 - confidence;
    double z = orc_normal_cdf_inverse(1.0 - alpha / 2.0);
    *low = static_cast<size_t>(floor(kmer_count * q - z * sqrt(varN)));
    *high = static_cast<size_t>(ceil(kmer_count * q + z * sqrt(varN)));
}

/* syncmer_model.hpp:38-50 (asserts are compiled out in the reference's Release build) */
extern "C" double orc_syncmer_match_ratio(uint64_t kmer_size, double error_rate)
{
    size_t row_index = ceil((1.0 - error_rate) * 100.0 - 80.0);
    size_t col_index = kmer_size - 10 - ((kmer_size - 10) / 2) - 1;
    if (row_index > 20 || col_index > 9) /* out of the table: undefined in the reference */
        return NAN;
    return matching_ratios[row_index][col_index];
}

/* threshold.hpp:22-49 (the std::cout model banner is the CLI's job) */
extern "C" void orc_threshold_init(orc_thresholder *t, uint32_t window_size, uint8_t kmer_size, double percentage,
                                   double error_rate, int use_syncmer, int fracminhash)
{
    t->kmer_size = kmer_size;
    t->error_rate = error_rate;
    t->percentage = 0.0;
    size_t kmers_per_window = (size_t)window_size - kmer_size + 1;
    if (percentage > 0.0 && percentage <= 1.0)
    {
        t->kind = ORC_THR_PERCENTAGE;
        t->percentage = percentage;
    }
    else if (use_syncmer)
        t->kind = ORC_THR_SYNCMER;
    else if (kmers_per_window == 1 && !fracminhash)
        t->kind = ORC_THR_KMER;
    else
        t->kind = ORC_THR_FRACMINHASH;
}

/* threshold.hpp:51-81 */
extern "C" uint64_t orc_threshold_get(const orc_thresholder *t, uint64_t minimiser_count, double scaling_factor)
{
    size_t fp_correction = minimiser_count * 0.0039;
    switch (t->kind)
    {
    case ORC_THR_SYNCMER:
    {
        double ratio = orc_syncmer_match_ratio(t->kmer_size, t->error_rate);
        return static_cast<size_t>(minimiser_count * ratio);
    }
    case ORC_THR_KMER:
    {
        uint64_t lo, hi;
        orc_kmer_ci(t->error_rate, (size_t)t->kmer_size, minimiser_count, 0.95, &lo, &hi);
        return minimiser_count - hi - fp_correction;
    }
    case ORC_THR_FRACMINHASH:
    {
        double z_alpha = orc_normal_cdf_inverse(1.0 - (1.0 - 0.95) / 2.0);
        double clow = expected_containment_index(t->error_rate, t->kmer_size)
                    - z_alpha * sqrt(variance_containment_index(t->error_rate, t->kmer_size, minimiser_count, scaling_factor));
        return static_cast<size_t>(clow * minimiser_count) - fp_correction;
    }
    default:
        return static_cast<size_t>(minimiser_count * t->percentage);
    }
}

/* ------------------------------------------------------------------------------------------------
 * IXF / HIXF
 * ---------------------------------------------------------------------------------------------- */

static ixfref_scheme g_scheme = {0, 0, 0, 21, 42};
extern "C" void orc_set_ixf_scheme(const uint32_t *s5)
{
    if (!s5)
        g_scheme = ixfref_scheme{0, 0, 0, 21, 42};
    else
        g_scheme = ixfref_scheme{s5[0], s5[1], s5[2], (s5[3] | s5[4]) ? s5[3] : 21u, (s5[3] | s5[4]) ? s5[4] : 42u};
}

extern "C" void orc_ixf_slots(uint64_t key, uint64_t seed, uint64_t seg_len,
                              uint8_t *f, uint64_t *p0, uint64_t *p1, uint64_t *p2)
{
    const uint64_t h = ixfref_mix(key, seed);
    *f = ixfref_fingerprint(h);
    *p0 = ixfref_slot(h, 0, seg_len);
    *p1 = ixfref_slot(h, 1, seg_len);
    *p2 = ixfref_slot(h, 2, seg_len);
}

/* ---- "tuned" CPU baseline (bench.py cpu_baseline.port_tuned; not used by the parity tests) ----
 * The same bulk_count, written the way a CPU implementation that cares would write it (the fork's agent is SIMD as well,
 * SURVEY 2.1): the three row addresses of value i+D are computed and prefetched while value i is reduced, rows are compared
 * 64 bins at a time (AVX-512BW when the build has it: xor, compare-equal against the broadcast fingerprint, byte counters
 * bumped through the mask), byte counters spill into the u32 counts every 255 values.  Same results, fewer stalls. */
static int g_tuned = 0;
extern "C" void orc_set_tuned(int on) { g_tuned = on; }
extern "C" int orc_build_flags(void)
{
    int f = 0;
#if defined(__AVX512BW__)
    f |= 1;
#endif
#if defined(__AVX2__)
    f |= 2;
#endif
    return f;
}

static void bulk_count_tuned(const orc_ixf *x, const uint64_t *values, uint64_t n, uint32_t *counts)
{
    constexpr int D = 12; /* prefetch distance in values: ~36 lines in flight per thread */
    const uint64_t rows = x->rows ? x->rows : 3 * x->seg_len;
    const uint64_t tb = x->tbins, chunks = tb / 64; /* tbins is a multiple of 64 (HBM/IXF row padding) */
    struct Slot
    {
        const uint8_t *r0, *r1, *r2;
        uint8_t f;
    } ring[D];
    std::vector<uint8_t> acc(tb, 0);
    std::memset(counts, 0, sizeof(uint32_t) * x->bins);
    auto flush = [&]() {
        for (uint64_t b = 0; b < x->bins; ++b)
            counts[b] += acc[b];
        std::memset(acc.data(), 0, tb);
    };
    uint32_t since_flush = 0;
    for (uint64_t i = 0; i < n + D; ++i)
    {
        if (i >= D)
        {
            const Slot &s = ring[i % D];
#if defined(__AVX512BW__)
            const __m512i fv = _mm512_set1_epi8((char)s.f);
            for (uint64_t c = 0; c < chunks; ++c)
            {
                const __m512i v = _mm512_xor_si512(_mm512_xor_si512(_mm512_loadu_si512(s.r0 + 64 * c), _mm512_loadu_si512(s.r1 + 64 * c)),
                                                   _mm512_loadu_si512(s.r2 + 64 * c));
                const __mmask64 m = _mm512_cmpeq_epi8_mask(v, fv);
                __m512i a = _mm512_loadu_si512(acc.data() + 64 * c);
                a = _mm512_sub_epi8(a, _mm512_movm_epi8(m)); /* mask bytes are 0xFF: subtracting adds one */
                _mm512_storeu_si512(acc.data() + 64 * c, a);
            }
#else
            for (uint64_t b = 0; b < tb; ++b)
                acc[b] += (uint8_t)((uint8_t)(s.r0[b] ^ s.r1[b] ^ s.r2[b]) == s.f);
#endif
            if (++since_flush == 255)
            {
                flush();
                since_flush = 0;
            }
        }
        if (i < n)
        {
            const uint64_t h = ixfref_mix_s(&g_scheme, values[i], x->seed);
            Slot &s = ring[i % D];
            s.f = ixfref_fingerprint_s(&g_scheme, h);
            s.r0 = x->data + ixfref_slot_s(&g_scheme, h, 0, x->seg_len, rows) * tb;
            s.r1 = x->data + ixfref_slot_s(&g_scheme, h, 1, x->seg_len, rows) * tb;
            s.r2 = x->data + ixfref_slot_s(&g_scheme, h, 2, x->seg_len, rows) * tb;
            for (uint64_t c = 0; c < chunks && c < 4; ++c) /* the first lines of a wide row; the hardware streams the rest */
            {
                __builtin_prefetch(s.r0 + 64 * c, 0, 0);
                __builtin_prefetch(s.r1 + 64 * c, 0, 0);
                __builtin_prefetch(s.r2 + 64 * c, 0, 0);
            }
        }
    }
    flush();
}

/* seqan3::interleaved_xor_filter<uint8_t>::counting_agent<uint32_t>().bulk_count (call site hixf.hpp:307-309).
 * PARITY UNPINNED (see ixf_ref.h): per value, fingerprint == xor of the three interleaved rows -> bin hit. */
extern "C" void orc_ixf_bulk_count(const orc_ixf *x, const uint64_t *values, uint64_t n, uint32_t *counts)
{
    if (g_tuned && x->tbins % 64 == 0)
    {
        bulk_count_tuned(x, values, n, counts);
        return;
    }
    std::memset(counts, 0, sizeof(uint32_t) * x->bins);
    for (uint64_t v = 0; v < n; ++v)
    {
        const uint64_t rows = x->rows ? x->rows : 3 * x->seg_len;
        const uint64_t h = ixfref_mix_s(&g_scheme, values[v], x->seed);
        const uint8_t f = ixfref_fingerprint_s(&g_scheme, h);
        const uint64_t p0 = ixfref_slot_s(&g_scheme, h, 0, x->seg_len, rows), p1 = ixfref_slot_s(&g_scheme, h, 1, x->seg_len, rows),
                       p2 = ixfref_slot_s(&g_scheme, h, 2, x->seg_len, rows);
        const uint8_t *r0 = x->data + p0 * x->tbins;
        const uint8_t *r1 = x->data + p1 * x->tbins;
        const uint8_t *r2 = x->data + p2 * x->tbins;
        for (uint64_t b = 0; b < x->bins; ++b)
            counts[b] += (uint8_t)(r0[b] ^ r1[b] ^ r2[b]) == f;
    }
}

namespace
{
struct dfs_out
{
    int64_t *ub;
    uint32_t *cnt;
    int64_t cap;
    int64_t n;
    uint64_t bytes;
};

/* hixf.hpp:303-340 */
void bulk_contains_impl(const orc_hixf *h, const uint64_t *values, uint64_t n, int64_t ixf_idx,
                        uint64_t threshold, dfs_out &o)
{
    const orc_ixf *x = &h->ixf[ixf_idx];
    std::vector<uint32_t> result(x->bins);
    orc_ixf_bulk_count(x, values, n, result.data());                               /* :307-309 */
    o.bytes += n * 3 * x->tbins + 8 * n;                                           /* SURVEY 8(d) algorithmic bytes */
    const int64_t *ub = h->bin_to_ub + h->bin_off[ixf_idx];
    const int64_t *nx = h->next_ixf_id + h->bin_off[ixf_idx];
    uint32_t sum = 0;
    for (size_t bin = 0; bin < result.size(); ++bin)
    {
        sum += result[bin];                                                        /* :315 */
        const int64_t current = ub[bin];                                           /* :317 */
        if (current < 0)                                                           /* merged bin :319-324 */
        {
            if (sum >= threshold)
                bulk_contains_impl(h, values, n, nx[bin], threshold, o);
            sum = 0u;
        }
        else if (bin + 1u == result.size() || current != ub[bin + 1])              /* end of split bin :325-334 */
        {
            if (sum >= threshold)
            {
                if (o.n < o.cap)
                {
                    o.ub[o.n] = current;
                    o.cnt[o.n] = sum;
                }
                ++o.n;
            }
            sum = 0u;
        }
    }
}
} // namespace

/* hixf.hpp:381-406 */
extern "C" int64_t orc_hixf_bulk_contains(const orc_hixf *h, const uint64_t *values, uint64_t n, uint64_t threshold,
                                          int64_t *out_ub, uint32_t *out_cnt, int64_t cap, uint64_t *visited_bytes)
{
    dfs_out o{out_ub, out_cnt, cap, 0, 0};
    bulk_contains_impl(h, values, n, 0, threshold, o);
    if (visited_bytes)
        *visited_bytes += o.bytes;
    return o.n <= cap ? o.n : -o.n;
}

/* ------------------------------------------------------------------------------------------------
 * whole per-read flow, taxor_search.cpp:196-313 (everything except string formatting)
 * ---------------------------------------------------------------------------------------------- */
extern "C" int orc_search_batch(const orc_hixf *h, const orc_search_params *p,
                                const uint8_t *codes, const uint64_t *off, uint64_t n_reads,
                                uint32_t *hash_count, uint64_t *threshold,
                                uint64_t *hit_off, int64_t *out_ub, uint32_t *out_cnt, uint64_t hit_cap,
                                uint64_t *raw_off, int64_t *raw_ub, uint32_t *raw_cnt, uint64_t raw_cap,
                                uint64_t *visited_bytes_total, int n_threads)
{
    orc_thresholder thr;
    orc_threshold_init(&thr, p->window_size, (uint8_t)p->k, p->percentage, p->error_rate, p->use_syncmer, 0);
    const uint64_t kseed = orc_adjust_seed((uint8_t)p->k);

    std::vector<std::vector<std::pair<int64_t, uint32_t>>> raw(n_reads), fin(n_reads);
    uint64_t bytes_total = 0;
#ifdef _OPENMP
    if (n_threads > 0)
        omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bytes_total)
    for (int64_t r = 0; r < (int64_t)n_reads; ++r)
    {
        const uint8_t *seq = codes + off[r];
        const int64_t len = (int64_t)(off[r + 1] - off[r]);
        std::vector<uint64_t> hashes;
        const int64_t windows = len >= p->k ? len - p->k + 1 : 0;
        std::vector<uint64_t> tmp((size_t)windows + 1);
        if (p->use_syncmer)                                                        /* :219-238 */
        {
            int64_t n = orc_syncmer_hashes(seq, len, p->k, p->s, p->t, tmp.data(), windows + 1);
            for (int64_t i = 0; i < n; ++i)
                if (p->scaling <= 1 || orc_scaling_keep(tmp[i], p->scaling))
                    hashes.push_back(tmp[i]);
        }
        else                                                                       /* :240-256 */
        {
            int64_t n = (int64_t)p->window_size > p->k
                            ? orc_minimiser_hashes(seq, len, p->k, (int)p->window_size, kseed, tmp.data(), windows + 1)
                            : orc_kmer_hashes(seq, len, p->k, kseed, tmp.data(), windows + 1);
            for (int64_t i = 0; i < n; ++i)
                if (p->scaling <= 1 || orc_scaling_keep(tmp[i], p->scaling))
                    hashes.push_back(tmp[i]);
        }
        const size_t hc = hashes.size();                                           /* :261 */
        const uint64_t t = orc_threshold_get(&thr, hc, (double)hc / ((double)len - (double)p->k + 1.0)); /* :263 */
        hash_count[r] = (uint32_t)hc;
        threshold[r] = t;

        int64_t cap = 64;
        std::vector<int64_t> ub((size_t)cap);
        std::vector<uint32_t> cnt((size_t)cap);
        uint64_t bytes = 0;
        int64_t n = orc_hixf_bulk_contains(h, hashes.data(), hc, t, ub.data(), cnt.data(), cap, &bytes); /* :265 */
        if (n < 0)
        {
            cap = -n;
            ub.resize((size_t)cap);
            cnt.resize((size_t)cap);
            n = orc_hixf_bulk_contains(h, hashes.data(), hc, t, ub.data(), cnt.data(), cap, nullptr);
        }
        bytes_total += bytes;
        uint64_t max_count = 0;                                                    /* :275-280 */
        for (int64_t i = 0; i < n; ++i)
        {
            raw[r].emplace_back(ub[i], cnt[i]);
            if (cnt[i] > max_count)
                max_count = cnt[i];
        }
        for (int64_t i = 0; i < n; ++i)                                            /* :282-286 */
        {
            if (static_cast<double>(cnt[i]) < static_cast<double>(max_count) * 0.8)
                continue;
            fin[r].emplace_back(ub[i], cnt[i]);
        }
    }
    if (visited_bytes_total)
        *visited_bytes_total = bytes_total;

    uint64_t nf = 0, nr = 0;
    int rc = 0;
    for (uint64_t r = 0; r < n_reads; ++r)
    {
        hit_off[r] = nf;
        for (auto &pr : fin[r])
        {
            if (nf < hit_cap)
            {
                out_ub[nf] = pr.first;
                out_cnt[nf] = pr.second;
            }
            ++nf;
        }
        if (raw_off)
        {
            raw_off[r] = nr;
            for (auto &pr : raw[r])
            {
                if (nr < raw_cap)
                {
                    raw_ub[nr] = pr.first;
                    raw_cnt[nr] = pr.second;
                }
                ++nr;
            }
        }
    }
    hit_off[n_reads] = nf;
    if (raw_off)
        raw_off[n_reads] = nr;
    if (nf > hit_cap)
        rc = -1;
    if (raw_off && nr > raw_cap)
        rc = -2;
    return rc;
}
