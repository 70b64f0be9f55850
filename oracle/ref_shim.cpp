/*
 * ref_shim.cpp -- C-ABI shim around the REFERENCE'S OWN SOURCES, compiled where they lie under
 * /root/reference (never copied).  TEST INFRASTRUCTURE: outputs go to oracle/_ref/ only.
 *
 * Compiled in place (see Makefile):
 *     src/hashing/syncmer.cpp                                   (syncmer scan, tie rules)
 *     src/hixf/build/hierarchical_interleaved_xor_filter.hpp    (membership_agent DFS)
 *     src/hixf/search/threshold.hpp + syncmer_model.hpp         (threshold dispatch + table)
 *     src/hixf/search/{kmer_model,fracminhash_model,gaussian_inverse}.cpp
 * The third-party headers those files include (SeqAn3 fork, ankerl::unordered_dense) are NOT in
 * /root/reference; oracle/stubs/ provides labelled stand-ins, so wyhash and the IXF probe stay
 * "parity unpinned" while the Taxor-owned logic above is the real thing.
 */
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include <syncmer.hpp>                                      // /root/reference/src/hashing
#include <build/hierarchical_interleaved_xor_filter.hpp>    // /root/reference/src/hixf
#include <search/threshold.hpp>                             // /root/reference/src/hixf
#include <search/gaussian_inverse.hpp>
#include <search/kmer_model.hpp>

extern "C" {

/* hashing::seq_to_syncmers (src/hashing/syncmer.cpp:157-165); codes 0..3 = ACGT, 4 = N */
int64_t ref_seq_to_syncmers(const uint8_t *codes, int64_t len, int k, int s, int t, uint64_t *out, int64_t cap)
{
    static const char alphabet[] = "ACGTN";
    seqan3::dna5_vector seq(static_cast<size_t>(len));
    for (int64_t i = 0; i < len; ++i)
        seq[static_cast<size_t>(i)].c = alphabet[codes[i] > 4 ? 4 : codes[i]];
    auto set = hashing::seq_to_syncmers(k, seq, s, t);
    int64_t n = 0;
    for (auto h : set)
    {
        if (n < cap)
            out[n] = h;
        ++n;
    }
    return n <= cap ? n : -n;
}

uint64_t ref_wyhash_stub(uint64_t x) { return ankerl::unordered_dense::detail::wyhash::hash(x); }

/* hixf::hierarchical_interleaved_xor_filter<uint8_t>::membership_agent::bulk_contains (hixf.hpp:381-406) */
struct ref_hixf
{
    hixf::hierarchical_interleaved_xor_filter<uint8_t> h;
};

void *ref_hixf_new(uint64_t n_ixf, const uint64_t *seed, const uint64_t *bins, const uint64_t *tbins,
                   const uint64_t *seg_len, const uint8_t *const *data, const uint64_t *bin_off,
                   const int64_t *next_ixf_id, const int64_t *bin_to_ub)
{
    auto *r = new ref_hixf{};
    r->h.user_bins.set_ixf_count(n_ixf);
    r->h.next_ixf_id.resize(n_ixf);
    for (uint64_t i = 0; i < n_ixf; ++i)
    {
        r->h.ixf_vector.emplace_back(bins[i], tbins[i], seg_len[i], seed[i], data[i]);
        r->h.next_ixf_id[i].assign(next_ixf_id + bin_off[i], next_ixf_id + bin_off[i + 1]);
        r->h.user_bins.bin_indices_of_ixf(i).assign(bin_to_ub + bin_off[i], bin_to_ub + bin_off[i + 1]);
    }
    return r;
}
void ref_hixf_free(void *p) { delete static_cast<ref_hixf *>(p); }

int64_t ref_bulk_contains(void *p, const uint64_t *values, uint64_t n, uint64_t threshold,
                          int64_t *out_ub, uint32_t *out_cnt, int64_t cap)
{
    auto *r = static_cast<ref_hixf *>(p);
    auto agent = r->h.membership_agent();
    std::vector<uint64_t> v(values, values + n);
    auto & result = agent.bulk_contains(v, static_cast<size_t>(threshold));
    int64_t k = 0;
    for (auto && pr : result)
    {
        if (k < cap)
        {
            out_ub[k] = pr.first;
            out_cnt[k] = pr.second;
        }
        ++k;
    }
    return k <= cap ? k : -k;
}

/* hixf::threshold::threshold (threshold.hpp:22-81) */
void *ref_threshold_new(uint32_t window_size, uint8_t kmer_size, double percentage, double error_rate,
                        int use_syncmer, int fracminhash)
{
    hixf::threshold_parameters par{};
    par.window_size = window_size;
    par.kmer_size = kmer_size;
    par.percentage = percentage;
    par.seq_error_rate = error_rate;
    par.use_syncmer = use_syncmer != 0;
    par.fracminhash = fracminhash != 0;
    std::ostringstream sink;                     // the ctor prints the model banner to std::cout
    auto *old = std::cout.rdbuf(sink.rdbuf());
    auto *t = new hixf::threshold::threshold{par};
    std::cout.rdbuf(old);
    return t;
}
void ref_threshold_free(void *p) { delete static_cast<hixf::threshold::threshold *>(p); }
uint64_t ref_threshold_get(void *p, uint64_t count, double scaling_factor)
{
    return static_cast<hixf::threshold::threshold *>(p)->get(static_cast<size_t>(count), scaling_factor);
}
double ref_syncmer_match_ratio(uint64_t k, double e) { return hixf::threshold::get_min_syncmer_match_ratio(k, e); }
double ref_normal_cdf_inverse(double p) { return hixf::threshold::NormalCDFInverse(p); }
void ref_kmer_ci(double r, uint64_t k, uint64_t n, double conf, uint64_t *lo, uint64_t *hi)
{
    auto ci = hixf::threshold::calculate_nmut_kmer_CI(r, k, n, conf);
    *lo = ci.first;
    *hi = ci.second;
}
uint64_t ref_adjust_seed(uint8_t k);
}

#include <build/adjust_seed.hpp>                            // /root/reference/src/hixf
extern "C" uint64_t ref_adjust_seed(uint8_t k) { return hixf::adjust_seed(k); }
