/*
 * taxor_oracle.h -- CPU ORACLE for the `taxor search` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C-ABI restatement of the reference algorithm (JensUweUlrich/Taxor) used as the
 * checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load it.  The product (taxor_b200/) never links or calls it.
 *
 * PARITY STATUS
 *   pinned   : syncmer scan + tie rules, HIXF DFS (split/merged logic), threshold models -- checked
 *              against the reference's OWN sources compiled in place (oracle/_ref, see Makefile) and
 *              against the known-answer vectors in tests/golden/.
 *   UNPINNED : (1) seqan3::interleaved_xor_filter<uint8_t> probe arithmetic + cereal field order
 *              (lives in the un-vendored fork JensUweUlrich/seqan3 @ master, absent from
 *              /root/reference) -- restated from the in-tree prototype src/main/xorfilter.hpp:22-45,
 *              60-62,336-350 + src/main/hashutil.hpp:50-61 and isolated in ixf_ref.h;
 *              (2) ankerl::unordered_dense v3.0.1 detail::wyhash::hash(uint64_t) -- restated from the
 *              published algorithm (mix(x, 0x9E3779B97F4A7C15), 128-bit product, lo ^ hi);
 *              (3) seqan3::views::minimiser_hash canonical k-mer semantics (SeqAn3 3.3.0 upstream).
 *              "parity unpinned" for those three third-party pieces; see DESIGN.md.
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#ifndef TAXOR_ORACLE_H
#define TAXOR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- scalar pieces ---------- */
uint64_t orc_wyhash_u64(uint64_t x);                    /* src/hashing/syncmer.cpp:73-77 (ankerl v3.0.1) */
uint64_t orc_adjust_seed(uint8_t kmer_size);            /* src/hixf/build/adjust_seed.hpp:40-44 */
int      orc_dna4_rank(unsigned char c);                /* src/hixf/build/dna4_traits.hpp:15-18 (seqan3::dna4 char_to_rank); -1 = illegal */
int      orc_t_syncmer(int k, int s);                   /* src/main/taxor_build.cpp:510 */
int      orc_scaling_keep(uint64_t h, uint16_t scaling);/* src/main/taxor_search.cpp:223-233 */

/* ---------- hashing ---------- */
/* codes[i] in {0,1,2,3} (A,C,G,T) or 4 ("N" restart branch, dead in search -- SURVEY 3.5).
 * Distinct hashes in first-insertion order (ankerl set iterates in insertion order).
 * Returns the number of distinct hashes, or -(needed) if cap is too small.            */
int64_t orc_syncmer_hashes(const uint8_t *codes, int64_t len, int k, int s, int t,
                           uint64_t *out, int64_t cap);  /* src/hashing/syncmer.cpp:80-165 */
/* Same scan but every emission is kept (no set): used to debug the GPU raw stage. */
int64_t orc_syncmer_hashes_raw(const uint8_t *codes, int64_t len, int k, int s, int t,
                               uint64_t *out, int64_t cap);
/* canonical k-mer mode, window == k: min(fwd ^ seed, rc ^ seed) per position, duplicates kept. */
int64_t orc_kmer_hashes(const uint8_t *codes, int64_t len, int k, uint64_t seed,
                        uint64_t *out, int64_t cap);     /* src/main/taxor_search.cpp:210-212,240-256 */

/* minimiser mode, window w > k: seqan3::views::minimiser over the values above, W = w-k+1 values per window,
 * rightmost minimum, re-reported whenever the tracked minimiser leaves the window (upstream SeqAn3 3.3.0). */
int64_t orc_minimiser_hashes(const uint8_t *codes, int64_t len, int k, int w, uint64_t seed,
                             uint64_t *out, int64_t cap);

/* ---------- thresholds ---------- */
enum { ORC_THR_FRACMINHASH = 0, ORC_THR_PERCENTAGE = 1, ORC_THR_KMER = 2, ORC_THR_SYNCMER = 3 };
typedef struct {
    int     kind;
    uint8_t kmer_size;
    double  percentage;
    double  error_rate;
} orc_thresholder;
void     orc_threshold_init(orc_thresholder *t, uint32_t window_size, uint8_t kmer_size, double percentage,
                            double error_rate, int use_syncmer, int fracminhash); /* threshold.hpp:22-49 */
uint64_t orc_threshold_get(const orc_thresholder *t, uint64_t count, double scaling_factor); /* threshold.hpp:51-81 */
double   orc_syncmer_match_ratio(uint64_t kmer_size, double error_rate);           /* syncmer_model.hpp:38-50 */
double   orc_normal_cdf_inverse(double p);                                          /* gaussian_inverse.cpp:28-52 */
void     orc_kmer_ci(double r, uint64_t kmer_size, uint64_t kmer_count, double confidence,
                     uint64_t *low, uint64_t *high);                                /* kmer_model.cpp:10-26 */

/* ---------- IXF / HIXF ---------- */
typedef struct {
    uint64_t seed;
    uint64_t bins;        /* user-visible technical bins == counting vector size            */
    uint64_t tbins;       /* stored row width in fingerprints (bins padded to 64)           */
    uint64_t seg_len;     /* slots per segment; 3*seg_len rows                              */
    const uint8_t *data;  /* data[slot * tbins + bin]                                       */
    uint64_t rows;        /* slots per bin, 0 = 3*seg_len (binary fuse: (segments+2)*seg_len) */
} orc_ixf;

/* The probe arithmetic is a hypothesis (see ixf_ref.h), so the oracle can be switched between candidate schemes:
 * scheme5 = {slots, mix, fingerprint, rot1, rot2}; NULL restores the prototype's.  Process-global (test infrastructure). */
void orc_set_ixf_scheme(const uint32_t *scheme5);
/* bulk_count with software prefetch + 64-bin SIMD compares (same results; the "port_tuned" CPU baseline of bench.py) */
void orc_set_tuned(int on);
int  orc_build_flags(void);   /* bit 0: compiled with AVX-512BW, bit 1: AVX2 */

typedef struct {
    uint64_t        n_ixf;
    const orc_ixf  *ixf;
    const uint64_t *bin_off;      /* n_ixf+1 offsets into the two per-bin arrays            */
    const int64_t  *next_ixf_id;  /* hixf.hpp:122                                           */
    const int64_t  *bin_to_ub;    /* user_bins.ixf_bin_to_filename_position, -1 = merged    */
} orc_hixf;

void orc_ixf_slots(uint64_t key, uint64_t seed, uint64_t seg_len,
                   uint8_t *f, uint64_t *p0, uint64_t *p1, uint64_t *p2);           /* ixf_ref.h */
void orc_ixf_bulk_count(const orc_ixf *x, const uint64_t *values, uint64_t n, uint32_t *counts);
/* DFS of hixf.hpp:303-340 + 381-406.  Writes (user_bin,count) pairs in DFS pre-order.
 * Returns number of pairs, or -(needed) when cap is too small. visited_bytes (may be NULL)
 * accumulates the algorithmic bytes  sum_x (n*3*tbins_x + 8*n)  of SURVEY 8(d).           */
int64_t orc_hixf_bulk_contains(const orc_hixf *h, const uint64_t *values, uint64_t n, uint64_t threshold,
                               int64_t *out_ub, uint32_t *out_cnt, int64_t cap, uint64_t *visited_bytes);

/* ---------- whole per-read flow (taxor_search.cpp:196-313) for a batch, OpenMP over reads ---------- */
typedef struct {
    int      k, s, t;
    int      use_syncmer;
    uint32_t window_size;
    uint16_t scaling;
    double   percentage;      /* <=0 -> model */
    double   error_rate;
} orc_search_params;

/* codes: concatenated base codes; off[n_reads+1].  Per read r writes hash_count[r], threshold[r],
 * and appends hits AFTER the 0.8*max filter (taxor_search.cpp:275-286) to out_* with hit_off[n_reads+1].
 * raw_* (optional, may be NULL) receive the unfiltered DFS result.  Returns 0 or <0 on overflow. */
int orc_search_batch(const orc_hixf *h, const orc_search_params *p,
                     const uint8_t *codes, const uint64_t *off, uint64_t n_reads,
                     uint32_t *hash_count, uint64_t *threshold,
                     uint64_t *hit_off, int64_t *out_ub, uint32_t *out_cnt, uint64_t hit_cap,
                     uint64_t *raw_off, int64_t *raw_ub, uint32_t *raw_cnt, uint64_t raw_cap,
                     uint64_t *visited_bytes_total, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
