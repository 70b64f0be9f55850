/*
 * taxor_b200.h -- C ABI of the B200-native `taxor search` hot path (libtaxor_b200.so).
 *
 * The reference (JensUweUlrich/Taxor) has no FFI layer; its natural seam is the body of the `worker`
 * lambda in src/main/taxor_search.cpp:196-313, i.e. exactly two calls per read
 *     hashing::seq_to_syncmers(k, seq, s, t)                       src/main/taxor_search.cpp:222
 *     membership_agent::bulk_contains(hashes, threshold)           src/main/taxor_search.cpp:265
 * plus thresholder.get() (:263) and the 0.8*max filter (:275-286).  A GPU needs batches, so every entry
 * point below is the batch-granular replacement of one of those calls; INTEGRATION.md shows the binding a
 * maintainer would add to search_single().  All `file:line` citations are relative to /root/reference/.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (TXR_OK) or a negative error code
 * and never throws; handles are opaque; the caller owns every host buffer it passes in; result views stay
 * valid until the next search call on the same context; one context per GPU; calls on one
 * context must be serialised by the caller; distinct contexts may be used from distinct host threads.
 * There is NO CPU fallback: without a CUDA device txr_ctx_create fails with TXR_ERR_CUDA.
 */
#ifndef TAXOR_B200_H
#define TAXOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TXR_OK 0
#define TXR_ERR_CUDA (-1)        /* CUDA runtime error; see txr_last_error() */
#define TXR_ERR_ARG (-2)         /* invalid argument */
#define TXR_ERR_STATE (-3)       /* call order (no index / params yet) */
#define TXR_ERR_UNSUPPORTED (-4) /* parameter combination outside the GPU path (e.g. window_size > k) */
#define TXR_ERR_FORMAT (-5)      /* malformed input (illegal base, bad .hixf) */
#define TXR_ERR_OVERFLOW (-6)    /* internal capacity exceeded after retries */
#define TXR_ERR_IO (-7)

typedef struct txr_ctx txr_ctx;
typedef struct txr_reads txr_reads;   /* a read set resident in HBM */

const char *txr_last_error(void);      /* thread-local, human-readable */
const char *txr_version(void);

/* ------------------------------------------------------------------------------------------------------
 * Index: replaces the in-memory hixf::hierarchical_interleaved_xor_filter<uint8_t>
 * (src/hixf/build/hierarchical_interleaved_xor_filter.hpp:78-160: ixf_vector, next_ixf_id, user_bins).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct
{
    uint64_t seed;       /* per-IXF hash seed (re-seeded on build failure, construct_ixf.cpp:101-108)      */
    uint64_t bins;       /* counting-vector size == ixf_bin_to_filename_position[i].size() (hixf.hpp:313)  */
    uint64_t tbins;      /* stored fingerprints per slot row (bins padded to a multiple of 64)             */
    uint64_t seg_len;    /* slots per segment (XOR3: the filter has 3*seg_len rows; FUSE3: a power of two) */
    const uint8_t *fp;   /* host pointer, fp[slot * tbins + bin] (or bin-major, see txr_ixf_scheme.layout) */
    uint64_t rows;       /* slots per bin; 0 = 3*seg_len.  FUSE3: (segment_count + 2) * seg_len            */
} txr_ixf_view;

/* The probe arithmetic of seqan3::interleaved_xor_filter<uint8_t> (call site hixf.hpp:307-309) lives in the SeqAn3 fork
 * JensUweUlrich/seqan3@master, which is in neither the reference tree nor this image: PARITY UNPINNED.  It is therefore a
 * descriptor handed over with the index, not a compile-time fact.  All-zero (or a NULL pointer) selects the arithmetic of
 * the same author's in-tree prototype (src/main/xorfilter.hpp:22-45,60-62,336-350 + src/main/hashutil.hpp:50-61). */
#define TXR_IXF_SLOTS_XOR3 0u   /* three equal segments: reduce((u32)rotl64(h, rot_i), seg_len) + i*seg_len             */
#define TXR_IXF_SLOTS_FUSE3 1u  /* 3-wise binary fuse (Graf & Lemire 2022): h0 = mulhi64(h, count*L), h1/h2 in the next two
                                   segments, low bits xor-ed with (h >> 18) / h  (the fork also ships an
                                   interleaved_binary_fuse_filter, main.cpp:22)                                          */
#define TXR_IXF_MIX_ADD_SEED 0u /* murmur fmix64(key + seed)   (hashutil.hpp:50-61)                                      */
#define TXR_IXF_MIX_XOR_SEED 1u /* murmur fmix64(key ^ seed)                                                            */
#define TXR_IXF_FP_FOLD32 0u    /* (u8)(h ^ (h >> 32))         (xorfilter.hpp:60-62)                                     */
#define TXR_IXF_FP_LOW8 1u      /* (u8)h                                                                                */
#define TXR_IXF_FP_HIGH8 2u     /* (u8)(h >> 56)                                                                        */
#define TXR_IXF_LAYOUT_SLOT_MAJOR 0u /* fp[slot * tbins + bin]: the bins of one slot are contiguous (interleaved)        */
#define TXR_IXF_LAYOUT_BIN_MAJOR 1u  /* fp[bin * rows + slot]: one plain filter after the other; re-laid-out on upload   */
typedef struct
{
    uint32_t slots, mix, fingerprint;
    uint32_t rot1, rot2;             /* XOR3 rotations for segments 1 and 2; 0,0 = the prototype's 21,42                 */
    uint32_t layout;                 /* how the HOST arrays are laid out; HBM is always slot-major                        */
} txr_ixf_scheme;

typedef struct
{
    uint64_t n_ixf;
    const txr_ixf_view *ixf;
    const uint64_t *bin_off;           /* n_ixf+1 offsets into the two per-bin arrays                      */
    const int64_t *next_ixf_id;        /* hixf.hpp:122                                                     */
    const int64_t *bin_to_user_bin;    /* user_bins.ixf_bin_to_filename_position (hixf.hpp:178), -1=merged */
    uint64_t n_user_bins;
    const txr_ixf_scheme *scheme;      /* NULL = the prototype's arithmetic, slot-major arrays             */
} txr_hixf_view;

/* Parameters that the reference takes from the .hixf (taxor_search.cpp:165-169) and from the CLI. */
typedef struct
{
    uint8_t kmer_size, syncmer_size, t_syncmer;
    uint8_t use_syncmer;               /* index.use_syncmer()                                              */
    uint32_t window_size;              /* index.window_size(); k-mer mode requires window_size == kmer_size */
    uint16_t scaling;                  /* FracMin scaling (taxor_search.cpp:223-233); 1 = off              */
    double percentage;                 /* --percentage; <= 0 selects the model (threshold.hpp:27-48)       */
    double error_rate;                 /* --error-rate (default 0.04, taxor_search_configuration.hpp:16)   */
} txr_params;

/* bulk_contains result for a batch (hixf.hpp:371-406), one segment per read, each segment in the
 * reference's DFS pre-order; `keep` marks the pairs that survive the 0.8*max filter (taxor_search.cpp:275-286). */
typedef struct
{
    uint64_t n_reads;
    const uint32_t *hash_count;        /* [n_reads]   hashes.size()  (taxor_search.cpp:261)                */
    const uint64_t *threshold;         /* [n_reads]   thresholder.get() (taxor_search.cpp:263)             */
    const uint64_t *hit_begin;         /* [n_reads+1]                                                      */
    const int64_t *user_bin;           /* [n_hits]    result[i].first                                      */
    const uint32_t *count;             /* [n_hits]    result[i].second                                     */
    const uint8_t *keep;               /* [n_hits]                                                         */
} txr_result;

/* Per-stage device time of the most recent search call, measured with CUDA events on the launch stream. */
typedef struct
{
    float h2d_ms, hash_ms, dedup_ms, query_ms, d2h_ms, total_ms;
    uint64_t query_launches, hash_launches, dedup_launches;
    uint64_t query_items;              /* (read, IXF) work items processed                                  */
    uint64_t query_bytes;              /* sum over items of Hp*3*tbins + 8*Hp, Hp = hashes actually probed (= H, SURVEY 8(d), without early exits) */
    uint64_t hash_bytes;               /* sum over reads of ceil(L/4) + 8*H_raw                              */
    uint64_t n_hashes;                 /* sum of hash_count                                                  */
    uint64_t skipped_hashes;           /* probes the early exit of kernel #2 saved (an item stops once no user bin can
                                          reach the read's threshold any more); their bytes are NOT in query_bytes   */
    uint64_t probe_launches;           /* launches of the probe kernels alone (query_launches also counts the
                                          small queue-grouping kernels between levels)                        */
} txr_timing;

/* ---- context ---- */
int txr_ctx_create(int device, txr_ctx **out);
void txr_ctx_destroy(txr_ctx *ctx);
/* The caller's CUDA stream (a cudaStream_t; NULL = legacy default stream).  Every search call forks its internal
 * streams from it and joins them back, so CUDA events recorded on it bracket the whole call. */
int txr_ctx_set_stream(txr_ctx *ctx, void *stream);
/* max reads / bases per internal batch and number of pipeline slots (defaults 262144 / 3e9 / 4). */
int txr_ctx_configure(txr_ctx *ctx, uint64_t max_batch_reads, uint64_t max_batch_bases, int n_slots);

/* optional: allocate now what the first search call of up to n_reads reads / n_bases bases would allocate (needs
 * txr_params_set; a multi-GPU driver calls it from each GPU's thread while the reads are still being parsed) */
int txr_ctx_reserve(txr_ctx *ctx, uint64_t n_reads, uint64_t n_bases);

/* one-time re-layout of the index into HBM; replaces load_index() + index.ixf() (load_index.hpp:27-38) */
int txr_index_upload(txr_ctx *ctx, const txr_hixf_view *index);
/* replicate the index already resident in `src` (same or another GPU) into `dst`, device to device (NVLink between
 * peers): the multi-GPU form of the reference's "load once, share between threads" (taxor_search.cpp:162-180) */
int txr_index_clone(txr_ctx *dst, txr_ctx *src);
int txr_params_set(txr_ctx *ctx, const txr_params *params);
/* hixf::threshold::threshold::get (src/hixf/search/threshold.hpp:51-81) with the context's parameters */
int txr_threshold_get(txr_ctx *ctx, uint64_t hash_count, double scaling_factor, uint64_t *out);
/* the same without a context (pure host code; usable on a machine without a GPU) */
int txr_threshold_eval(const txr_params *params, uint64_t hash_count, double scaling_factor, uint64_t *out);

/* ---- reads: 2-bit packing (A0 C1 G2 T3, seqan3::dna4 collapse of IUPAC, src/hixf/build/dna4_traits.hpp:15-18) ----
 * Layout: read r occupies 64-bit words [word_off[r], word_off[r] + ceil(len/32)] (one zero pad word),
 * base i of a read sits in word i/32 at bits [62-2*(i%32), 63-2*(i%32)] (first base most significant). */
uint64_t txr_packed_words(uint64_t n_bases);                                 /* ceil(n/32) + 1 */
int txr_pack_2bit(const char *ascii, uint64_t len, uint64_t *dst_words);     /* TXR_ERR_FORMAT on a non-IUPAC char */
int txr_pack_codes(const uint8_t *codes, uint64_t len, uint64_t *dst_words); /* codes 0..3 */
int txr_unpack_codes(const uint64_t *words, uint64_t len, uint8_t *codes);
void *txr_host_alloc(size_t bytes);                                          /* pinned host memory */
void *txr_ctx_host_alloc(txr_ctx *ctx, size_t bytes);   /* the same from any thread: makes ctx's device current first */
void txr_host_free(void *p);

/* ---- search: replaces the per-read body of `worker` (taxor_search.cpp:196-313) for n_reads reads ---- */
/* end to end from HOST buffers (pinned or pageable): H2D, hash, dedup, query levels, D2H, host ordering. */
int txr_search(txr_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len,
               uint64_t n_reads, txr_result *out);
/* reads resident in HBM (measurement of the device path without host<->device copies) */
int txr_reads_upload(txr_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len,
                     uint64_t n_reads, txr_reads **out);
void txr_reads_free(txr_ctx *ctx, txr_reads *reads);
/* fetch != 0: also copy the hits back and fill *out (out may be NULL when fetch == 0) */
int txr_search_resident(txr_ctx *ctx, txr_reads *reads, int fetch, txr_result *out);
int txr_get_timing(txr_ctx *ctx, txr_timing *out);

/* ---- build side: hash generation of `taxor build` on the GPU ----
 * Replaces the per-user-bin loop of compute_hashes (src/hixf/build/compute_hashes.cpp:76-142): every sequence of a user
 * bin is hashed with the context's parameters (syncmers / k-mers / minimisers, FracMin scaling filter) and the DISTINCT
 * hashes of the bin are returned (order unspecified: the consumer builds a set / an XOR filter).  Long sequences are cut
 * into segments at windows whose state does not depend on what came before (unique window minimum), so the result is
 * bit-identical to the sequential scan.  seq_bin[] must be non-decreasing; sequences use the packed layout above.
 * The XOR-filter construction itself stays on the CPU (as in the reference).  Buffers are owned by the context. */
typedef struct
{
    uint64_t n_bins;
    const uint64_t *bin_off;           /* [n_bins+1] offsets into hashes                                      */
    const uint64_t *hashes;
    uint64_t n_segments;               /* how many independent pieces the sequences were hashed in            */
} txr_bin_hashes;
/* The cutting alone (pure host code, no GPU): k-mer window indices at which a sequence of `len` bases may be cut into
 * independently hashed pieces of about target_windows windows (>= 64).  cuts[0] is 0; piece i covers the windows
 * [cuts[i], cuts[i+1]), i.e. the bases [cuts[i], cuts[i+1] + span - 1) with span = k (syncmers, k-mers) or window_size. */
int txr_plan_segments(const txr_params *params, const uint64_t *words, uint64_t len, uint64_t target_windows,
                      uint64_t *cuts, uint64_t cap, uint64_t *n_cuts);
int txr_hash_user_bins(txr_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len,
                       uint64_t n_seqs, const uint32_t *seq_bin, uint64_t n_bins, txr_bin_hashes *out);

/* ---- `taxor profile` head fed from in-memory results (host code, no GPU) ----
 * Replaces the TSV round trip between `taxor search` and `taxor profile`: parse_search_results
 * (src/main/taxor_profile.cpp:93-163) becomes txr_profile_add_batch on txr_result batches (or txr_profile_add_file for a
 * result file written earlier), and the three reference-filter rounds that open tax_profile (:796-825:
 * remove_matches_to_nonunique_refs :186-234, remove_low_confidence_references(3, 0.01) :269-282, filter_ref_associations
 * :289-462) run on the table.  The EM loop and the CAMI writers stay with the reference binary. */
typedef struct txr_profile txr_profile;
typedef struct
{
    uint64_t user_bin, seq_len;                    /* Species::user_bin, Species::seq_len (src/taxonomy/Species.hpp:14-21) */
    const char *accession_id, *taxid, *taxnames_string, *taxid_string;
} txr_profile_species;
typedef struct                                     /* the table in std::map order (by read id), hits in arrival order */
{
    uint64_t n_reads;
    const char *const *read_id;                    /* [n_reads]                                                        */
    const uint64_t *hit_begin;                     /* [n_reads+1]                                                      */
    const char *const *accession_id;               /* [n_hits] "-" = unclassified (taxonomy::Search_Result fields)     */
    const char *const *tax_id;
    const uint64_t *ref_len, *query_len, *query_hash_count, *query_hash_match;
    uint64_t n_taxa;                               /* references left after round 3 (filter_ref_associations' result)  */
} txr_profile_view;
const char *txr_profile_last_error(void);
int txr_profile_create(txr_profile **out);
void txr_profile_destroy(txr_profile *p);
/* read_ids[r] / read_len[r] describe read r of `result` (the id is cut at its first space, :125-126); only hits with
 * keep != 0 count (the result file holds no others); reads without hits enter as "-" (:129-133) */
int txr_profile_add_batch(txr_profile *p, const txr_result *result, const char *const *read_ids, const uint32_t *read_len,
                          const txr_profile_species *species, uint64_t n_species);
int txr_profile_add_file(txr_profile *p, const char *search_file);
/* runs the filter rounds not yet run, up to `rounds` (1..3) */
int txr_profile_filter(txr_profile *p, int rounds);
int txr_profile_get(txr_profile *p, txr_profile_view *out);
/* line format for tests and hand-over: "R <id> <n>", n x "H <accession> <taxid> <ref_len> <query_len> <hashes> <matches>",
 * then "T <accession> <ref_len>" per remaining reference and "P <accession> <taxid string> <taxname string>" (tab separated) */
int txr_profile_text(txr_profile *p, const char **text, uint64_t *len);

/* ---- kernel-level entry points for parity tests ---- */
/* kernel #1 alone: seq_to_syncmers / minimiser_hash for a batch.  hash_off[n_reads+1]; hashes of read r are
 * hashes[hash_off[r] .. hash_off[r+1]) (distinct, unordered, in syncmer mode; position order in k-mer mode).
 * dedup == 0 returns the raw emission list (syncmer mode).  Buffers are owned by the context. */
int txr_hash_batch(txr_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len,
                   uint64_t n_reads, int dedup, const uint64_t **hash_off, const uint64_t **hashes);
/* kernel #2 alone: interleaved_xor_filter::counting_agent().bulk_count(values) (call site hixf.hpp:307-309)
 * for one IXF of the uploaded index; counts[bins]. */
int txr_ixf_bulk_count(txr_ctx *ctx, uint64_t ixf_idx, const uint64_t *values, uint64_t n, uint32_t *counts);

#ifdef __cplusplus
}
#endif
#endif
