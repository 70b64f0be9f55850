/*
 * ixf_ref.h -- ORACLE-side statement of the interleaved-XOR-filter probe arithmetic.  TEST INFRASTRUCTURE.
 *
 * *** PARITY UNPINNED ***  The real arithmetic lives in seqan3::interleaved_xor_filter<uint8_t> of the
 * un-vendored fork JensUweUlrich/seqan3 (GIT_TAG master, /root/reference/src/seqan/CMakeLists.txt.in:39-42),
 * which is not present in /root/reference and cannot be fetched.  What follows restates the only in-tree
 * statement of the same author's XOR filter:
 *     src/main/hashutil.hpp:50-61   SimpleMixSplit::murmur64 / operator()  (key + seed, fmix64)
 *     src/main/xorfilter.hpp:22-28  rotl64
 *     src/main/xorfilter.hpp:36-39  reduce  ((u64)hash * n >> 32)
 *     src/main/xorfilter.hpp:42-45  getHashFromHash (index*21 rotation, + index*blockLength)
 *     src/main/xorfilter.hpp:60-62  fingerprint  (hash ^ hash>>32, truncated)
 *     src/main/xorfilter.hpp:336-350 Contain
 *     src/main/xorfilter.hpp:67-68  arrayLength = 32 + 1.23*size ; blockLength = arrayLength/3
 * Everything that depends on the fork is confined to this header (oracle) and to
 * taxor_b200/csrc/ixf_arith.cuh (product); fix both in one place once the fork is visible.
 */
#ifndef TAXOR_ORACLE_IXF_REF_H
#define TAXOR_ORACLE_IXF_REF_H
#include <stdint.h>

static inline uint64_t ixfref_fmix64(uint64_t h)
{
    h ^= h >> 33;
    h *= UINT64_C(0xff51afd7ed558ccd);
    h ^= h >> 33;
    h *= UINT64_C(0xc4ceb9fe1a85ec53);
    h ^= h >> 33;
    return h;
}
static inline uint64_t ixfref_rotl64(uint64_t n, unsigned c)
{
    c &= 63u;
    return c ? (n << c) | (n >> (64u - c)) : n;
}
static inline uint32_t ixfref_reduce(uint32_t hash, uint32_t n)
{
    return (uint32_t)(((uint64_t)hash * n) >> 32);
}
static inline uint64_t ixfref_mix(uint64_t key, uint64_t seed) { return ixfref_fmix64(key + seed); }
static inline uint8_t  ixfref_fingerprint(uint64_t hash) { return (uint8_t)(hash ^ (hash >> 32)); }
static inline uint64_t ixfref_slot(uint64_t hash, int index, uint64_t seg_len)
{
    uint32_t r = (uint32_t)ixfref_rotl64(hash, (unsigned)index * 21u);
    return (uint64_t)ixfref_reduce(r, (uint32_t)seg_len) + (uint64_t)index * seg_len;
}
/* ---- descriptor-driven form (test-side twin of txr_ixf_scheme / taxor_b200/csrc/ixf_arith.cuh, written independently) ----
 * Until the fork is visible the arithmetic is a hypothesis, so it is data: which slot derivation, which seed mixing, which
 * fingerprint fold.  {0,0,0,21,42} is the prototype above.  slots == 1 is the 3-wise binary fuse filter as published by
 * Graf & Lemire (2022) and implemented in FastFilter's binaryfusefilter.h (binary_fuse8_contain):
 *     hi = (u64)(((u128)hash * SegmentCountLength) >> 64);  h0 = hi;  h1 = h0 + SegmentLength;  h2 = h1 + SegmentLength;
 *     h1 ^= (hash >> 18) & SegmentLengthMask;  h2 ^= hash & SegmentLengthMask;                                          */
typedef struct
{
    uint32_t slots;       /* 0 xor3 | 1 fuse3 */
    uint32_t mix;         /* 0 fmix64(key + seed) | 1 fmix64(key ^ seed) */
    uint32_t fingerprint; /* 0 (u8)(h ^ h>>32) | 1 (u8)h | 2 (u8)(h>>56) */
    uint32_t rot1, rot2;  /* xor3 rotations of segments 1, 2 */
} ixfref_scheme;
static inline uint64_t ixfref_mix_s(const ixfref_scheme *s, uint64_t key, uint64_t seed)
{
    return ixfref_fmix64(s->mix ? (key ^ seed) : (key + seed));
}
static inline uint8_t ixfref_fingerprint_s(const ixfref_scheme *s, uint64_t hash)
{
    switch (s->fingerprint)
    {
    case 1: return (uint8_t)hash;
    case 2: return (uint8_t)(hash >> 56);
    default: return (uint8_t)(hash ^ (hash >> 32));
    }
}
/* rows: slots per bin; seg_len: slots per segment (xor3) / SegmentLength (fuse3, power of two) */
static inline uint64_t ixfref_slot_s(const ixfref_scheme *s, uint64_t hash, int index, uint64_t seg_len, uint64_t rows)
{
    if (s->slots == 1)
    {
        const uint64_t count_len = rows - 2 * seg_len; /* SegmentCount * SegmentLength */
        const uint64_t h0 = (uint64_t)(((unsigned __int128)hash * count_len) >> 64);
        if (index == 0)
            return h0;
        if (index == 1)
            return (h0 + seg_len) ^ ((hash >> 18) & (seg_len - 1));
        return (h0 + 2 * seg_len) ^ (hash & (seg_len - 1));
    }
    const unsigned rot = index == 0 ? 0u : index == 1 ? s->rot1 : s->rot2;
    return (uint64_t)ixfref_reduce((uint32_t)ixfref_rotl64(hash, rot), (uint32_t)seg_len) + (uint64_t)index * seg_len;
}
/* slots per segment for a bin capacity (prototype formula, xorfilter.hpp:67-68) */
static inline uint64_t ixfref_seg_len(uint64_t max_bin_elements)
{
    uint64_t array_length = (uint64_t)(32 + 1.23 * (double)max_bin_elements);
    return array_length / 3;
}
#endif
