// ixf_arith.cuh -- the ONE place in the product that states the interleaved-XOR-filter probe arithmetic.
//
// *** PARITY UNPINNED *** The authoritative arithmetic is seqan3::interleaved_xor_filter<uint8_t> of the fork
// JensUweUlrich/seqan3 (master), which is not vendored in the reference tree.  This header follows the same
// author's in-tree prototype:  src/main/hashutil.hpp:50-61 (murmur64 finaliser of key+seed),
// src/main/xorfilter.hpp:22-45 (rotl64 / reduce / getHashFromHash), :60-62 (fingerprint), :336-350 (Contain),
// :67-68 (arrayLength = 32 + 1.23*size, blockLength = arrayLength/3).
// Used by the CUDA query kernel (device) and by the CPU synthetic-index builder (host).  If the fork turns
// out to differ, this file (and its independent test-side twin) is all that changes.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TXR_HD __host__ __device__ __forceinline__
#else
#define TXR_HD inline
#endif

namespace txr
{
TXR_HD uint64_t ixf_mix(uint64_t key, uint64_t seed)
{
    uint64_t h = key + seed;
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
TXR_HD uint32_t ixf_fingerprint(uint64_t h) { return (uint32_t)((h ^ (h >> 32)) & 0xffu); }
// (uint32)rotl64(h, 21*i) reduced onto [0, seg_len) by multiply-shift, plus the segment base
TXR_HD uint32_t ixf_reduce(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }
TXR_HD void ixf_slots(uint64_t h, uint32_t seg_len, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
    p0 = ixf_reduce((uint32_t)h, seg_len);
    p1 = ixf_reduce((uint32_t)((h << 21) | (h >> 43)), seg_len) + seg_len;
    p2 = ixf_reduce((uint32_t)((h << 42) | (h >> 22)), seg_len) + 2u * seg_len;
}
inline uint64_t ixf_seg_len_for(uint64_t max_bin_elements)
{
    uint64_t array_length = (uint64_t)(32 + 1.23 * (double)max_bin_elements);
    return array_length / 3;
}
// ankerl::unordered_dense v3.0.1 detail::wyhash::hash(uint64_t) (call site src/hashing/syncmer.cpp:73-77);
// third-party, restated from the published algorithm: lo64(x*C) ^ hi64(x*C), C = 0x9E3779B97F4A7C15.
TXR_HD uint64_t wyhash_u64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return (x * 0x9E3779B97F4A7C15ULL) ^ __umul64hi(x, 0x9E3779B97F4A7C15ULL);
#else
    __uint128_t r = (__uint128_t)x * 0x9E3779B97F4A7C15ULL;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
#endif
}
} // namespace txr
