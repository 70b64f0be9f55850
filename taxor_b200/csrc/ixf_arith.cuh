// ixf_arith.cuh -- the ONE place in the product that states the interleaved-XOR-filter probe arithmetic.
//
// *** PARITY UNPINNED *** The authoritative arithmetic is seqan3::interleaved_xor_filter<uint8_t> of the fork
// JensUweUlrich/seqan3 (master), which is not vendored in the reference tree and is not on the GPU boxes either
// (profiles/r2_fork_probe.txt).  Because it cannot be read, the arithmetic is a DESCRIPTOR (IxfScheme, the device form of
// txr_ixf_scheme in include/taxor_b200.h) chosen when an index is uploaded, not a compile-time fact:
//
//   default (all zeros but the rotations) = the same author's in-tree prototype:
//       src/main/hashutil.hpp:50-61 (murmur64 finaliser of key+seed), src/main/xorfilter.hpp:22-45 (rotl64 / reduce /
//       getHashFromHash), :60-62 (fingerprint), :336-350 (Contain), :67-68 (arrayLength = 32 + 1.23*size, blockLength =
//       arrayLength/3).  This path is compiled with constants (ixf_mix / ixf_slots / ixf_fingerprint below).
//   slots = FUSE3 = the 3-wise binary fuse filter of Graf & Lemire ("Binary Fuse Filters", 2022; FastFilter's
//       binaryfusefilter.h): main.cpp:22 of the reference shows the fork also ships an interleaved_binary_fuse_filter.
//       h0 = mulhi64(h, segment_count * segment_length), h1 = h0 + L ^ ((h >> 18) & (L-1)), h2 = h0 + 2L ^ (h & (L-1)).
//   mix / fingerprint / rotations: the small variations a re-implementation of the same filter could have made.
//
// Used by the CUDA query kernels (device), by the CPU synthetic-index builder (host) and by the .hixf reader's
// consistency checks.  The test side keeps its own, independently written statement of the same descriptor.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define TXR_HD __host__ __device__ __forceinline__
#else
#define TXR_HD inline
#endif

namespace txr
{
// ---- the prototype's arithmetic, constants folded (default scheme) ----
TXR_HD uint64_t ixf_fmix64(uint64_t h)
{
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
TXR_HD uint64_t ixf_mix(uint64_t key, uint64_t seed) { return ixf_fmix64(key + seed); }
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

// ---- the descriptor ----
enum : uint32_t { kIxfSlotsXor3 = 0, kIxfSlotsFuse3 = 1 };
enum : uint32_t { kIxfMixAddSeed = 0, kIxfMixXorSeed = 1 };
enum : uint32_t { kIxfFpFold32 = 0, kIxfFpLow8 = 1, kIxfFpHigh8 = 2 };
struct IxfScheme
{
    uint32_t slots;       // kIxfSlotsXor3: three equal segments, rotl + multiply-shift | kIxfSlotsFuse3: binary fuse window
    uint32_t mix;         // kIxfMixAddSeed: fmix64(key + seed) | kIxfMixXorSeed: fmix64(key ^ seed)
    uint32_t fingerprint; // kIxfFpFold32: (u8)(h ^ h >> 32) | kIxfFpLow8: (u8)h | kIxfFpHigh8: (u8)(h >> 56)
    uint32_t rot1, rot2;  // Xor3: left rotation of h before the reduce for segments 1 and 2 (prototype: 21, 42)
};
TXR_HD IxfScheme ixf_default_scheme() { return IxfScheme{kIxfSlotsXor3, kIxfMixAddSeed, kIxfFpFold32, 21u, 42u}; }
TXR_HD bool ixf_scheme_is_default(const IxfScheme &s)
{
    return s.slots == kIxfSlotsXor3 && s.mix == kIxfMixAddSeed && s.fingerprint == kIxfFpFold32 && s.rot1 == 21u && s.rot2 == 42u;
}
inline bool ixf_scheme_valid(const IxfScheme &s)
{
    return s.slots <= kIxfSlotsFuse3 && s.mix <= kIxfMixXorSeed && s.fingerprint <= kIxfFpHigh8 && s.rot1 < 64 && s.rot2 < 64;
}

TXR_HD uint64_t ixf_mix_g(uint64_t key, uint64_t seed, const IxfScheme &s)
{
    return ixf_fmix64(s.mix == kIxfMixXorSeed ? key ^ seed : key + seed);
}
TXR_HD uint32_t ixf_fingerprint_g(uint64_t h, const IxfScheme &s)
{
    const uint64_t f = s.fingerprint == kIxfFpFold32 ? h ^ (h >> 32) : s.fingerprint == kIxfFpLow8 ? h : h >> 56;
    return (uint32_t)(f & 0xffu);
}
TXR_HD uint64_t ixf_rotl64(uint64_t h, uint32_t c) { return c ? (h << c) | (h >> (64u - c)) : h; }
TXR_HD uint64_t ixf_mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((__uint128_t)a * b) >> 64);
#endif
}
// seg_len: slots per segment (Xor3) / segment length, a power of two (Fuse3); count_len: segment_count * seg_len (Fuse3)
TXR_HD void ixf_slots_g(uint64_t h, uint32_t seg_len, uint32_t count_len, const IxfScheme &s, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
    if (s.slots == kIxfSlotsFuse3)
    {
        const uint32_t mask = seg_len - 1u;
        p0 = (uint32_t)ixf_mulhi64(h, (uint64_t)count_len);
        p1 = (p0 + seg_len) ^ ((uint32_t)(h >> 18) & mask);
        p2 = (p0 + 2u * seg_len) ^ ((uint32_t)h & mask);
    }
    else
    {
        p0 = ixf_reduce((uint32_t)h, seg_len);
        p1 = ixf_reduce((uint32_t)ixf_rotl64(h, s.rot1), seg_len) + seg_len;
        p2 = ixf_reduce((uint32_t)ixf_rotl64(h, s.rot2), seg_len) + 2u * seg_len;
    }
}

// rows (slots) of one bin and their partition, from the bin capacity
struct IxfGeometry
{
    uint64_t seg_len;   // Xor3: rows / 3; Fuse3: segment length (power of two)
    uint64_t rows;      // Xor3: 3 * seg_len; Fuse3: (segment_count + 2) * seg_len
    uint64_t count_len; // Fuse3: segment_count * seg_len (the range of p0); Xor3: 0
};
inline IxfGeometry ixf_geometry_for(const IxfScheme &s, uint64_t max_bin_elements)
{
    IxfGeometry g{};
    if (s.slots == kIxfSlotsFuse3)
    {
        // binary_fuse8_allocate (FastFilter binaryfusefilter.h), arity 3
        const uint64_t size = max_bin_elements;
        uint64_t L = 4;
        if (size > 0)
        {
            L = 1ull << (int)std::floor(std::log((double)size) / std::log(3.33) + 2.25);
            if (L > 262144)
                L = 262144;
        }
        const double factor = size <= 1 ? 0.0 : std::fmax(1.125, 0.875 + 0.25 * std::log(1000000.0) / std::log((double)size));
        const uint64_t capacity = size <= 1 ? 0 : (uint64_t)std::llround((double)size * factor);
        const uint64_t init_count = (capacity + L - 1) / L; // may be < 2: clamped below
        uint64_t array_len = (init_count >= 2 ? init_count : 2) * L;
        uint64_t count = (array_len + L - 1) / L;
        count = count <= 2 ? 1 : count - 2;
        g.seg_len = L;
        g.count_len = count * L;
        g.rows = (count + 2) * L;
    }
    else
    {
        g.seg_len = ixf_seg_len_for(max_bin_elements);
        g.rows = 3 * g.seg_len;
        g.count_len = 0;
    }
    return g;
}
// is (seg_len, rows) a geometry this scheme can address?  Returns count_len through `count_len`.
inline bool ixf_geometry_ok(const IxfScheme &s, uint64_t seg_len, uint64_t rows, uint64_t &count_len)
{
    count_len = 0;
    if (seg_len == 0 || seg_len >= (1ull << 31) || rows >= (1ull << 32))
        return false;
    if (s.slots == kIxfSlotsFuse3)
    {
        if ((seg_len & (seg_len - 1)) != 0 || rows % seg_len != 0 || rows < 3 * seg_len)
            return false;
        count_len = rows - 2 * seg_len;
        return true;
    }
    return rows == 3 * seg_len;
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
