// pack_simd.cpp -- the fast lane of txr_pack_2bit (read ingest): 32 plain A/C/G/T bases (either case) ->
// one 64-bit word, first base most significant, with AVX2 + BMI2 when the CPU has them (checked at run time; the build itself
// stays baseline x86-64).  Anything else in a 32-base block -- IUPAC codes, which collapse like seqan3::dna4, or an illegal
// character -- sends that block, and only that block, back to the table-driven scalar loop in engine.cu.
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace txr
{
#if defined(__x86_64__)
// code = ((c >> 1) ^ (c >> 2)) & 3 maps A,a -> 0  C,c -> 1  G,g -> 2  T,t -> 3
__attribute__((target("avx2,bmi2"))) static size_t pack_words_avx2(const char *ascii, size_t n_words, uint64_t *dst)
{
    const __m256i up = _mm256_set1_epi8((char)0xDF), a = _mm256_set1_epi8('A'), c = _mm256_set1_epi8('C'), g = _mm256_set1_epi8('G'),
                  t = _mm256_set1_epi8('T');
    for (size_t w = 0; w < n_words; ++w)
    {
        const char *p = ascii + 32 * w;
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
        const __m256i u = _mm256_and_si256(x, up);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, a), _mm256_cmpeq_epi8(u, c)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, g), _mm256_cmpeq_epi8(u, t)));
        if ((uint32_t)_mm256_movemask_epi8(ok) != 0xffffffffu)
            return w; // something other than ACGT/acgt: the caller's scalar loop decides what it is
        uint64_t q[4];
        memcpy(q, p, 32);
        uint64_t out = 0;
        for (int i = 0; i < 4; ++i)
        {
            const uint64_t v = __builtin_bswap64(q[i]);                       // first base of the 8 into the top byte
            const uint64_t codes = ((v >> 1) ^ (v >> 2)) & 0x0303030303030303ull;
            out = (out << 16) | _pext_u64(codes, 0x0303030303030303ull);     // 8 x 2 bits, top byte first
        }
        dst[w] = out;
    }
    return n_words;
}
#endif

// packs full 32-base words from `ascii` while they hold only A/C/G/T (either case); returns how many words were written
size_t pack_plain_words(const char *ascii, size_t n_words, uint64_t *dst)
{
#if defined(__x86_64__)
    static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    if (fast)
        return pack_words_avx2(ascii, n_words, dst);
#endif
    (void)ascii;
    (void)n_words;
    (void)dst;
    return 0;
}
} // namespace txr
