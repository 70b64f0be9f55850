// inflate_fuzz.cpp -- the word-at-a-time inflater of the read ingest (taxor_b200/csrc/inflate_fast.cpp) against zlib, meant to be
// built with -fsanitize=address,undefined.
//   1. round trips: data of several kinds (random bytes, DNA, FASTQ-like, runs, text) deflated by zlib at every level and
//      strategy (stored, fixed and dynamic blocks, RLE = distance-1 matches, Huffman-only = no distance tree), with flush
//      points, as one or several gzip members, read back through GzipStream in pieces of random size: byte-identical;
//   2. the whole-buffer form on raw DEFLATE streams of 0 .. 70 000 bytes (the BGZF block path), exact and wrong output sizes;
//   3. damage: flipped, dropped, inserted bytes and truncation -- an exception or (only if the CRC still matches) the
//      original bytes, never a crash or a read outside the buffers;
//   4. the carry-less-multiply CRC-32 against zlib's on random lengths and alignments.
// usage: inflate_fuzz <iterations> <seed>
#include "../../taxor_b200/csrc/gzip_parallel.hpp"
#include "../../taxor_b200/csrc/inflate_fast.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>

using namespace txr;
using Bytes = std::vector<uint8_t>;
static bool same(const uint8_t *a, const uint8_t *b, size_t n) { return n == 0 || memcmp(a, b, n) == 0; }

static Bytes make_data(std::mt19937_64 &rng)
{
    static const size_t sizes[] = {0, 1, 2, 7, 100, 257, 258, 259, 4000, 32767, 32768, 32769, 70000, 300000, 1500000};
    size_t n = sizes[rng() % (sizeof sizes / sizeof *sizes)];
    if (rng() % 3 == 0)
        n = rng() % 200000;
    Bytes d(n);
    switch (rng() % 7)
    {
    case 0: // incompressible
        for (auto &c : d)
            c = (uint8_t)rng();
        break;
    case 1: // DNA
        for (auto &c : d)
            c = (uint8_t) "ACGT"[rng() % 4];
        break;
    case 2: // FASTQ-like: header, bases, +, skewed qualities
    {
        size_t i = 0;
        while (i < n)
        {
            char hdr[64];
            const int h = snprintf(hdr, sizeof hdr, "@read%llu ch=%u\n", (unsigned long long)(rng() % 1000000), (unsigned)(rng() % 512));
            for (int k = 0; k < h && i < n; ++k)
                d[i++] = (uint8_t)hdr[k];
            const size_t L = 50 + rng() % 3000;
            for (size_t k = 0; k < L && i < n; ++k)
                d[i++] = (uint8_t) "ACGT"[rng() % 4];
            for (const char *s = "\n+\n"; *s && i < n; ++s)
                d[i++] = (uint8_t)*s;
            for (size_t k = 0; k < L && i < n; ++k)
            {
                unsigned q = 0;
                while (q < 40 && rng() % 8)
                    ++q;
                d[i++] = (uint8_t)(35 + q);
            }
            if (i < n)
                d[i++] = '\n';
        }
        break;
    }
    case 3: // long runs (distance 1, maximal lengths)
    {
        size_t i = 0;
        while (i < n)
        {
            const uint8_t c = (uint8_t)rng();
            const size_t run = 1 + rng() % 2000;
            for (size_t k = 0; k < run && i < n; ++k)
                d[i++] = c;
        }
        break;
    }
    case 4: // short periods (distances 2..15)
    {
        const size_t period = 2 + rng() % 14;
        for (size_t i = 0; i < n; ++i)
            d[i] = i < period ? (uint8_t)rng() : (rng() % 200 ? d[i - period] : (uint8_t)rng());
        break;
    }
    case 5: // many distinct symbols with a long tail: deep Huffman trees (codes beyond the primary table)
        for (auto &c : d)
        {
            unsigned v = 0;
            while (v < 255 && rng() % 16)
                ++v;
            c = (uint8_t)(v * 37);
        }
        break;
    default: // far matches: copies from anywhere in the last 32 KiB
        for (size_t i = 0; i < n;)
        {
            if (i > 100 && rng() % 2)
            {
                const size_t dist = 1 + rng() % std::min<size_t>(i, 32768), len = 3 + rng() % 300;
                for (size_t k = 0; k < len && i < n; ++k, ++i)
                    d[i] = d[i - dist];
            }
            else
                d[i++] = (uint8_t)rng();
        }
    }
    return d;
}

static Bytes deflate_with(const Bytes &data, int level, int strategy, int window_bits, std::mt19937_64 &rng, bool flushes)
{
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, window_bits, 1 + (int)(rng() % 9), strategy) != Z_OK)
        throw std::runtime_error("deflateInit2");
    Bytes out(deflateBound(&zs, (uLong)data.size()) + 4096 + data.size() / 8);
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    size_t pos = 0;
    while (pos < data.size() && flushes)
    {
        const size_t piece = std::min<size_t>(data.size() - pos, 1 + rng() % 50000);
        zs.next_in = const_cast<Bytef *>(data.data() + pos);
        zs.avail_in = (uInt)piece;
        static const int kinds[] = {Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_PARTIAL_FLUSH, Z_BLOCK};
        if (deflate(&zs, kinds[rng() % 5]) != Z_OK || zs.avail_in)
            throw std::runtime_error("deflate");
        pos += piece;
    }
    zs.next_in = const_cast<Bytef *>(data.data() + pos);
    zs.avail_in = (uInt)(data.size() - pos);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END)
        throw std::runtime_error("deflate finish");
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

static bool read_all(const Bytes &gz, std::mt19937_64 &rng, Bytes &out, std::string &err)
{
    out.clear();
    try
    {
        GzipStream gs(gz.data(), gz.size());
        Bytes piece;
        for (;;)
        {
            static const size_t caps[] = {1, 2, 100, 4096, 65536, 1 << 20, 3 << 20};
            piece.resize(caps[rng() % 7]);
            const size_t n = gs.read(piece.data(), piece.size());
            out.insert(out.end(), piece.begin(), piece.begin() + (long)n);
            if (n < piece.size())
                break;
        }
        return true;
    }
    catch (std::runtime_error const &e)
    {
        err = e.what();
        return false;
    }
}

// the same through the multi-threaded reader, with pieces small enough that even these inputs are cut many times
static bool read_all_parallel(const Bytes &gz, std::mt19937_64 &rng, Bytes &out, std::string &err, ParallelGzip::Stats *stats = nullptr)
{
    out.clear();
    try
    {
        static const size_t piece_sizes[] = {1024, 3000, 10000, 65536, 300000};
        ParallelGzip pg(gz.data(), gz.size(), 1 + (unsigned)(rng() % 4), piece_sizes[rng() % 5]);
        Bytes piece;
        for (;;)
        {
            static const size_t caps[] = {1, 100, 4096, 65536, 1 << 20, 3 << 20};
            piece.resize(caps[rng() % 6]);
            const size_t n = pg.read(piece.data(), piece.size());
            out.insert(out.end(), piece.begin(), piece.begin() + (long)n);
            if (n < piece.size())
                break;
        }
        if (stats)
            *stats = pg.stats();
        return true;
    }
    catch (std::runtime_error const &e)
    {
        err = e.what();
        return false;
    }
}

int main(int argc, char **argv)
{
    if (argc < 3)
        return 2;
    const long iters = atol(argv[1]);
    std::mt19937_64 rng((uint64_t)atoll(argv[2]));
    long n_round = 0, n_damaged = 0, n_rejected = 0, n_raw = 0, par_bytes = 0, seq_bytes = 0;
    static const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
    for (long it = 0; it < iters; ++it)
    {
        // 1. round trips through GzipStream
        Bytes whole, gz;
        const int members = 1 + (int)(rng() % 3 == 0 ? rng() % 3 : 0);
        for (int m = 0; m < members; ++m)
        {
            const Bytes d = make_data(rng);
            const Bytes z = deflate_with(d, (int)(rng() % 10), strategies[rng() % 5], 15 + 16, rng, rng() % 2);
            whole.insert(whole.end(), d.begin(), d.end());
            gz.insert(gz.end(), z.begin(), z.end());
        }
        if (rng() % 4 == 0) // bytes behind the last member that are not another member are ignored
            gz.insert(gz.end(), 1 + rng() % 20, 0);
        Bytes back;
        std::string err;
        if (!read_all(gz, rng, back, err))
        {
            printf("iteration %ld: a valid stream was rejected: %s\n", it, err.c_str());
            return 1;
        }
        if (back != whole)
        {
            printf("iteration %ld: %zu bytes in, %zu bytes out, contents differ\n", it, whole.size(), back.size());
            return 1;
        }
        ++n_round;
        {
            ParallelGzip::Stats st;
            if (!read_all_parallel(gz, rng, back, err, &st))
            {
                printf("iteration %ld: the parallel reader rejected a valid stream: %s\n", it, err.c_str());
                return 1;
            }
            if (back != whole)
            {
                printf("iteration %ld: parallel reader: %zu bytes in, %zu bytes out, contents differ\n", it, whole.size(), back.size());
                return 1;
            }
            par_bytes += (long)st.parallel_bytes;
            seq_bytes += (long)st.sequential_bytes;
        }

        // 2. the whole-buffer form on a raw stream
        {
            Bytes d = make_data(rng);
            if (d.size() > 70000)
                d.resize(70000);
            const Bytes z = deflate_with(d, (int)(rng() % 10), strategies[rng() % 5], -15, rng, rng() % 2);
            Bytes out(d.size() + 16, 0xAA);
            if (!inflate_raw_exact(z.data(), z.size(), out.data(), d.size()) || !same(out.data(), d.data(), d.size()))
            {
                printf("iteration %ld: raw stream of %zu bytes not reproduced\n", it, d.size());
                return 1;
            }
            for (size_t k = d.size(); k < out.size(); ++k)
                if (out[k] != 0xAA)
                {
                    printf("iteration %ld: wrote behind the end of the output\n", it);
                    return 1;
                }
            if (!d.empty() && inflate_raw_exact(z.data(), z.size(), out.data(), d.size() - 1))
            {
                printf("iteration %ld: a too small output was accepted\n", it);
                return 1;
            }
            if (inflate_raw_exact(z.data(), z.size(), out.data(), d.size() + 1))
            {
                printf("iteration %ld: a too large output was accepted\n", it);
                return 1;
            }
            if (z.size() > 1 && inflate_raw_exact(z.data(), z.size() - 1 - rng() % std::min<size_t>(z.size() - 1, 8), out.data(), d.size()))
            {
                // cutting the last bytes may only remove padding bits of the final byte's successor -- a stream that still
                // decodes completely is fine, anything else must have been refused; verify the bytes
                if (!same(out.data(), d.data(), d.size()))
                {
                    printf("iteration %ld: a truncated raw stream decoded to different bytes\n", it);
                    return 1;
                }
            }
            ++n_raw;
        }

        // 3. damage
        for (int k = 0; k < 4 && !gz.empty(); ++k)
        {
            Bytes bad = gz;
            switch (rng() % 5)
            {
            case 0:
                for (int j = 0; j < 1 + (int)(rng() % 3); ++j)
                    bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
                break;
            case 1:
                bad.resize(rng() % bad.size());
                break;
            case 2:
                bad.erase(bad.begin() + (long)(rng() % bad.size()));
                break;
            case 3:
                bad.insert(bad.begin() + (long)(rng() % bad.size()), (uint8_t)rng());
                break;
            default:
                for (size_t j = rng() % bad.size(), e = std::min(bad.size(), j + 1 + rng() % 64); j < e; ++j)
                    bad[j] = (uint8_t)rng();
            }
            {
                // the parallel reader on the same damaged bytes: it may differ from the serial one only in WHERE it notices
                Bytes got2;
                std::string err2;
                if (read_all_parallel(bad, rng, got2, err2))
                {
                    if (got2.size() > whole.size() || !same(got2.data(), whole.data(), got2.size()))
                    {
                        printf("iteration %ld: parallel reader accepted damaged input with different contents\n", it);
                        return 1;
                    }
                }
                else if (err2.find("corrupt or truncated") == std::string::npos)
                {
                    printf("iteration %ld: parallel reader, unexpected error text: %s\n", it, err2.c_str());
                    return 1;
                }
            }
            Bytes got;
            ++n_damaged;
            if (!read_all(bad, rng, got, err))
            {
                ++n_rejected;
                if (err.find("corrupt or truncated") == std::string::npos)
                {
                    printf("iteration %ld: unexpected error text: %s\n", it, err.c_str());
                    return 1;
                }
            }
            // accepted: the damage hit a header field that is not checked (time stamp, OS, file name), bytes behind the last
            // member, or produced a stream whose CRC and length still match -- then the bytes are the original ones or a prefix
            // of whole members
            else if (got.size() > whole.size() || !same(got.data(), whole.data(), got.size()))
            {
                printf("iteration %ld: damaged input accepted with different contents (%zu vs %zu bytes)\n", it, got.size(), whole.size());
                return 1;
            }
        }

        // 4. CRC
        {
            const size_t n = rng() % 5000, off = rng() % 16;
            Bytes d(n + off);
            for (auto &c : d)
                c = (uint8_t)rng();
            const uint32_t seed = (uint32_t)rng();
            if (crc32_fast(seed, d.data() + off, n) != (uint32_t)crc32(seed, d.data() + off, (uInt)n))
            {
                printf("iteration %ld: CRC mismatch at %zu bytes\n", it, n);
                return 1;
            }
        }
    }
    printf("fuzz ok: %ld round trips (parallel reader: %ld bytes from pieces, %ld decoded sequentially), %ld raw streams, %ld damaged inputs (%ld rejected)\n",
           n_round, par_bytes, seq_bytes, n_raw, n_damaged, n_rejected);
    return 0;
}
