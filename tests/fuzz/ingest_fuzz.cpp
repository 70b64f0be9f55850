// ingest_fuzz.cpp -- in-process mutation fuzzer for the CLI's sequence-file ingest (taxor_b200/csrc/ingest.cpp), meant to
// be built with -fsanitize=address,undefined: valid FASTA/FASTQ text is mutated (byte flips, truncation, spliced
// garbage, CRLF, missing final newline), written to a file and pushed through BOTH paths -- the streaming RecordScanner
// and the mapped byte-range path (guess + exact scan + in-order acceptance).  Either path may reject the input with an
// exception; when both accept it they must agree record for record.  Any crash / sanitizer report fails the run.
// usage: ingest_fuzz <tmp file> <iterations> <seed>
#include "../../taxor_b200/csrc/ingest.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>

using namespace txr;

static std::string make_valid(std::mt19937_64 &rng)
{
    std::string s;
    const bool fastq = rng() & 1;
    const int n = 1 + (int)(rng() % 12);
    const char *eol = (rng() % 4 == 0) ? "\r\n" : "\n";
    for (int i = 0; i < n; ++i)
    {
        const size_t L = rng() % 5 == 0 ? 0 : rng() % 300;
        std::string seq(L, 'A'), qual(L, 'I');
        for (auto &c : seq)
            c = "ACGTNacgt"[rng() % 9];
        for (auto &c : qual)
            c = (char)(33 + rng() % 60);
        s += fastq ? "@" : ">";
        s += "r" + std::to_string(i) + " d";
        s += eol;
        const size_t width = (rng() % 3 == 0 && !fastq) ? 1 + rng() % 80 : 0;
        if (width)
            for (size_t a = 0; a < L; a += width)
                s += seq.substr(a, width) + eol;
        else
            s += seq + eol;
        if (fastq)
        {
            s += "+";
            s += eol;
            s += qual + eol;
        }
        if (rng() % 5 == 0)
            s += eol;
    }
    return s;
}

static void mutate(std::string &s, std::mt19937_64 &rng)
{
    const int kind = (int)(rng() % 6);
    if (s.empty())
        return;
    if (kind == 0)
        for (int i = 0; i < 3; ++i)
            s[rng() % s.size()] = (char)(rng() % 256);
    else if (kind == 1)
        s.resize(rng() % s.size());
    else if (kind == 2)
        s.insert(rng() % s.size(), std::string(rng() % 40, "@>+\n\r"[rng() % 5]));
    else if (kind == 3 && s.back() == '\n')
        s.pop_back();
    else if (kind == 4)
        s.erase(rng() % s.size(), rng() % 20);
    // kind 5: keep valid
}

// BGZF as bgzip writes it (one gzip member per block, block size in the 'BC' extra field)
static std::string to_bgzf(const std::string &text, std::mt19937_64 &rng)
{
    std::string out;
    const size_t block = 64 + rng() % 4000;
    for (size_t a = 0; a <= text.size(); a += block)
    {
        const size_t n = std::min(block, text.size() - a); // the last iteration may write the empty EOF block
        std::vector<unsigned char> comp(n + n / 10 + 64);
        z_stream zs{};
        deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = reinterpret_cast<unsigned char *>(const_cast<char *>(text.data() + a));
        zs.avail_in = (unsigned)n;
        zs.next_out = comp.data();
        zs.avail_out = (unsigned)comp.size();
        deflate(&zs, Z_FINISH);
        const size_t c = zs.total_out;
        deflateEnd(&zs);
        const unsigned bsize = (unsigned)(18 + c + 8 - 1);
        const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 255), (unsigned char)(bsize >> 8)};
        out.append(reinterpret_cast<const char *>(hdr), 18);
        out.append(reinterpret_cast<const char *>(comp.data()), c);
        const unsigned long crc = crc32(0, reinterpret_cast<const unsigned char *>(text.data() + a), (unsigned)n);
        for (int i = 0; i < 4; ++i)
            out.push_back((char)(crc >> (8 * i)));
        for (int i = 0; i < 4; ++i)
            out.push_back((char)(n >> (8 * i)));
        if (n == 0)
            break;
    }
    return out;
}

struct Rec
{
    std::string id, seq;
    bool operator==(const Rec &o) const { return id == o.id && seq == o.seq; }
};

static void collect(const char *base, const std::vector<RecordRef> &recs, std::vector<Rec> &out)
{
    std::string joined;
    for (const auto &r : recs)
    {
        Rec x;
        x.id.assign(base + r.id_off, r.id_len);
        if (r.single_line)
            x.seq.assign(base + r.seq_off, r.seq_len);
        else
        {
            join_record(base, r, joined);
            x.seq = joined;
        }
        if (x.seq.size() != r.seq_len)
            throw std::logic_error("seq_len disagrees with the joined sequence");
        out.push_back(std::move(x));
    }
}

int main(int argc, char **argv)
{
    if (argc < 4)
        return 2;
    const char *path = argv[1];
    const long iters = atol(argv[2]);
    std::mt19937_64 rng((uint64_t)atoll(argv[3]));
    long accepted = 0, rejected = 0;
    const bool verbose = getenv("FUZZ_VERBOSE") != nullptr;
    for (long it = 0; it < iters; ++it)
    {
        if (verbose)
            fprintf(stderr, "it %ld\n", it);
        std::string text = make_valid(rng);
        mutate(text, rng);
        if (rng() % 4 == 0)
        {
            // blocked gzip, sometimes with damaged bytes: the scanner must inflate it to the same records or reject it
            std::string z = to_bgzf(text, rng);
            const bool damaged = rng() % 2;
            if (damaged)
                mutate(z, rng);
            FILE *g = fopen(path, "wb");
            if (!g)
                return 3;
            fwrite(z.data(), 1, z.size(), g);
            fclose(g);
            std::vector<Rec> zr, pr;
            bool ok_z = true, ok_p = true;
            try
            {
                RecordScanner sc(path);
                std::vector<char> buf;
                std::vector<RecordRef> recs;
                const size_t target = 16 + rng() % 5000;
                while (sc.next(buf, recs, target))
                    collect(buf.data(), recs, zr);
            }
            catch (std::runtime_error const &)
            {
                ok_z = false;
            }
            if (!damaged)
            {
                g = fopen(path, "wb");
                fwrite(text.data(), 1, text.size(), g);
                fclose(g);
                try
                {
                    RecordScanner sc(path);
                    std::vector<char> buf;
                    std::vector<RecordRef> recs;
                    while (sc.next(buf, recs, 4096))
                        collect(buf.data(), recs, pr);
                }
                catch (std::runtime_error const &)
                {
                    ok_p = false;
                }
                if (ok_z != ok_p || (ok_z && !(zr == pr)))
                {
                    fprintf(stderr, "BGZF and plain text disagree at iteration %ld\n", it);
                    return 1;
                }
            }
            ok_z ? ++accepted : ++rejected;
            continue;
        }
        FILE *f = fopen(path, "wb");
        if (!f)
            return 3;
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);
        std::vector<Rec> a, b;
        bool ok_a = true, ok_b = true;
        try
        {
            RecordScanner sc(path);
            std::vector<char> buf;
            std::vector<RecordRef> recs;
            const size_t target = 16 + rng() % 400;
            while (sc.next(buf, recs, target))
                collect(buf.data(), recs, a);
        }
        catch (std::runtime_error const &)
        {
            ok_a = false;
        }
        try
        {
            MappedFile mf(path);
            if (!mf.ok())
                throw std::runtime_error("open");
            const char *data = mf.data();
            const size_t size = mf.size();
            size_t first = 0;
            while (first < size && (data[first] == '\n' || data[first] == '\r'))
                ++first;
            if (first < size)
            {
                if (data[first] != '>' && data[first] != '@')
                    throw std::runtime_error("start");
                const size_t seg = 8 + rng() % 300, n_seg = (size + seg - 1) / seg;
                std::vector<SegmentScan> segs(n_seg);
                for (size_t k = 0; k < n_seg; ++k)
                    scan_byte_range(data, size, first, data[first], k * seg, std::min(size, (k + 1) * seg), segs[k]);
                size_t expected = first;
                for (size_t k = 0; k < n_seg; ++k)
                {
                    expected = accept_byte_range(data, size, expected, std::min(size, (k + 1) * seg), segs[k]);
                    collect(data + segs[k].begin, segs[k].recs, b);
                }
            }
        }
        catch (std::runtime_error const &)
        {
            ok_b = false;
        }
        if (ok_a != ok_b || (ok_a && !(a == b)))
        {
            fprintf(stderr, "paths disagree at iteration %ld (stream %s %zu records, mapped %s %zu records)\n", it, ok_a ? "ok" : "rejects",
                    a.size(), ok_b ? "ok" : "rejects", b.size());
            return 1;
        }
        ok_a ? ++accepted : ++rejected;
    }
    printf("fuzz ok: %ld inputs, %ld accepted, %ld rejected\n", iters, accepted, rejected);
    return 0;
}
