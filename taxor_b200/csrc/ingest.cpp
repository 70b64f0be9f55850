#include "ingest.hpp"
#include "gzip_parallel.hpp"
#include "inflate_fast.hpp"

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstring>
#include <dlfcn.h>
#include <fcntl.h>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace txr
{
namespace
{
// BGZF block at p (n bytes left): gzip member with FEXTRA whose 'B','C' subfield holds the block size - 1.
// Returns the block size, 0 if this is not a BGZF block header.
size_t bgzf_block_size(const unsigned char *p, size_t n)
{
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4))
        return 0;
    const size_t xlen = p[10] | (size_t)p[11] << 8;
    if (12 + xlen > n)
        return 0;
    for (size_t q = 12; q + 4 <= 12 + xlen;)
    {
        const size_t slen = p[q + 2] | (size_t)p[q + 3] << 8;
        if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen)
        {
            const size_t bsize = (p[q + 4] | (size_t)p[q + 5] << 8) + 1;
            return bsize >= 12 + xlen + 8 && bsize <= n ? bsize : 0;
        }
        q += 4 + slen;
    }
    return 0;
}
} // namespace

// ---- bzip2 through libbz2.so.1.0, bound at run time ----
namespace
{
struct BzStream // bz_stream of bzlib.h (libbz2 1.0.x; the ABI has been frozen since 1.0.0)
{
    char *next_in;
    unsigned int avail_in, total_in_lo32, total_in_hi32;
    char *next_out;
    unsigned int avail_out, total_out_lo32, total_out_hi32;
    void *state;
    void *(*bzalloc)(void *, int, int);
    void (*bzfree)(void *, void *);
    void *opaque;
};
struct BzApi
{
    int (*init)(BzStream *, int, int){nullptr};
    int (*run)(BzStream *){nullptr};
    int (*end)(BzStream *){nullptr};
    bool ok{false};
};
const BzApi &bz_api()
{
    static const BzApi api = [] {
        BzApi a;
        void *h = nullptr;
        for (const char *name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"})
            if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL)))
                break;
        if (h)
        {
            a.init = reinterpret_cast<int (*)(BzStream *, int, int)>(dlsym(h, "BZ2_bzDecompressInit"));
            a.run = reinterpret_cast<int (*)(BzStream *)>(dlsym(h, "BZ2_bzDecompress"));
            a.end = reinterpret_cast<int (*)(BzStream *)>(dlsym(h, "BZ2_bzDecompressEnd"));
            a.ok = a.init && a.run && a.end;
        }
        return a;
    }();
    return api;
}
struct BzState
{
    BzStream zs{};
    bool open{false};
    std::vector<char> in;
    BzState() : in(1 << 20) {}
};
constexpr int kBzOk = 0, kBzStreamEnd = 4;
} // namespace

RecordScanner::RecordScanner(const std::string &path)
{
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0)
        return;
    unsigned char magic[18] = {0};
    const ssize_t n = ::pread(fd_, magic, sizeof magic, 0);
    if (n >= 4 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h' && magic[3] >= '1' && magic[3] <= '9')
    {
        if (!bz_api().ok)
        {
            open_error_ = "bzip2 input needs libbz2.so.1.0 at run time and it could not be loaded; decompress the file (bzip2 -d) or "
                          "recompress it with gzip / bgzip";
            ::close(fd_);
            fd_ = -1;
            return;
        }
        bz_ = new BzState;
        return; // fd_ stays open: fill_bz2 reads the compressed bytes from it
    }
    struct stat st;
    if (n == 18 && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && st.st_size >= 28)
    {
        // blocked gzip: look at the first header only (BSIZE must fit the file), then map the whole file
        unsigned char probe[18];
        memcpy(probe, magic, 18);
        const size_t xlen = probe[10] | (size_t)probe[11] << 8;
        if (probe[0] == 0x1f && probe[1] == 0x8b && probe[2] == 8 && (probe[3] & 4) && xlen >= 6 && probe[12] == 'B' && probe[13] == 'C')
        {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m != MAP_FAILED && bgzf_block_size(static_cast<const unsigned char *>(m), (size_t)st.st_size))
            {
                bgzf_data_ = static_cast<const unsigned char *>(m);
                bgzf_size_ = (size_t)st.st_size;
                madvise(m, bgzf_size_, MADV_SEQUENTIAL);
                ::close(fd_);
                fd_ = -1;
                return;
            }
            if (m != MAP_FAILED)
                munmap(m, (size_t)st.st_size);
        }
    }
    if (n >= 2 && magic[0] == 0x1f && magic[1] == 0x8b && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
        !(getenv("TAXOR_GZIP") && !strcmp(getenv("TAXOR_GZIP"), "zlib")))
    {
        // gzip in a regular file: map it and decode with the word-at-a-time inflater
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m != MAP_FAILED)
        {
            gzmap_ = static_cast<const unsigned char *>(m);
            gzmap_size_ = (size_t)st.st_size;
            madvise(m, gzmap_size_, MADV_SEQUENTIAL);
            // one stream, several threads: pieces of the file are inflated speculatively and chained (gzip_parallel.hpp);
            // TAXOR_GZIP=serial keeps the one-thread decoder
            unsigned threads = 1;
#ifdef _OPENMP
            threads = (unsigned)std::max(1, omp_get_max_threads());
#endif
            const char *mode = getenv("TAXOR_GZIP");
            gz_parallel_ = threads > 1 && gzmap_size_ >= (size_t(4) << 20) && !(mode && !strcmp(mode, "serial"));
            if (gz_parallel_)
                gzs_ = new ParallelGzip(gzmap_, gzmap_size_, threads);
            else
                gzs_ = new GzipStream(gzmap_, gzmap_size_);
            ::close(fd_);
            fd_ = -1;
            return;
        }
    }
    if (n >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) // gzip that cannot be mapped: inflate through zlib; everything else is read() directly
    {
        gz_ = gzdopen(fd_, "rb");
        if (!gz_)
        {
            ::close(fd_);
            fd_ = -1;
            return;
        }
        fd_ = -1; // owned by gz_ now
        gzbuffer(gz_, 1 << 20);
    }
#ifdef POSIX_FADV_SEQUENTIAL
    else
        posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
}

RecordScanner::~RecordScanner()
{
    if (bz_)
    {
        BzState *b = static_cast<BzState *>(bz_);
        if (b->open)
            bz_api().end(&b->zs);
        delete b;
    }
    if (bgzf_data_)
        munmap(const_cast<unsigned char *>(bgzf_data_), bgzf_size_);
    if (gz_parallel_)
        delete static_cast<ParallelGzip *>(gzs_);
    else
        delete static_cast<GzipStream *>(gzs_);
    if (gzmap_)
        munmap(const_cast<unsigned char *>(gzmap_), gzmap_size_);
    if (gz_)
        gzclose(gz_);
    if (fd_ >= 0)
        ::close(fd_);
}

// Inflates as many whole BGZF blocks as fit into [dst, dst+cap) in parallel; a block that does not fit is inflated into
// bgzf_rest_ and handed out piecewise.
size_t RecordScanner::fill_bgzf(char *dst, size_t cap)
{
    size_t got = 0;
    while (got < cap)
    {
        if (bgzf_rest_pos_ < bgzf_rest_.size())
        {
            const size_t n = std::min(cap - got, bgzf_rest_.size() - bgzf_rest_pos_);
            memcpy(dst + got, bgzf_rest_.data() + bgzf_rest_pos_, n);
            bgzf_rest_pos_ += n;
            got += n;
            continue;
        }
        if (bgzf_pos_ >= bgzf_size_)
        {
            eof_ = true;
            break;
        }
        struct Block
        {
            size_t in, in_len, out, out_len;
            uint32_t crc;
        };
        std::vector<Block> blocks;
        size_t pos = bgzf_pos_, out = got;
        while (pos < bgzf_size_)
        {
            const unsigned char *p = bgzf_data_ + pos;
            const size_t bsize = bgzf_block_size(p, bgzf_size_ - pos);
            if (!bsize)
                throw std::runtime_error("read error (corrupt BGZF block header)");
            const size_t xlen = p[10] | (size_t)p[11] << 8, hdr = 12 + xlen;
            const unsigned char *t = p + bsize - 8;
            const uint32_t crc = t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
            const size_t isize = t[4] | (size_t)t[5] << 8 | (size_t)t[6] << 16 | (size_t)t[7] << 24;
            if (isize > (1u << 16)) // a BGZF block holds at most 64 KiB: a larger claim is damage, not a reason to allocate
                throw std::runtime_error("read error (corrupt BGZF block size)");
            if (isize > cap - out)
                break;
            blocks.push_back(Block{pos + hdr, bsize - hdr - 8, out, isize, crc});
            out += isize;
            pos += bsize;
        }
        if (blocks.empty())
        {
            // the next block is larger than what is left of the caller's buffer: inflate it aside
            const unsigned char *p = bgzf_data_ + pos;
            const size_t bsize = bgzf_block_size(p, bgzf_size_ - pos);
            const size_t xlen = p[10] | (size_t)p[11] << 8, hdr = 12 + xlen;
            const unsigned char *t = p + bsize - 8;
            const size_t isize = t[4] | (size_t)t[5] << 8 | (size_t)t[6] << 16 | (size_t)t[7] << 24;
            bgzf_rest_.resize(isize);
            bgzf_rest_pos_ = 0;
            z_stream zs{};
            if (inflateInit2(&zs, -15) != Z_OK)
                throw std::runtime_error("zlib initialisation failed");
            zs.next_in = const_cast<unsigned char *>(p + hdr);
            zs.avail_in = (unsigned)(bsize - hdr - 8);
            zs.next_out = reinterpret_cast<unsigned char *>(bgzf_rest_.data());
            zs.avail_out = (unsigned)isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            const uint32_t crc = t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
            if (rc != Z_STREAM_END || zs.total_out != isize ||
                crc32(0, reinterpret_cast<const unsigned char *>(bgzf_rest_.data()), (unsigned)isize) != crc)
                throw std::runtime_error("read error (corrupt BGZF block)");
            bgzf_pos_ = pos + bsize;
            continue;
        }
        int bad = 0;
        static const bool use_zlib = getenv("TAXOR_GZIP") && !strcmp(getenv("TAXOR_GZIP"), "zlib");
#pragma omp parallel for schedule(dynamic, 4) if (blocks.size() > 8)
        for (long i = 0; i < (long)blocks.size(); ++i)
        {
            const Block &b = blocks[(size_t)i];
            uint8_t *const o = reinterpret_cast<uint8_t *>(dst + b.out);
            bool good;
            if (use_zlib) // TAXOR_GZIP=zlib: the round-1 path, kept for A/B runs
            {
                z_stream zs{};
                good = inflateInit2(&zs, -15) == Z_OK;
                if (good)
                {
                    zs.next_in = const_cast<unsigned char *>(bgzf_data_ + b.in);
                    zs.avail_in = (unsigned)b.in_len;
                    zs.next_out = o;
                    zs.avail_out = (unsigned)b.out_len;
                    const int rc = inflate(&zs, Z_FINISH);
                    inflateEnd(&zs);
                    good = rc == Z_STREAM_END && zs.total_out == b.out_len && crc32(0, o, (unsigned)b.out_len) == b.crc;
                }
            }
            else
                good = inflate_raw_exact(bgzf_data_ + b.in, b.in_len, o, b.out_len) && crc32_fast(0, o, b.out_len) == b.crc;
            if (!good)
            {
#pragma omp atomic write
                bad = 1;
            }
        }
        if (bad)
            throw std::runtime_error("read error (corrupt BGZF block)");
        bgzf_pos_ = pos;
        got = out;
        if (pos < bgzf_size_)
            break; // buffer (nearly) full: the next block does not fit
    }
    return got;
}

// one bzip2 stream after the other (pbzip2 / `cat a.bz2 b.bz2` files are concatenations), single-threaded like zlib's
size_t RecordScanner::fill_bz2(char *dst, size_t cap)
{
    BzState &b = *static_cast<BzState *>(bz_);
    const BzApi &api = bz_api();
    size_t got = 0;
    while (got < cap && !eof_)
    {
        if (b.zs.avail_in == 0)
        {
            const ssize_t n = ::read(fd_, b.in.data(), b.in.size());
            if (n < 0)
                throw std::runtime_error("read error");
            if (n == 0)
            {
                if (b.open)
                    throw std::runtime_error("truncated bzip2 stream");
                eof_ = true;
                break;
            }
            b.zs.next_in = b.in.data();
            b.zs.avail_in = (unsigned)n;
        }
        if (!b.open)
        {
            char *keep_in = b.zs.next_in;
            const unsigned keep_avail = b.zs.avail_in;
            b.zs = BzStream{};
            b.zs.next_in = keep_in;
            b.zs.avail_in = keep_avail;
            if (api.init(&b.zs, 0, 0) != kBzOk)
                throw std::runtime_error("bzip2 initialisation failed");
            b.open = true;
        }
        b.zs.next_out = dst + got;
        b.zs.avail_out = (unsigned)std::min<size_t>(cap - got, 1u << 30);
        const unsigned before = b.zs.avail_out;
        const int rc = api.run(&b.zs);
        got += before - b.zs.avail_out;
        if (rc == kBzStreamEnd)
        {
            api.end(&b.zs);
            b.open = false; // another stream may follow
        }
        else if (rc != kBzOk)
            throw std::runtime_error("read error (corrupt bzip2 stream?)");
    }
    return got;
}

size_t RecordScanner::fill(char *dst, size_t cap)
{
    if (bgzf_data_)
        return fill_bgzf(dst, cap);
    if (bz_)
        return fill_bz2(dst, cap);
    if (gzs_)
    {
        const size_t n = gz_parallel_ ? static_cast<ParallelGzip *>(gzs_)->read(reinterpret_cast<uint8_t *>(dst), cap)
                                      : static_cast<GzipStream *>(gzs_)->read(reinterpret_cast<uint8_t *>(dst), cap);
        if (n < cap)
            eof_ = true;
        return n;
    }
    size_t got = 0;
    while (got < cap && !eof_)
    {
        long n;
        if (gz_)
        {
            n = gzread(gz_, dst + got, (unsigned)std::min<size_t>(cap - got, 1u << 30));
            if (n < 0)
                throw std::runtime_error("read error (corrupt gzip stream?)");
        }
        else
        {
            n = ::read(fd_, dst + got, cap - got);
            if (n < 0)
                throw std::runtime_error("read error");
        }
        if (n == 0)
            eof_ = true;
        got += (size_t)n;
    }
    return got;
}

namespace
{
inline size_t line_len(const char *b, const char *e) // without one trailing '\r'
{
    return (size_t)(e - b) - ((e > b && e[-1] == '\r') ? 1 : 0);
}

// Scans one record starting at data[pos] (which is '>' or '@').  Returns false when the record is not complete
// in [pos, n) and more data may follow.
bool scan_record(const char *data, size_t n, size_t pos, bool at_eof, RecordRef &r, size_t &next_pos)
{
    const char marker = data[pos];
    const char *end = data + n;
    const char *p = data + pos;
    const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
    if (!nl && !at_eof)
        return false;
    const char *hdr_end = nl ? nl : end;
    const char *id = p + 1;
    if (marker == '>') // SeqAn3's FASTA reader skips the blanks between '>' and the id ("> id" and ">id" are the same record)
        while (id < hdr_end && (*id == ' ' || *id == '\t'))
            ++id;
    r.fasta = marker == '>';
    r.id_off = (uint32_t)(id - data);
    r.id_len = (uint32_t)line_len(id, hdr_end);
    p = nl ? nl + 1 : end;
    r.seq_off = (uint32_t)(p - data);
    size_t seq_len = 0, lines = 0;
    const char *seq_end = p;
    const char stop = marker == '>' ? '>' : '+';
    while (true)
    {
        if (p >= end)
        {
            if (!at_eof)
                return false;
            if (marker == '@')
                throw std::runtime_error("FASTQ record without a '+' line");
            break;
        }
        if (*p == stop)
            break;
        const char *e = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
        if (!e && !at_eof)
            return false;
        const char *le = e ? e : end;
        const size_t ll = line_len(p, le);
        if (ll)
        {
            if (lines == 0)
                r.seq_off = (uint32_t)(p - data);
            ++lines;
            seq_end = p + ll;
        }
        seq_len += ll;
        p = e ? e + 1 : end;
    }
    if (seq_len > 0xffffffffull)
        throw std::runtime_error("sequence longer than 2^32 bases");
    r.seq_len = (uint32_t)seq_len;
    r.seq_span = (uint32_t)(seq_end - (data + r.seq_off));
    r.single_line = lines <= 1;
    if (marker == '@')
    {
        // '+' line, then quality lines until as many characters as bases were seen
        const char *e = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
        if (!e && !at_eof)
            return false;
        p = e ? e + 1 : end;
        size_t q = 0;
        while (q < seq_len)
        {
            if (p >= end)
            {
                if (!at_eof)
                    return false;
                break;
            }
            const char *qe = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
            if (!qe && !at_eof)
                return false;
            const char *le = qe ? qe : end;
            q += line_len(p, le);
            p = qe ? qe + 1 : end;
        }
        if (q != seq_len)
            throw std::runtime_error("FASTQ quality string length differs from the sequence length");
    }
    next_pos = (size_t)(p - data);
    return true;
}
} // namespace

bool RecordScanner::next(std::vector<char> &buf, std::vector<RecordRef> &recs, size_t target)
{
    recs.clear();
    if (!ok())
        return false;
    size_t have = carry_.size();
    if (buf.size() < std::max(target, have + (1u << 16)))
        buf.resize(std::max(target, have + (1u << 16)));
    if (have)
        memcpy(buf.data(), carry_.data(), have);
    carry_.clear();
    while (true)
    {
        have += fill(buf.data() + have, buf.size() - have);
        const char *data = buf.data();
        size_t pos = 0;
        while (true)
        {
            while (pos < have && (data[pos] == '\n' || data[pos] == '\r')) // blank lines between records
                ++pos;
            if (pos >= have)
                break;
            if (data[pos] != '>' && data[pos] != '@')
                throw std::runtime_error("sequence file: record does not start with '>' or '@'");
            RecordRef r{};
            size_t next_pos = pos;
            if (!scan_record(data, have, pos, eof_, r, next_pos))
                break;
            recs.push_back(r);
            pos = next_pos;
        }
        if (!recs.empty() || (eof_ && pos >= have))
        {
            carry_.assign(data + pos, data + have); // incomplete tail (empty at end of file)
            return !recs.empty();
        }
        if (eof_)
            throw std::runtime_error("sequence file: truncated record at end of file");
        // not even one complete record fits: grow and read on
        if (buf.size() > (size_t)0xfffffff0ull)
            throw std::runtime_error("sequence record larger than 4 GiB");
        buf.resize(std::min<size_t>(buf.size() * 2, 0xfffffff0ull));
    }
}

MappedFile::MappedFile(const std::string &path)
{
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0)
        return;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode))
    {
        ::close(fd);
        return;
    }
    size_ = (size_t)st.st_size;
    ok_ = true;
    if (size_)
    {
        void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED)
        {
            ok_ = false;
            size_ = 0;
        }
        else
        {
            data_ = static_cast<const char *>(p);
            madvise(p, size_, MADV_SEQUENTIAL);
            gzip_ = (size_ >= 2 && (unsigned char)data_[0] == 0x1f && (unsigned char)data_[1] == 0x8b) ||
                    (size_ >= 4 && data_[0] == 'B' && data_[1] == 'Z' && data_[2] == 'h' && data_[3] >= '1' && data_[3] <= '9');
        }
    }
    ::close(fd);
}

MappedFile::~MappedFile()
{
    if (data_)
        munmap(const_cast<char *>(data_), size_);
}

size_t guess_record_start(const char *data, size_t size, size_t from, char marker)
{
    size_t pos = from;
    if (pos > 0) // move to the next line start unless `from` is one
    {
        if (data[pos - 1] != '\n')
        {
            const char *nl = static_cast<const char *>(memchr(data + pos, '\n', size - pos));
            if (!nl)
                return size;
            pos = (size_t)(nl - data) + 1;
        }
    }
    const char *end = data + size;
    while (pos < size)
    {
        const char *l0 = data + pos;
        const char *e0 = static_cast<const char *>(memchr(l0, '\n', (size_t)(end - l0)));
        if (*l0 == marker)
        {
            if (marker == '>')
                return pos;
            // FASTQ, 4-line layout: header / bases / '+' / qualities of the same length
            if (e0)
            {
                const char *l1 = e0 + 1;
                const char *e1 = l1 < end ? static_cast<const char *>(memchr(l1, '\n', (size_t)(end - l1))) : nullptr;
                if (e1 && e1 + 1 < end && e1[1] == '+')
                {
                    const char *e2 = static_cast<const char *>(memchr(e1 + 1, '\n', (size_t)(end - e1 - 1)));
                    if (e2)
                    {
                        const char *l3 = e2 + 1;
                        const char *e3 = l3 < end ? static_cast<const char *>(memchr(l3, '\n', (size_t)(end - l3))) : nullptr;
                        const char *q_end = e3 ? e3 : end;
                        if (line_len(l1, e1) == line_len(l3, q_end))
                            return pos;
                    }
                }
            }
        }
        if (!e0)
            return size;
        pos = (size_t)(e0 - data) + 1;
    }
    return size;
}

size_t scan_segment(const char *data, size_t size, size_t begin, size_t end_hint, std::vector<RecordRef> &recs)
{
    recs.clear();
    const char *base = data + begin;
    const size_t n = size - begin;
    size_t pos = 0;
    while (true)
    {
        while (pos < n && (base[pos] == '\n' || base[pos] == '\r'))
            ++pos;
        if (pos >= n || begin + pos >= end_hint)
            break;
        if (base[pos] != '>' && base[pos] != '@')
            throw std::runtime_error("sequence file: record does not start with '>' or '@'");
        if (pos > 0xf0000000ull)
            throw std::runtime_error("sequence record larger than 4 GiB");
        RecordRef r{};
        size_t next_pos = pos;
        scan_record(base, n, pos, true, r, next_pos); // the whole file is mapped: never incomplete
        if (next_pos > 0xfffffff0ull)
            throw std::runtime_error("sequence record larger than 4 GiB");
        recs.push_back(r);
        pos = next_pos;
    }
    return begin + pos;
}

void scan_byte_range(const char *data, size_t size, size_t first_record, char marker, size_t lo, size_t hi, SegmentScan &out)
{
    try
    {
        out.begin = lo == 0 ? first_record : guess_record_start(data, size, lo, marker);
        out.end = out.begin < hi ? scan_segment(data, size, out.begin, hi, out.recs) : out.begin;
    }
    catch (std::exception const &e)
    {
        out.error = e.what();
    }
}

size_t accept_byte_range(const char *data, size_t size, size_t expected, size_t hi, SegmentScan &sg)
{
    if (expected >= hi) // the record before spans this whole range
    {
        sg.recs.clear();
        sg.begin = sg.end = expected;
        return expected;
    }
    if (!sg.error.empty() || sg.begin != expected)
    {
        sg.error.clear();
        sg.begin = expected;
        sg.end = scan_segment(data, size, expected, hi, sg.recs);
    }
    return sg.end;
}

void clean_record(const char *raw, const RecordRef &r, std::string &out)
{
    out.clear();
    out.reserve(r.seq_len);
    const char *p = raw + r.seq_off, *end = p + r.seq_span;
    for (; p < end; ++p)
    {
        const unsigned char c = (unsigned char)*p;
        if (c == ' ' || (c >= '\t' && c <= '\r') || (r.fasta && c >= '0' && c <= '9'))
            continue;
        out.push_back((char)c);
    }
}

void join_record(const char *raw, const RecordRef &r, std::string &out)
{
    out.clear();
    out.reserve(r.seq_len);
    const char *p = raw + r.seq_off, *end = p + r.seq_span;
    while (p < end)
    {
        const char *e = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
        // the span ends where the last line's bases end: its line terminator ('\r') is already outside
        out.append(p, e ? line_len(p, e) : (size_t)(end - p));
        p = e ? e + 1 : end;
    }
}
} // namespace txr
