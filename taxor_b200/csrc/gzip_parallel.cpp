// gzip_parallel.cpp -- see gzip_parallel.hpp.
#include "gzip_parallel.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <zlib.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace txr
{
namespace
{
[[noreturn]] void bad(const char *what)
{
    throw std::runtime_error(std::string("read error (corrupt or truncated gzip data: ") + what + ")");
}
inline uint64_t load64(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
inline uint32_t le32(const uint8_t *p) { return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// a bit reader that never throws and never reads behind `size`: the header probe runs on millions of wrong positions
struct Probe
{
    const uint8_t *data;
    size_t size;
    uint64_t bit;
    bool ok{true};
    uint32_t take(unsigned n) // n <= 16
    {
        const size_t byte = (size_t)(bit >> 3);
        if (byte + 4 > size)
        {
            ok = false;
            return 0;
        }
        uint32_t v;
        memcpy(&v, data + byte, 4);
        v = (v >> (bit & 7)) & ((1u << n) - 1);
        bit += n;
        return v;
    }
};

// Kraft sum of a set of code lengths, in units of 2^-15
inline uint32_t kraft(const uint8_t *lens, unsigned n, unsigned &used)
{
    uint32_t sum = 0;
    used = 0;
    for (unsigned i = 0; i < n; ++i)
        if (lens[i])
        {
            sum += 1u << (15 - lens[i]);
            ++used;
        }
    return sum;
}

// The rest of a dynamic block header behind the 17 bits already looked at: code-length code, the two sets of code lengths.
// Stricter than a decoder has to be -- a position that fails here is merely not used as a starting point.
bool plausible_dynamic_header(const uint8_t *data, size_t size, uint64_t bit, unsigned hlit, unsigned hdist, unsigned hclen)
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    Probe pr{data, size, bit + 17};
    uint8_t pre[19] = {0};
    for (unsigned i = 0; i < hclen; ++i)
        pre[order[i]] = (uint8_t)pr.take(3);
    if (!pr.ok)
        return false;
    unsigned used;
    uint32_t sum = 0;
    for (unsigned i = 0; i < 19; ++i)
        if (pre[i])
            sum += 1u << (7 - pre[i]);
    if (sum != 128) // the code-length code must be complete
        return false;
    // canonical code of at most 7 bits -> 128-entry table (symbol << 4 | length)
    uint8_t table[128];
    {
        unsigned count[8] = {0}, next[8], code = 0;
        for (unsigned i = 0; i < 19; ++i)
            ++count[pre[i]];
        count[0] = 0;
        for (unsigned l = 1; l <= 7; ++l)
        {
            code = (code + count[l - 1]) << 1;
            next[l] = code;
        }
        for (unsigned s = 0; s < 19; ++s)
        {
            const unsigned l = pre[s];
            if (!l)
                continue;
            unsigned c = next[l]++, rev = 0;
            for (unsigned i = 0; i < l; ++i, c >>= 1)
                rev = rev << 1 | (c & 1);
            for (unsigned i = rev; i < 128; i += 1u << l)
                table[i] = (uint8_t)(s << 3 | l); // s <= 18 needs 5 bits, l <= 7 needs 3
        }
    }
    uint8_t lens[286 + 30 + 140];
    const unsigned total = hlit + hdist;
    unsigned i = 0;
    while (i < total)
    {
        const size_t byte = (size_t)(pr.bit >> 3);
        if (byte + 4 > size)
            return false;
        uint32_t v;
        memcpy(&v, data + byte, 4);
        v >>= (pr.bit & 7);
        const unsigned e = table[v & 127], l = e & 7, sym = e >> 3;
        pr.bit += l;
        if (sym < 16)
        {
            lens[i++] = (uint8_t)sym;
            continue;
        }
        unsigned rep;
        uint8_t fill = 0;
        if (sym == 16)
        {
            if (!i)
                return false;
            fill = lens[i - 1];
            rep = 3 + pr.take(2);
        }
        else if (sym == 17)
            rep = 3 + pr.take(3);
        else
            rep = 11 + pr.take(7);
        if (!pr.ok || i + rep > total)
            return false;
        memset(lens + i, fill, rep);
        i += rep;
    }
    if (!lens[256])
        return false;
    if (kraft(lens, hlit, used) != 32768) // literal/length code: complete
        return false;
    sum = kraft(lens + hlit, hdist, used);
    return sum == 32768 || used == 0 || (used == 1 && sum == 16384); // distance code: complete, absent, or the single 1-bit code
}

// 16-bit elements -> bytes; markers (>= 0x8000) index the 32 KiB in front of dst
void resolve_markers(const uint16_t *sym, size_t n, uint8_t *dst)
{
    const uint8_t *const window = dst - 32768;
    size_t i = 0;
#if defined(__x86_64__)
    const __m128i high = _mm_set1_epi16((short)0xff00), zero = _mm_setzero_si128();
    for (; i + 16 <= n; i += 16)
    {
        const __m128i a = _mm_loadu_si128((const __m128i *)(sym + i)), b = _mm_loadu_si128((const __m128i *)(sym + i + 8));
        if (_mm_movemask_epi8(_mm_cmpeq_epi16(_mm_and_si128(_mm_or_si128(a, b), high), zero)) == 0xffff)
            _mm_storeu_si128((__m128i *)(dst + i), _mm_packus_epi16(a, b));
        else
            for (size_t k = i; k < i + 16; ++k)
                dst[k] = sym[k] < 256 ? (uint8_t)sym[k] : window[sym[k] & 0x7fff];
    }
#endif
    for (; i < n; ++i)
        dst[i] = sym[i] < 256 ? (uint8_t)sym[i] : window[sym[i] & 0x7fff];
}
} // namespace

const uint8_t *gzip_member_body(const uint8_t *p, const uint8_t *end)
{
    if (end - p < 2 || p[0] != 0x1f || p[1] != 0x8b)
        return nullptr;
    if (end - p < 10)
        bad("the data ends inside a gzip header");
    if (p[2] != 8 || (p[3] & 0xe0))
        bad("unknown compression method or header flags");
    const unsigned flags = p[3];
    p += 10;
    if (flags & 4) // FEXTRA
    {
        if (end - p < 2)
            bad("the data ends inside a gzip header");
        const size_t xlen = p[0] | (size_t)p[1] << 8;
        p += 2;
        if ((size_t)(end - p) < xlen)
            bad("the data ends inside a gzip header");
        p += xlen;
    }
    for (unsigned bit : {8u, 16u}) // FNAME, FCOMMENT: zero-terminated
        if (flags & bit)
        {
            const void *z = memchr(p, 0, (size_t)(end - p));
            if (!z)
                bad("the data ends inside a gzip header");
            p = static_cast<const uint8_t *>(z) + 1;
        }
    if (flags & 2) // FHCRC
    {
        if (end - p < 2)
            bad("the data ends inside a gzip header");
        p += 2;
    }
    return p;
}

uint64_t find_dynamic_block(const uint8_t *data, size_t size, uint64_t from_bit, uint64_t to_bit)
{
    if (size < 16)
        return ~uint64_t(0);
    const uint64_t last = std::min<uint64_t>(to_bit, (uint64_t)(size - 12) * 8);
    for (uint64_t bit = from_bit; bit < last; ++bit)
    {
        const uint32_t v = (uint32_t)(load64(data + (bit >> 3)) >> (bit & 7));
        if ((v & 7) != 4) // BFINAL = 0, BTYPE = 2 (dynamic Huffman)
            continue;
        const unsigned hlit = ((v >> 3) & 31) + 257, hdist = ((v >> 8) & 31) + 1, hclen = ((v >> 13) & 15) + 4;
        if (hlit > 286 || hdist > 30)
            continue;
        if (plausible_dynamic_header(data, size, bit, hlit, hdist, hclen))
            return bit;
    }
    return ~uint64_t(0);
}

// ---------------------------------------------------------------------------------------------------------------------------
ParallelGzip::ParallelGzip(const uint8_t *data, size_t size, unsigned threads, size_t piece_bytes) : data_(data), size_(size)
{
    if (!piece_bytes)
    {
        piece_bytes = size_t(1) << 20;
        if (const char *e = getenv("TAXOR_GZIP_PIECE"))
            piece_bytes = (size_t)std::max(1024L, atol(e));
    }
    piece_bytes_ = piece_bytes;
    const uint8_t *body = gzip_member_body(data, data + size);
    if (!body)
        bad("not a gzip file");
    first_body_ = (size_t)(body - data);
    pos_bit_ = (uint64_t)first_body_ * 8;
    n_pieces_ = std::max<size_t>(1, (size - first_body_ + piece_bytes_ - 1) / piece_bytes_);
    pieces_.resize(n_pieces_);
    n_threads_ = std::max(1u, threads);
    window_pieces_ = 2 * (size_t)n_threads_ + 2;
    window_.assign(kWindow, 0);
    for (unsigned t = 0; t < n_threads_; ++t)
        threads_.emplace_back([this] { worker(); });
}

ParallelGzip::~ParallelGzip()
{
    {
        std::lock_guard<std::mutex> l(m_);
        quit_ = true;
    }
    cv_work_.notify_all();
    for (auto &t : threads_)
        t.join();
    for (auto &pc : pieces_)
        if (pc)
            free(pc->sym);
    for (auto &b : out_q_)
        if (b->piece)
            free(b->piece->sym);
    if (cur_ && cur_->piece)
        free(cur_->piece->sym);
    for (auto &b : pool_)
        free(b.p);
}

ParallelGzip::Buffer ParallelGzip::get_buffer(size_t cap)
{
    {
        std::lock_guard<std::mutex> l(m_);
        if (!pool_.empty())
        {
            // the largest one: pieces of one file inflate to similar sizes, so it rarely has to grow again
            auto it = std::max_element(pool_.begin(), pool_.end(), [](const Buffer &a, const Buffer &b) { return a.cap < b.cap; });
            const Buffer b = *it;
            pool_.erase(it);
            return b;
        }
    }
    const Buffer b{static_cast<uint16_t *>(malloc(cap * sizeof(uint16_t))), cap};
    if (!b.p)
        throw std::bad_alloc();
    return b;
}

void ParallelGzip::put_buffer(Piece &pc) // m_ held by the caller
{
    if (pc.sym)
        pool_.push_back(Buffer{pc.sym, pc.cap});
    pc.sym = nullptr;
    pc.cap = 0;
}

std::vector<uint8_t> ParallelGzip::get_bytes(size_t n)
{
    std::vector<uint8_t> v;
    {
        std::lock_guard<std::mutex> l(m_);
        if (!byte_pool_.empty())
        {
            auto it = std::max_element(byte_pool_.begin(), byte_pool_.end(),
                                       [](const std::vector<uint8_t> &a, const std::vector<uint8_t> &b) { return a.size() < b.size(); });
            v = std::move(*it);
            byte_pool_.erase(it);
        }
    }
    if (v.size() < n)
        v.resize(n);
    return v;
}

void ParallelGzip::put_bytes(std::vector<uint8_t> &&v)
{
    std::lock_guard<std::mutex> l(m_);
    byte_pool_.push_back(std::move(v));
}

void ParallelGzip::worker()
{
    for (;;)
    {
        Block *job = nullptr;
        Piece *pc = nullptr;
        size_t k = 0;
        {
            std::unique_lock<std::mutex> l(m_);
            cv_work_.wait(l, [&] {
                return quit_ || !resolve_jobs_.empty() || (next_piece_ < n_pieces_ && next_piece_ < consumed_ + window_pieces_);
            });
            if (quit_)
                return;
            if (!resolve_jobs_.empty()) // finishing what the reader waits for comes before decoding further ahead
            {
                job = resolve_jobs_.front();
                resolve_jobs_.pop_front();
            }
            else
            {
                k = next_piece_++;
                pieces_[k].reset(new Piece);
                pc = pieces_[k].get();
            }
        }
        if (job)
            resolve_block(*job);
        else
            decode_piece(k, *pc);
        {
            std::lock_guard<std::mutex> l(m_);
            if (job)
            {
                put_buffer(*job->piece);
                job->piece.reset();
                job->ready = true;
            }
            else
                pc->done = true;
        }
        cv_done_.notify_all();
    }
}

// Stage 1.  A piece: from the first plausible block start in its byte range (piece 0: from the start of the data) to the first
// block boundary at or behind the end of the range, or to the end of the member.
void ParallelGzip::decode_piece(size_t k, Piece &pc)
{
    const size_t c0 = first_body_ + k * piece_bytes_, c1 = std::min(size_, c0 + piece_bytes_);
    const uint64_t stop_from = k + 1 < n_pieces_ ? (uint64_t)c1 * 8 : ~uint64_t(0);
    try
    {
        uint64_t start = (uint64_t)first_body_ * 8;
        if (k)
        {
            start = find_dynamic_block(data_, size_, (uint64_t)c0 * 8, (uint64_t)c1 * 8);
            if (start == ~uint64_t(0))
                return;
        }
        pc.found = true;
        pc.start_bit = start;
        // elements: a piece that inflates beyond this (more than 16x at the default piece size) is left to the sequential path,
        // which needs no memory to speak of -- sequence files compress 2-8x
        constexpr size_t kMaxOut = size_t(16) << 20;
        {
            const Buffer b = get_buffer(kWindow + std::max<size_t>(4 * (c1 - c0), 1 << 16));
            pc.sym = b.p;
            pc.cap = b.cap;
        }
        size_t cap = pc.cap;
        for (size_t w = 0; w < kWindow; ++w)
            pc.sym[w] = (uint16_t)(0x8000 + w);
        size_t out = kWindow;
        Inflater inf;
        inf.reset_at_bit(data_, size_, start);
        inf.stop_at_block_from(stop_from);
        for (;;)
        {
            if (cap - out < 4 * Inflater::kMargin)
            {
                if (cap - kWindow >= kMaxOut)
                    bad("piece inflates too far");
                cap = kWindow + std::min(kMaxOut, (cap - kWindow) * 2);
                uint16_t *grown = static_cast<uint16_t *>(realloc(pc.sym, cap * sizeof(uint16_t)));
                if (!grown)
                    throw std::bad_alloc();
                pc.sym = grown;
                pc.cap = cap;
            }
            uint16_t *const base = pc.sym;
            // piece 0 starts a member: nothing before it may be referenced; the others may reach into the marker prefix
            uint16_t *const o = inf.run16(k ? base : base + kWindow, base + out, base + cap);
            out = (size_t)(o - base);
            if (inf.stopped())
            {
                pc.end_bit = inf.bit_pos();
                break;
            }
            if (inf.done())
            {
                const uint8_t *t = inf.input_pos();
                if (data_ + size_ - t < 8)
                    bad("the data ends before the gzip trailer");
                pc.member_end = true;
                pc.want_crc = le32(t);
                pc.want_isize = le32(t + 4);
                pc.end_bit = (uint64_t)(t + 8 - data_) * 8;
                break;
            }
        }
        pc.n_out = out - kWindow;
    }
    catch (std::exception const &e) // damage, a wrong starting point, or no memory: the reader decodes this stretch itself
    {
        pc.error = e.what();
    }
}

// Stage 3.  The history the piece did not know is in front of the block's bytes by now.
void ParallelGzip::resolve_block(Block &b)
{
    resolve_markers(b.piece->sym + kWindow, b.n, b.buf.data() + kWindow);
    b.crc = crc32_fast(0, b.buf.data() + kWindow, b.n);
}

void ParallelGzip::push_window(const uint8_t *bytes, size_t n)
{
    if (n >= kWindow)
        memcpy(window_.data(), bytes + n - kWindow, kWindow);
    else
    {
        memmove(window_.data(), window_.data() + n, kWindow - n);
        memcpy(window_.data() + kWindow - n, bytes, n);
    }
}

// Stage 2 for a piece that continues the chain: only its last 32 KiB are resolved here (they are the history of whatever
// follows); the rest is left to a worker.
void ParallelGzip::chain_take_piece(std::unique_ptr<Piece> pc)
{
    const size_t n = pc->n_out;
    std::unique_ptr<Block> b(new Block);
    b->buf = get_bytes(kWindow + n);
    memcpy(b->buf.data(), window_.data(), kWindow);
    b->n = n;
    b->member_end = pc->member_end;
    b->want_crc = pc->want_crc;
    b->want_isize = pc->want_isize;
    {
        const size_t tail = std::min(n, kWindow);
        std::vector<uint8_t> scratch(kWindow + tail);
        memcpy(scratch.data(), window_.data(), kWindow);
        // a marker names a byte of the window in front of the PIECE, also in the tail of a long piece
        const uint16_t *src = pc->sym + kWindow + (n - tail);
        for (size_t i = 0; i < tail; ++i)
            scratch[kWindow + i] = src[i] < 256 ? (uint8_t)src[i] : scratch[src[i] & 0x7fff];
        push_window(scratch.data() + kWindow, tail);
    }
    pos_bit_ = pc->end_bit;
    member_bytes_ += n;
    if (pc->member_end)
        in_member_ = false;
    stats_.parallel_bytes += n;
    ++stats_.pieces_used;
    b->piece = std::move(pc);
    {
        std::lock_guard<std::mutex> l(m_);
        out_q_.push_back(std::move(b));
        resolve_jobs_.push_back(out_q_.back().get());
    }
    cv_work_.notify_one();
}

// Stage 2 for a stretch no piece covers: decoded here, with the real history, up to the boundary the caller asked for (set when
// seq_ was started) or the end of the member, at most one block buffer per call.
void ParallelGzip::chain_sequential()
{
    constexpr size_t kChunk = size_t(4) << 20;
    std::unique_ptr<Block> b(new Block);
    b->buf = get_bytes(kWindow + kChunk);
    memcpy(b->buf.data(), window_.data(), kWindow);
    uint8_t *const from = b->buf.data() + kWindow;
    const size_t reach = (size_t)std::min<uint64_t>(kWindow, member_bytes_); // a member cannot refer to what was before it
    uint8_t *const to = seq_.run(from - reach, from, from + kChunk);
    const size_t n = (size_t)(to - from);
    b->n = n;
    b->crc = crc32_fast(0, from, n);
    b->ready = true;
    member_bytes_ += n;
    stats_.sequential_bytes += n;
    push_window(from, n);
    if (seq_.stopped())
    {
        pos_bit_ = seq_.bit_pos();
        seq_active_ = false;
    }
    else if (seq_.done())
    {
        const uint8_t *t = seq_.input_pos();
        if (data_ + size_ - t < 8)
            bad("the data ends before the gzip trailer");
        b->member_end = true;
        b->want_crc = le32(t);
        b->want_isize = le32(t + 4);
        pos_bit_ = (uint64_t)(t + 8 - data_) * 8;
        in_member_ = false;
        seq_active_ = false;
    }
    else if (!n)
        bad("decoder made no progress");
    std::lock_guard<std::mutex> l(m_);
    out_q_.push_back(std::move(b));
}

void ParallelGzip::advance_chain(bool may_wait)
{
    for (;;)
    {
        if (chain_finished_)
            return;
        {
            std::lock_guard<std::mutex> l(m_);
            if (out_q_.size() >= (size_t)n_threads_ + 2)
                return;
        }
        if (seq_active_)
        {
            chain_sequential();
            continue;
        }
        if (!in_member_)
        {
            const uint8_t *p = data_ + (size_t)(pos_bit_ >> 3);
            const uint8_t *body = p < data_ + size_ ? gzip_member_body(p, data_ + size_) : nullptr;
            if (!body) // the end, or bytes behind the last member that are not another member: ignored, as gzread does
            {
                chain_finished_ = true;
                return;
            }
            pos_bit_ = (uint64_t)(body - data_) * 8;
            in_member_ = true;
            member_bytes_ = 0;
        }
        // the piece whose byte range holds pos_bit_: usable only if it starts exactly here and has not been looked at before
        const size_t k = std::min(n_pieces_ - 1, ((size_t)(pos_bit_ >> 3) - first_body_) / piece_bytes_);
        std::unique_ptr<Piece> pc;
        if (k >= cur_piece_)
        {
            std::unique_lock<std::mutex> l(m_);
            // pieces before k cover nothing that is still needed: those not started yet never will be, the others are dropped
            // once their workers are through with them
            next_piece_ = std::max(next_piece_, k);
            consumed_ = std::max(consumed_, k);
            cv_work_.notify_all();
            auto decoded = [&] {
                for (size_t i = cur_piece_; i <= k; ++i)
                    if (pieces_[i] ? !pieces_[i]->done : i == k)
                        return false;
                return true;
            };
            if (!decoded())
            {
                if (!may_wait)
                    return;
                cv_done_.wait(l, decoded);
            }
            for (size_t i = cur_piece_; i < k; ++i)
                if (pieces_[i])
                {
                    put_buffer(*pieces_[i]);
                    pieces_[i].reset();
                }
            pc = std::move(pieces_[k]);
            stats_.pieces += k + 1 - cur_piece_;
            cur_piece_ = k + 1;
            consumed_ = cur_piece_;
            cv_work_.notify_all();
        }
        if (pc && pc->found && pc->error.empty() && pc->start_bit == pos_bit_)
        {
            chain_take_piece(std::move(pc));
            continue;
        }
        if (pc)
        {
            std::lock_guard<std::mutex> l(m_);
            put_buffer(*pc);
        }
        const uint64_t range_end = (uint64_t)std::min(size_, first_body_ + (k + 1) * piece_bytes_) * 8;
        seq_.reset_at_bit(data_, size_, pos_bit_);
        seq_.stop_at_block_from(k + 1 < n_pieces_ ? std::max(range_end, pos_bit_ + 1) : ~uint64_t(0));
        seq_active_ = true;
    }
}

size_t ParallelGzip::read(uint8_t *dst, size_t cap)
{
    size_t got = 0;
    while (got < cap)
    {
        if (cur_ && cur_rd_ < cur_->n)
        {
            const size_t n = std::min(cap - got, cur_->n - cur_rd_);
            memcpy(dst + got, cur_->buf.data() + kWindow + cur_rd_, n);
            cur_rd_ += n;
            got += n;
            continue;
        }
        if (cur_)
        {
            put_bytes(std::move(cur_->buf));
            cur_.reset();
        }
        if (finished_)
            break;
        advance_chain(false);
        {
            std::unique_lock<std::mutex> l(m_);
            if (out_q_.empty())
            {
                l.unlock();
                if (chain_finished_)
                {
                    finished_ = true;
                    break;
                }
                advance_chain(true);
                continue;
            }
            cv_done_.wait(l, [&] { return out_q_.front()->ready; });
            cur_ = std::move(out_q_.front());
            out_q_.pop_front();
        }
        cur_rd_ = 0;
        crc_ = (uint32_t)crc32_combine(crc_, cur_->crc, (z_off_t)cur_->n);
        isize_ += cur_->n;
        if (cur_->member_end)
        {
            if (cur_->want_crc != crc_ || cur_->want_isize != (uint32_t)isize_)
                bad("CRC or length mismatch");
            crc_ = 0;
            isize_ = 0;
        }
    }
    return got;
}
} // namespace txr
