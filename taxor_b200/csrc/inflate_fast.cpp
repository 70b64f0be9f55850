// inflate_fast.cpp -- see inflate_fast.hpp.  Written from RFC 1951 / RFC 1952; zlib is only linked for the CRC fallback.
#include "inflate_fast.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace txr
{
namespace
{
// table entry: [31:16] payload | kLit | kEnd | kSub | [12:8] code length (length / distance entries) or index bits of the
// subtable (kSub) | [7:0] bits to consume (code + extra bits; for kSub the primary index width)
// payload: literal byte / base length / base distance / first entry of the subtable; 0 = no code ends here (invalid)
// a literal entry of the primary literal/length table with kPair set carries TWO literals (payload = first | second << 8) and
// consumes the bits of both codes
constexpr uint32_t kLit = 0x8000, kEnd = 0x4000, kSub = 0x2000, kPair = 0x0100;
inline void store16(uint8_t *p, uint32_t v)
{
    const uint16_t w = (uint16_t)v;
    memcpy(p, &w, 2);
}

inline uint64_t load64(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
inline void store64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }
inline uint32_t low_bits(uint64_t v, unsigned n) { return (uint32_t)(v & ((uint64_t(1) << n) - 1)); }

[[noreturn]] void bad(const char *what)
{
    throw std::runtime_error(std::string("read error (corrupt or truncated gzip data: ") + what + ")");
}

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

uint32_t litlen_entry(unsigned s)
{
    if (s < 256)
        return kLit | s << 16;
    if (s == 256)
        return kEnd;
    if (s < 286)
        return (uint32_t)kLenBase[s - 257] << 16 | (uint32_t)kLenExtra[s - 257] << 8;
    return kEnd | 1u << 16; // 286, 287: have codes in the fixed tree, never valid in data
}
uint32_t dist_entry(unsigned s)
{
    if (s < 30)
        return (uint32_t)kDistBase[s] << 16 | (uint32_t)kDistExtra[s] << 8;
    return kEnd | 1u << 16; // 30, 31
}
uint32_t precode_entry(unsigned s) { return s << 16; }

inline unsigned reverse_bits(unsigned code, unsigned len)
{
    unsigned r = 0;
    for (unsigned i = 0; i < len; ++i, code >>= 1)
        r = r << 1 | (code & 1);
    return r;
}

// Canonical Huffman code (RFC 1951 3.2.2) -> lookup table indexed by the next `tb` stream bits (codes are packed most significant
// bit first into a least-significant-bit-first stream, hence the reversal).  Codes longer than tb go through one subtable per
// tb-bit prefix, all of the size the longest code needs.  Same acceptance rules as zlib: over-subscribed sets are damage, an
// incomplete set only passes as the single one-bit code of a literal/length or distance tree.
void build_table(const uint8_t *lens, unsigned n, unsigned tb, uint32_t *tab, size_t cap, uint32_t (*entry)(unsigned), bool may_be_single)
{
    unsigned count[16] = {0};
    for (unsigned i = 0; i < n; ++i)
        ++count[lens[i]];
    count[0] = 0;
    unsigned maxlen = 15;
    while (maxlen && !count[maxlen])
        --maxlen;
    memset(tab, 0, sizeof(uint32_t) << tb);
    if (!maxlen)
        return; // no code at all (a block without matches has no distance tree): every lookup is invalid
    int left = 1;
    for (unsigned len = 1; len <= 15; ++len)
    {
        left <<= 1;
        left -= (int)count[len];
        if (left < 0)
            bad("over-subscribed Huffman code");
    }
    if (left > 0 && !(may_be_single && maxlen == 1))
        bad("incomplete Huffman code");
    unsigned next[16];
    unsigned code = 0;
    for (unsigned len = 1; len <= 15; ++len)
    {
        code = (code + count[len - 1]) << 1;
        next[len] = code;
    }
    size_t free_at = size_t(1) << tb;
    const unsigned sub_bits = maxlen > tb ? maxlen - tb : 0;
    for (unsigned s = 0; s < n; ++s)
    {
        const unsigned l = lens[s];
        if (!l)
            continue;
        const unsigned rev = reverse_bits(next[l]++, l);
        // low byte: bits to consume (code + extra bits); bits 8..12 of a length / distance entry: the code length alone
        const uint32_t e0 = entry(s), extra = (e0 >> 8) & 31, base = e0 & ~0x1f00u;
        const bool plain = (e0 & (kLit | kEnd)) != 0;
        const uint32_t e = plain ? base | l : base | l << 8 | (l + extra);
        if (l <= tb)
        {
            for (unsigned i = rev; i < (1u << tb); i += 1u << l)
                tab[i] = e;
            continue;
        }
        uint32_t &pe = tab[rev & ((1u << tb) - 1)];
        if (!(pe & kSub))
        {
            if (free_at + (size_t(1) << sub_bits) > cap)
                bad("Huffman table overflow");
            memset(tab + free_at, 0, sizeof(uint32_t) << sub_bits);
            pe = kSub | (uint32_t)free_at << 16 | sub_bits << 8 | tb;
            free_at += size_t(1) << sub_bits;
        }
        const uint32_t start = pe >> 16;
        const unsigned sl = l - tb;
        const uint32_t es = plain ? base | sl : base | sl << 8 | (sl + extra);
        for (unsigned i = rev >> tb; i < (1u << sub_bits); i += 1u << sl)
            tab[start + i] = es;
    }
}

// Literal-dominated data (bases, qualities) decodes one table lookup per byte, and every lookup waits for the shift of the one
// before.  Where the index bits behind a short literal code hold another complete literal code, the entry is rewritten to
// deliver both at once.  Descending order: tab[idx >> l1] is below idx (or idx itself at 0) and therefore still single.
void pair_literals(uint32_t *tab, unsigned tb)
{
    for (int idx = (1 << tb) - 1; idx >= 0; --idx)
    {
        const uint32_t e = tab[idx];
        const unsigned l1 = e & 0xff;
        if (!(e & kLit) || l1 >= tb)
            continue;
        const uint32_t e2 = tab[idx >> l1];
        const unsigned l2 = e2 & 0xff;
        if (!(e2 & kLit) || (e2 & kPair) || l1 + l2 > tb)
            continue;
        tab[idx] = kLit | kPair | ((e >> 16) | (e2 >> 16) << 8) << 16 | (l1 + l2);
    }
}
} // namespace

void Inflater::reset(const uint8_t *in, size_t n)
{
    base_ = in;
    stop_bit_ = ~uint64_t(0);
    stopped_ = false;
    p_ = in;
    end_ = in + n;
    buf_ = 0;
    cnt_ = 0;
    overrun_ = 0;
    state_ = State::header;
    final_ = false;
    stored_left_ = 0;
}

inline void Inflater::refill_safe()
{
    if (end_ - p_ >= 8)
    {
        buf_ |= load64(p_) << cnt_;
        p_ += (63 - cnt_) >> 3;
        cnt_ |= 56;
        return;
    }
    while (cnt_ <= 56)
    {
        if (p_ < end_)
            buf_ |= (uint64_t)*p_++ << cnt_;
        else
            ++overrun_; // an imaginary zero byte: decoding may look at it, consuming one of its bits is an error
        cnt_ += 8;
    }
}

inline uint32_t Inflater::take(unsigned n)
{
    if (cnt_ < n)
        refill_safe();
    const uint32_t v = low_bits(buf_, n);
    buf_ >>= n;
    cnt_ -= n;
    return v;
}

inline void Inflater::check_not_past_end() const
{
    if (overrun_ * 8 > cnt_)
        bad("the data ends inside a block");
}

const uint8_t *Inflater::input_pos() const
{
    const size_t in_buffer = cnt_ >> 3; // whole bytes not consumed yet, the last overrun_ of them imaginary
    return p_ - (in_buffer > overrun_ ? in_buffer - overrun_ : 0);
}

void Inflater::use_fixed_tables()
{
    if (lit_fixed_.empty())
    {
        uint8_t lens[288 + 32];
        for (unsigned s = 0; s < 288; ++s)
            lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        for (unsigned s = 0; s < 32; ++s)
            lens[288 + s] = 5;
        lit_fixed_.assign(kLitCap, 0);
        dist_fixed_.assign(kDistCap, 0);
        build_table(lens, 288, kLitBits, lit_fixed_.data(), kLitCap, litlen_entry, false);
        build_table(lens + 288, 32, kDistBits, dist_fixed_.data(), kDistCap, dist_entry, false);
        pair_literals(lit_fixed_.data(), kLitBits);
    }
    lit_ = lit_fixed_.data();
    dist_ = dist_fixed_.data();
}

void Inflater::read_dynamic_tables()
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const unsigned hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
    if (hlit > 286 || hdist > 30)
        bad("too many length or distance symbols");
    uint8_t pre_lens[19] = {0};
    for (unsigned i = 0; i < hclen; ++i)
        pre_lens[order[i]] = (uint8_t)take(3);
    uint32_t pre[128];
    build_table(pre_lens, 19, 7, pre, 128, precode_entry, false);
    uint8_t lens[286 + 30];
    const unsigned total = hlit + hdist;
    unsigned i = 0;
    while (i < total)
    {
        refill_safe();
        const uint32_t e = pre[buf_ & 127];
        if (!e)
            bad("invalid code-length code");
        buf_ >>= (e & 0xff);
        cnt_ -= (e & 0xff);
        const unsigned sym = e >> 16;
        if (sym < 16)
        {
            lens[i++] = (uint8_t)sym;
            continue;
        }
        unsigned rep;
        uint8_t v = 0;
        if (sym == 16)
        {
            if (!i)
                bad("code-length repeat with nothing before it");
            v = lens[i - 1];
            rep = 3 + take(2);
        }
        else if (sym == 17)
            rep = 3 + take(3);
        else
            rep = 11 + take(7);
        if (i + rep > total)
            bad("code-length repeat beyond the table");
        memset(lens + i, v, rep);
        i += rep;
    }
    check_not_past_end();
    if (!lens[256])
        bad("no end-of-block code");
    if (lit_dyn_.empty())
    {
        lit_dyn_.assign(kLitCap, 0);
        dist_dyn_.assign(kDistCap, 0);
    }
    build_table(lens, hlit, kLitBits, lit_dyn_.data(), kLitCap, litlen_entry, true);
    build_table(lens + hlit, hdist, kDistBits, dist_dyn_.data(), kDistCap, dist_entry, true);
    pair_literals(lit_dyn_.data(), kLitBits);
    lit_ = lit_dyn_.data();
    dist_ = dist_dyn_.data();
}

void Inflater::read_block_header()
{
    final_ = take(1) != 0;
    const unsigned type = take(2);
    check_not_past_end();
    if (type == 0)
    {
        buf_ >>= cnt_ & 7; // to the byte boundary
        cnt_ -= cnt_ & 7;
        const uint32_t len = take(16), nlen = take(16);
        check_not_past_end();
        if ((len ^ nlen) != 0xffff)
            bad("stored block length check");
        // hand the whole bytes still in the bit buffer back to the byte pointer
        const size_t in_buffer = cnt_ >> 3;
        p_ -= in_buffer - overrun_; // check_not_past_end: in_buffer >= overrun_
        buf_ = 0;
        cnt_ = 0;
        overrun_ = 0;
        stored_left_ = len;
        state_ = State::stored;
    }
    else if (type == 1)
    {
        use_fixed_tables();
        state_ = State::huffman;
    }
    else if (type == 2)
    {
        read_dynamic_tables();
        state_ = State::huffman;
    }
    else
        bad("reserved block type");
}

// ---- the symbol loops of a Huffman block ----
#define TXR_CONSUME(n)           \
    do                           \
    {                            \
        const unsigned n_ = (n); \
        buf >>= n_;              \
        cnt -= n_;               \
    } while (0)
// base + extra bits of a length / distance entry: the low byte is code + extra bits (consumed in one shift), bits 8..12 the code
// length alone, so the value comes from the bits as they were BEFORE the shift and stays off the buf -> index -> entry chain
#define TXR_VALUE(e, saved) ((size_t)((e) >> 16) + (low_bits((saved), (e)&0xff) >> (((e) >> 8) & 31)))

// Fast: the caller guarantees >= 32 input bytes and >= kMargin bytes of room at entry; the loop re-checks both before every
// refill and leaves when one no longer holds.  No per-symbol bounds checks inside.  The literal/length entry of the NEXT
// symbol is loaded before a match is copied, so that its latency hides behind the copy.
template <class T> __attribute__((always_inline)) inline T *Inflater::fast_body(const T *hist, T *out, T *out_end, bool &block_ended)
{
    constexpr uint32_t lmask = (1u << kLitBits) - 1, dmask = (1u << kDistBits) - 1;
    const uint32_t *const lit = lit_, *const dist = dist_;
    const uint8_t *p = p_, *const end = end_;
    uint64_t buf = buf_;
    unsigned cnt = cnt_;
    block_ended = false;
#define TXR_REFILL()                \
    do                              \
    {                               \
        buf |= load64(p) << cnt;    \
        p += (63 - cnt) >> 3;       \
        cnt |= 56;                  \
    } while (0)
#define TXR_LITERAL(e)                            \
    do                                            \
    {                                             \
        TXR_CONSUME((e)&0xff);                    \
        if (sizeof(T) == 1)                       \
            store16((uint8_t *)out, (e) >> 16);   \
        else                                      \
        {                                         \
            out[0] = (T)(((e) >> 16) & 0xff);     \
            out[1] = (T)((e) >> 24);              \
        }                                         \
        out += 1 + (((e) >> 8) & 1);              \
    } while (0)
#define TXR_ROOM() (end - p >= 32 && out_end - out >= (ptrdiff_t)kMargin)
    TXR_REFILL();
    uint32_t e = lit[buf & lmask];
    for (;;)
    {
        if (e & kLit)
        {
            // up to three literal lookups (one or two bytes each) per refill: 3 x 11 bits, or 2 x 11 + a 15-bit length code + 5
            // extra bits, fit the >= 56 bits.  Two bytes are always stored; the second one only counts for a pair.
            TXR_LITERAL(e);
            e = lit[buf & lmask];
            if (e & kLit)
            {
                TXR_LITERAL(e);
                e = lit[buf & lmask];
                if (e & kLit)
                {
                    TXR_LITERAL(e);
                    if (!TXR_ROOM())
                        break;
                    TXR_REFILL();
                    e = lit[buf & lmask];
                    continue;
                }
            }
        }
        if (e & kSub)
        {
            TXR_CONSUME(e & 0xff);
            e = lit[(e >> 16) + low_bits(buf, (e >> 8) & 31)];
            if (e & kLit)
            {
                TXR_LITERAL(e);
                if (!TXR_ROOM())
                    break;
                TXR_REFILL();
                e = lit[buf & lmask];
                continue;
            }
        }
        if (!e)
            bad("invalid literal/length code");
        if (e & kEnd)
        {
            if (e >> 16)
                bad("invalid literal/length symbol");
            TXR_CONSUME(e & 0xff);
            block_ended = true;
            break;
        }
        uint64_t saved = buf;
        TXR_CONSUME(e & 0xff);
        const size_t length = TXR_VALUE(e, saved);
        TXR_REFILL();
        e = dist[buf & dmask];
        if (e & kSub)
        {
            TXR_CONSUME(e & 0xff);
            e = dist[(e >> 16) + low_bits(buf, (e >> 8) & 31)];
        }
        if (!e || (e & kEnd))
            bad("invalid distance code");
        saved = buf;
        TXR_CONSUME(e & 0xff);
        const size_t distance = TXR_VALUE(e, saved);
        if (distance > (size_t)(out - hist))
            bad("match distance reaches before the start of the data");
        T *const stop = out + length;
        const bool more = end - p >= 32 && out_end - stop >= (ptrdiff_t)kMargin;
        if (more)
        {
            TXR_REFILL();
            e = lit[buf & lmask];
        }
        // the copy, in bytes: chunks of 16 or 8 where source and destination are that far apart, else element by element
        {
            uint8_t *o = reinterpret_cast<uint8_t *>(out), *const o_stop = reinterpret_cast<uint8_t *>(stop);
            const size_t gap = distance * sizeof(T);
            const uint8_t *src = o - gap;
            if (gap >= 16)
            {
                do
                {
                    store64(o, load64(src));
                    store64(o + 8, load64(src + 8));
                    o += 16;
                    src += 16;
                } while (o < o_stop);
            }
            else if (gap >= 8)
            {
                do
                {
                    store64(o, load64(src));
                    o += 8;
                    src += 8;
                } while (o < o_stop);
            }
            else if (sizeof(T) == 1 && distance == 1)
            {
                const uint64_t v = 0x0101010101010101ull * *src;
                do
                {
                    store64(o, v);
                    o += 8;
                } while (o < o_stop);
            }
            else
            {
                const T *s = out - distance;
                do
                    *out++ = *s++;
                while (out < stop);
            }
        }
        out = stop;
        if (!more)
            break;
    }
#undef TXR_REFILL
#undef TXR_LITERAL
#undef TXR_ROOM
    p_ = p;
    buf_ = buf;
    cnt_ = cnt;
    return out;
}

template <class T> T *Inflater::fast_generic(const T *hist, T *out, T *out_end, bool &block_ended)
{
    return fast_body<T>(hist, out, out_end, block_ended);
}
#if defined(__x86_64__)
// the same loop compiled for BMI2: shrx / shlx / bzhi instead of shifts through %cl and mask arithmetic
template <class T> __attribute__((target("bmi2"))) T *Inflater::fast_bmi2(const T *hist, T *out, T *out_end, bool &block_ended)
{
    return fast_body<T>(hist, out, out_end, block_ended);
}
#endif

// Careful: any amount of input and room; every symbol is checked against the room that is left and undone if it does not fit;
// bits of the imaginary bytes behind the input may be looked at but not consumed.
template <class T> T *Inflater::careful_loop(const T *hist, T *out, T *out_end, bool &block_ended)
{
    constexpr uint32_t lmask = (1u << kLitBits) - 1, dmask = (1u << kDistBits) - 1;
    const uint32_t *const lit = lit_, *const dist = dist_;
    block_ended = false;
    for (;;)
    {
        const uint8_t *const save_p = p_;
        const uint64_t save_buf = buf_;
        const unsigned save_cnt = cnt_;
        const size_t save_overrun = overrun_;
        auto undo = [&] {
            p_ = save_p;
            buf_ = save_buf;
            cnt_ = save_cnt;
            overrun_ = save_overrun;
        };
        refill_safe();
        uint64_t &buf = buf_;
        unsigned &cnt = cnt_;
        uint32_t e = lit[buf & lmask];
        if (e & kSub)
        {
            TXR_CONSUME(e & 0xff);
            e = lit[(e >> 16) + low_bits(buf, (e >> 8) & 31)];
        }
        if (!e)
            bad("invalid literal/length code");
        if (e & kLit)
        {
            const size_t n_lit = 1 + ((e >> 8) & 1);
            if ((size_t)(out_end - out) < n_lit)
            {
                undo();
                return out;
            }
            TXR_CONSUME(e & 0xff);
            check_not_past_end();
            *out++ = (T)((e >> 16) & 0xff);
            if (n_lit == 2)
                *out++ = (T)(e >> 24);
            continue;
        }
        if (e & kEnd)
        {
            if (e >> 16)
                bad("invalid literal/length symbol");
            TXR_CONSUME(e & 0xff);
            check_not_past_end();
            block_ended = true;
            return out;
        }
        uint64_t saved = buf;
        TXR_CONSUME(e & 0xff);
        const size_t length = TXR_VALUE(e, saved);
        refill_safe();
        e = dist[buf & dmask];
        if (e & kSub)
        {
            TXR_CONSUME(e & 0xff);
            e = dist[(e >> 16) + low_bits(buf, (e >> 8) & 31)];
        }
        if (!e || (e & kEnd))
            bad("invalid distance code");
        saved = buf;
        TXR_CONSUME(e & 0xff);
        const size_t distance = TXR_VALUE(e, saved);
        check_not_past_end();
        if (distance > (size_t)(out - hist))
            bad("match distance reaches before the start of the data");
        if (length > (size_t)(out_end - out))
        {
            undo();
            return out;
        }
        const T *src = out - distance;
        for (size_t i = 0; i < length; ++i)
            out[i] = src[i];
        out += length;
    }
}
#undef TXR_CONSUME
#undef TXR_VALUE

uint8_t *Inflater::run(const uint8_t *hist, uint8_t *out, uint8_t *out_end) { return run_t<uint8_t>(hist, out, out_end); }
uint16_t *Inflater::run16(const uint16_t *hist, uint16_t *out, uint16_t *out_end) { return run_t<uint16_t>(hist, out, out_end); }

uint64_t Inflater::bit_pos() const { return (uint64_t)(p_ - base_) * 8 + 8 * overrun_ - cnt_; }

void Inflater::reset_at_bit(const uint8_t *in, size_t n, uint64_t bit)
{
    reset(in, n);
    p_ = in + (bit >> 3);
    if (bit & 7)
        (void)take((unsigned)(bit & 7));
}

template <class T> T *Inflater::run_t(const T *hist, T *out, T *out_end)
{
    stopped_ = false;
    for (;;)
    {
        switch (state_)
        {
        case State::header:
            if (bit_pos() >= stop_bit_) // a block boundary at or behind the position the caller asked to stop at
            {
                stopped_ = true;
                return out;
            }
            read_block_header();
            break;
        case State::stored:
        {
            const size_t n = std::min({stored_left_, (size_t)(out_end - out), (size_t)(end_ - p_)});
            if (sizeof(T) == 1)
                memcpy(out, p_, n);
            else
                for (size_t i = 0; i < n; ++i)
                    out[i] = (T)p_[i];
            out += n;
            p_ += n;
            stored_left_ -= n;
            if (stored_left_)
            {
                if (p_ == end_)
                    bad("the data ends inside a stored block");
                return out; // no room
            }
            state_ = final_ ? State::done : State::header;
            break;
        }
        case State::huffman:
        {
            bool ended = false;
            if (end_ - p_ >= 32 && out_end - out >= (ptrdiff_t)kMargin)
            {
#if defined(__x86_64__)
                // TAXOR_INFLATE_ISA=generic keeps the baseline build of the loop (tests run both)
                static const bool bmi2 = __builtin_cpu_supports("bmi2") && !(getenv("TAXOR_INFLATE_ISA") && !strcmp(getenv("TAXOR_INFLATE_ISA"), "generic"));
                out = bmi2 ? fast_bmi2<T>(hist, out, out_end, ended) : fast_generic<T>(hist, out, out_end, ended);
#else
                out = fast_generic<T>(hist, out, out_end, ended);
#endif
            }
            if (!ended)
            {
                out = careful_loop<T>(hist, out, out_end, ended);
                if (!ended)
                    return out; // the next symbol does not fit
            }
            state_ = final_ ? State::done : State::header;
            break;
        }
        case State::done:
            return out;
        }
    }
}

bool inflate_raw_exact(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    thread_local Inflater inf;
    try
    {
        inf.reset(in, in_len);
        uint8_t *const o = inf.run(out, out, out + out_len); // comes back early only if a literal or match does not fit
        return inf.done() && o == out + out_len;
    }
    catch (std::runtime_error const &)
    {
        return false;
    }
}

// ---- gzip members ----
GzipStream::GzipStream(const uint8_t *data, size_t size) : in_(data), end_(data + size), win_((1u << 20) + 32768 + 2 * Inflater::kMargin) {}

bool GzipStream::start_member()
{
    if (in_ >= end_)
        return false;
    if (end_ - in_ < 2 || in_[0] != 0x1f || in_[1] != 0x8b)
    {
        if (!members_)
            bad("not a gzip file");
        return false; // bytes behind the last member that are not another member: ignored, as gzread does
    }
    if (end_ - in_ < 10)
        bad("the data ends inside a gzip header");
    if (in_[2] != 8 || (in_[3] & 0xe0))
        bad("unknown compression method or header flags");
    const unsigned flags = in_[3];
    const uint8_t *p = in_ + 10;
    if (flags & 4) // FEXTRA
    {
        if (end_ - p < 2)
            bad("the data ends inside a gzip header");
        const size_t xlen = p[0] | (size_t)p[1] << 8;
        p += 2;
        if ((size_t)(end_ - p) < xlen)
            bad("the data ends inside a gzip header");
        p += xlen;
    }
    for (unsigned bit : {8u, 16u}) // FNAME, FCOMMENT: zero-terminated
        if (flags & bit)
        {
            const void *z = memchr(p, 0, (size_t)(end_ - p));
            if (!z)
                bad("the data ends inside a gzip header");
            p = static_cast<const uint8_t *>(z) + 1;
        }
    if (flags & 2) // FHCRC
    {
        if (end_ - p < 2)
            bad("the data ends inside a gzip header");
        p += 2;
    }
    inf_.reset(p, (size_t)(end_ - p));
    crc_ = 0;
    isize_ = 0;
    ++members_;
    return true;
}

size_t GzipStream::read(uint8_t *dst, size_t cap)
{
    size_t got = 0;
    for (;;)
    {
        if (rd_ < wr_)
        {
            const size_t n = std::min(cap - got, wr_ - rd_);
            memcpy(dst + got, win_.data() + rd_, n);
            rd_ += n;
            got += n;
        }
        if (got == cap || phase_ == Phase::end)
            return got;
        if (phase_ == Phase::header)
        {
            if (!start_member())
            {
                phase_ = Phase::end;
                continue;
            }
            phase_ = Phase::body;
            floor_ = wr_; // a member cannot refer to the data of the one before it
        }
        // everything decoded so far has been handed out: keep the last 32 KiB of this member as history, decode the next piece
        if (win_.size() - wr_ < (1u << 19))
        {
            const size_t keep = std::min<size_t>(wr_ - floor_, 32768);
            memmove(win_.data(), win_.data() + wr_ - keep, keep);
            floor_ = 0;
            rd_ = wr_ = keep;
        }
        uint8_t *const from = win_.data() + wr_;
        uint8_t *const to = inf_.run(win_.data() + floor_, from, win_.data() + win_.size());
        const size_t n = (size_t)(to - from);
        crc_ = crc32_fast(crc_, from, n);
        isize_ += n;
        wr_ += n;
        if (inf_.done())
        {
            const uint8_t *t = inf_.input_pos();
            if (end_ - t < 8)
                bad("the data ends before the gzip trailer");
            const uint32_t crc = t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
            const uint32_t isz = t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
            if (crc != crc_ || isz != (uint32_t)isize_)
                bad("CRC or length mismatch");
            in_ = t + 8;
            phase_ = Phase::header;
        }
        else if (!n)
            bad("decoder made no progress");
    }
}

// ---- CRC-32 (the gzip polynomial, reflected) ----
#if defined(__x86_64__)
namespace
{
// Folding with carry-less multiplies (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ", Intel 2009):
// four 128-bit lanes are folded 64 bytes ahead per step, reduced to one lane, then to 64 and 32 bits (Barrett).  The constants
// are x^(n) mod P for the distances involved, bit-reflected; crc32_fast checks this routine against zlib's once, at start-up.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_pclmul(uint32_t crc, const uint8_t *buf, size_t len) // len >= 64, len % 16 == 0
{
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
    const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
    __m128i x1 = _mm_loadu_si128((const __m128i *)(buf + 0)), x2 = _mm_loadu_si128((const __m128i *)(buf + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i *)(buf + 32)), x4 = _mm_loadu_si128((const __m128i *)(buf + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    buf += 64;
    len -= 64;
    while (len >= 64)
    {
        const __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i *)(buf + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i *)(buf + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i *)(buf + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i *)(buf + 48)));
        buf += 64;
        len -= 64;
    }
#define TXR_FOLD(x, next) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128((x), k3k4, 0x11), _mm_clmulepi64_si128((x), k3k4, 0x00)), (next))
    x1 = TXR_FOLD(x1, x2);
    x1 = TXR_FOLD(x1, x3);
    x1 = TXR_FOLD(x1, x4);
    while (len >= 16)
    {
        x1 = TXR_FOLD(x1, _mm_loadu_si128((const __m128i *)buf));
        buf += 16;
        len -= 16;
    }
#undef TXR_FOLD
    // 128 -> 64 bits
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, mask32);
    x1 = _mm_clmulepi64_si128(x1, k5, 0x00);
    x1 = _mm_xor_si128(x1, t);
    // Barrett reduction to 32 bits
    t = _mm_and_si128(x1, mask32);
    t = _mm_clmulepi64_si128(t, poly, 0x10);
    t = _mm_and_si128(t, mask32);
    t = _mm_clmulepi64_si128(t, poly, 0x00);
    x1 = _mm_xor_si128(x1, t);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

bool pclmul_usable()
{
    static const bool ok = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1"))
            return false;
        uint8_t probe[208];
        for (size_t i = 0; i < sizeof probe; ++i)
            probe[i] = (uint8_t)(i * 151 + 7);
        for (size_t n : {size_t(64), size_t(80), size_t(128), size_t(208)})
            if (~crc32_pclmul(~0x12345678u, probe, n) != (uint32_t)::crc32(0x12345678u, probe, (unsigned)n))
                return false;
        return true;
    }();
    return ok;
}
} // namespace
#endif

uint32_t crc32_fast(uint32_t crc, const uint8_t *data, size_t n)
{
#if defined(__x86_64__)
    if (n >= 64 && pclmul_usable())
    {
        const size_t body = n & ~size_t(15);
        crc = ~crc32_pclmul(~crc, data, body);
        data += body;
        n -= body;
    }
#endif
    while (n)
    {
        const size_t step = std::min<size_t>(n, 1u << 30);
        crc = (uint32_t)::crc32(crc, data, (unsigned)step);
        data += step;
        n -= step;
    }
    return crc;
}
} // namespace txr
