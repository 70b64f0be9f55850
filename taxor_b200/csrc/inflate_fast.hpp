// inflate_fast.hpp -- DEFLATE (RFC 1951) / gzip (RFC 1952) decoder for the read ingest of `taxor search`.
//
// The reference reads .gz input through SeqAn3 -> zlib, one byte-oriented state machine (taxor_search.cpp:181-184).  Nanopore
// FASTQ is literal-dominated (bases cost ~2 bits, qualities 3-5), which is zlib's worst case: one table lookup, one bounds
// check and one state transition per output byte.  This decoder works on whole mapped files instead: a 64-bit bit buffer
// refilled once per iteration, an 11-bit primary table, up to three literals per refill, matches copied in words, and the
// per-symbol bounds checks replaced by a margin on both buffers (the tail runs through a careful copy of the same loop).
// Written from the RFCs; zlib stays the checker (tests/test_ingest.py compares the two on every block type) and the fallback
// for input that cannot be mapped (TAXOR_GZIP=zlib forces it).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace txr
{
// One raw DEFLATE stream held in memory, decoded piecewise into caller-provided output space.
class Inflater
{
public:
    static constexpr size_t kMargin = 320; // room the fast loop wants: 3 literals + one 258-byte match + copy overshoot

    void reset(const uint8_t *in, size_t n);
    // Decodes into [out, out_end).  A match may reach back to `hist` (at most 32768 bytes are ever used).  Returns the new end
    // of the output.  Comes back when the stream has ended (done()) or when the next symbol does not fit the room that is left
    // (always possible to continue with >= 258 bytes of room).  Throws std::runtime_error on damaged or truncated input.
    uint8_t *run(const uint8_t *hist, uint8_t *out, uint8_t *out_end);
    bool done() const { return state_ == State::done; }
    const uint8_t *input_pos() const; // after done(): the first byte behind the stream

    // ---- for decoding a stream in pieces, from several threads (gzip_parallel.cpp) ----
    // The same with 16-bit output elements: a piece that starts in the middle of a stream does not know the 32 KiB before it;
    // the caller presets them with marker values >= 256, matches copy the markers along, and they are replaced later.
    uint16_t *run16(const uint16_t *hist, uint16_t *out, uint16_t *out_end);
    void reset_at_bit(const uint8_t *in, size_t n, uint64_t bit); // a block header starts at this bit of `in`
    uint64_t bit_pos() const;                                     // bits of `in` consumed so far
    void stop_at_block_from(uint64_t bit) { stop_bit_ = bit; }    // run() returns at the first block boundary at or behind `bit`
    bool stopped() const { return stopped_; }                     // ... and says so here; bit_pos() is that boundary

private:
    enum class State
    {
        header,
        stored,
        huffman,
        done
    };
    static constexpr unsigned kLitBits = 11, kDistBits = 8;
    static constexpr size_t kLitCap = (1u << kLitBits) + 288 * 16, kDistCap = (1u << kDistBits) + 32 * 128;

    void refill_safe();
    uint32_t take(unsigned n); // n <= 32 bits, safe refill
    void check_not_past_end() const;
    void read_block_header();
    void read_dynamic_tables();
    void use_fixed_tables();
    template <class T> T *run_t(const T *hist, T *out, T *out_end);
    template <class T> T *fast_body(const T *hist, T *out, T *out_end, bool &block_ended);
    template <class T> T *fast_generic(const T *hist, T *out, T *out_end, bool &block_ended);
    template <class T> T *fast_bmi2(const T *hist, T *out, T *out_end, bool &block_ended);
    template <class T> T *careful_loop(const T *hist, T *out, T *out_end, bool &block_ended);

    const uint8_t *base_{nullptr};
    uint64_t stop_bit_{~uint64_t(0)};
    bool stopped_{false};
    const uint8_t *p_{nullptr}, *end_{nullptr};
    uint64_t buf_{0};
    unsigned cnt_{0};
    size_t overrun_{0}; // imaginary zero bytes appended behind the input (consuming one of their bits = truncated input)
    State state_{State::done};
    bool final_{false};
    size_t stored_left_{0};
    const uint32_t *lit_{nullptr}, *dist_{nullptr}; // tables of the current block
    std::vector<uint32_t> lit_dyn_, dist_dyn_, lit_fixed_, dist_fixed_;
};

// Whole-buffer form (BGZF blocks): exactly out_len bytes must come out of exactly one stream.  Returns false on any damage.
bool inflate_raw_exact(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

// A gzip file (one member or several, RFC 1952) held in memory, handed out as a byte stream; CRC-32 and ISIZE of every member
// are checked.  Bytes behind the last member that do not start another one are ignored, as zlib's gzread does.
class GzipStream
{
public:
    GzipStream(const uint8_t *data, size_t size);
    size_t read(uint8_t *dst, size_t cap); // 0 = end of data; throws std::runtime_error on damage
    bool eof() const { return phase_ == Phase::end && rd_ == wr_; }

private:
    enum class Phase
    {
        header,
        body,
        end
    };
    bool start_member();
    const uint8_t *in_, *end_;
    Inflater inf_;
    std::vector<uint8_t> win_;
    size_t rd_{0}, wr_{0}, floor_{0}; // handed out / decoded / first byte of the current member still in win_
    uint32_t crc_{0};
    uint64_t isize_{0};
    uint64_t members_{0};
    Phase phase_{Phase::header};
};

uint32_t crc32_fast(uint32_t crc, const uint8_t *data, size_t n); // gzip CRC-32, PCLMULQDQ where the CPU has it, else zlib's
} // namespace txr
