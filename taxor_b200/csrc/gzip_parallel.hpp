// gzip_parallel.hpp -- one gzip stream inflated by several threads (reads arrive as .fastq.gz).
//
// DEFLATE has no index: a decoder that starts in the middle of a stream neither knows where a block begins nor what the
// 32 KiB before it were.  Both can be worked around (the two-pass idea of pugz / rapidgzip, written here from the format):
//   * block starts are FOUND: from a byte offset on, every bit position is tried as the header of a dynamic-Huffman block and
//     kept only if the whole header is consistent (complete code-length code, complete literal/length code with an end-of-block
//     symbol, plausible distance code);
//   * the unknown window is CARRIED SYMBOLICALLY: the piece decodes into 16-bit elements whose 32 KiB of pre-history are marker
//     values; matches copy markers like bytes, and once the piece before is finished the markers are replaced by its last
//     32 KiB.
// Nothing is trusted: a piece is only used if it starts at exactly the bit where the piece before it ended (a chain from the
// first byte of the file), anything else is decoded again sequentially, and every member's CRC-32 and length are checked over
// the final bytes.  Input with fixed or stored blocks at the piece borders, damaged input and tiny files simply take the
// sequential path more often.
#pragma once
#include "inflate_fast.hpp"

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace txr
{
// First byte of the DEFLATE data of the gzip member that starts at p (RFC 1952 header skipped); nullptr if [p, end) does not
// start with the gzip magic; throws std::runtime_error on a damaged or cut header.
const uint8_t *gzip_member_body(const uint8_t *p, const uint8_t *end);

// Bit position >= from_bit (and < to_bit) of the first header of a non-final dynamic-Huffman block that passes every consistency
// check, or ~0 if there is none.  A found position is a candidate, not a fact.
uint64_t find_dynamic_block(const uint8_t *data, size_t size, uint64_t from_bit, uint64_t to_bit);

class ParallelGzip
{
public:
    // piece_bytes: compressed bytes per piece (0: default 1 MiB, TAXOR_GZIP_PIECE overrides -- tests use small pieces)
    ParallelGzip(const uint8_t *data, size_t size, unsigned threads, size_t piece_bytes = 0);
    ~ParallelGzip();
    ParallelGzip(const ParallelGzip &) = delete;
    ParallelGzip &operator=(const ParallelGzip &) = delete;
    size_t read(uint8_t *dst, size_t cap); // 0 = end of data; throws std::runtime_error on damage

    struct Stats
    {
        uint64_t pieces{0}, pieces_used{0}, sequential_bytes{0}, parallel_bytes{0};
    };
    const Stats &stats() const { return stats_; }

private:
    static constexpr size_t kWindow = 32768;
    struct Piece // stage 1 (worker): a stretch of the stream decoded with a symbolic window
    {
        bool found{false}, done{false}, member_end{false};
        uint64_t start_bit{0}, end_bit{0};
        uint16_t *sym{nullptr}; // kWindow marker elements, then the output (a buffer of the pool below)
        size_t cap{0}, n_out{0};
        uint32_t want_crc{0}, want_isize{0}; // the member's trailer, if the piece ends with the member
        std::string error;
    };
    struct Block // stage 3 (worker or reader): final bytes, in stream order
    {
        std::vector<uint8_t> buf; // kWindow bytes of history, then n bytes
        size_t n{0};
        uint32_t crc{0};
        bool ready{false}, member_end{false};
        uint32_t want_crc{0}, want_isize{0};
        std::unique_ptr<Piece> piece; // still to be resolved into buf (null for a block the reader decoded itself)
    };
    // buffers are recycled: a fresh 10-30 MB allocation per piece costs more in page faults than the decoding saves
    struct Buffer
    {
        uint16_t *p;
        size_t cap;
    };
    Buffer get_buffer(size_t cap);
    void put_buffer(Piece &pc);
    std::vector<uint8_t> get_bytes(size_t n);
    void put_bytes(std::vector<uint8_t> &&v);

    void worker();
    void decode_piece(size_t k, Piece &pc);
    void resolve_block(Block &b);
    // stage 2 (reader, cheap): follows the chain of block boundaries from the first byte, turns the pieces that fit into blocks
    // to resolve and decodes the stretches no piece covers itself
    void advance_chain(bool may_wait);
    void chain_take_piece(std::unique_ptr<Piece> pc);
    void chain_sequential();
    void push_window(const uint8_t *bytes, size_t n);

    const uint8_t *data_;
    size_t size_;
    size_t piece_bytes_;
    size_t first_body_{0}; // byte offset of the first member's DEFLATE data
    size_t n_pieces_{0};
    unsigned n_threads_{1};

    std::mutex m_; // everything below up to the chain state
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::unique_ptr<Piece>> pieces_;
    size_t next_piece_{0}, consumed_{0}, window_pieces_{0};
    std::deque<Block *> resolve_jobs_;
    std::deque<std::unique_ptr<Block>> out_q_; // blocks in stream order, ready or not
    std::vector<Buffer> pool_;
    std::vector<std::vector<uint8_t>> byte_pool_;
    bool quit_{false};
    std::vector<std::thread> threads_;

    // chain state (reader thread only)
    uint64_t pos_bit_{0}; // the stream up to this bit is in out_q_ (a block boundary or a member start)
    size_t cur_piece_{0}; // next piece to look at
    bool in_member_{true}, chain_finished_{false};
    uint64_t member_bytes_{0}; // bytes of the current member in the chain so far (how far back a match may reach)
    Inflater seq_;
    bool seq_active_{false};
    std::vector<uint8_t> window_; // 2 * kWindow: the last kWindow bytes of the chain sit in the upper half
    // hand-out state (reader thread only)
    std::unique_ptr<Block> cur_;
    size_t cur_rd_{0};
    uint32_t crc_{0};
    uint64_t isize_{0};
    bool finished_{false};
    Stats stats_;
};
} // namespace txr
