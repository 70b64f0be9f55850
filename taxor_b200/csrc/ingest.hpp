// ingest.hpp -- parallel FASTA / FASTQ (plain or gzip) ingest for the `taxor search` driver.
//
// Replaces seqan3::sequence_file_input<dna4_traits, fields<id, seq>> + views::chunk(1024) on the main thread
// (src/main/taxor_search.cpp:181-184, 315-321), which parses serially while the workers idle.  Here ONE thread
// streams the file (inflating if it is gzip) and only finds record boundaries (memchr per line, no copying);
// the byte work -- IUPAC -> dna4 collapse and 2-bit packing straight into the pinned batch buffer, id strings --
// is done by a pool of threads on disjoint record ranges.  Record semantics (those of seqan3's FASTA/FASTQ formats as
// the search uses them): id = header line without '>' / '@', sequence = all sequence lines joined, one trailing
// '\r' stripped per line, blank lines between records skipped, FASTQ quality length checked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <zlib.h>

namespace txr
{
struct RecordRef // one record inside a raw buffer
{
    uint32_t id_off, id_len;   // header line without the marker and without "\r"
    uint32_t seq_off, seq_span; // first sequence byte .. end of the last sequence line (may contain line breaks)
    uint32_t seq_len;          // bases (line breaks excluded)
    uint32_t single_line;      // 1: the seq_len bases are contiguous at seq_off
    uint32_t fasta;            // 1: '>' record (SeqAn3 drops digits inside FASTA sequence lines, not inside FASTQ ones)
};

// Streams a sequence file as raw buffers that each start at a record boundary.
class RecordScanner
{
public:
    explicit RecordScanner(const std::string &path);
    ~RecordScanner();
    RecordScanner(const RecordScanner &) = delete;
    RecordScanner &operator=(const RecordScanner &) = delete;
    bool ok() const { return fd_ >= 0 || gz_ != nullptr || bgzf_data_ != nullptr || gzs_ != nullptr; }
    const std::string &open_error() const { return open_error_; } // why ok() is false, when there is more to say than "cannot open"
    // Fills `buf` (resized as needed; `target` bytes unless one record needs more) and appends the descriptors of
    // the complete records it holds to `recs` (cleared first).  The incomplete tail is kept for the next call.
    // Returns false when the file is exhausted and nothing was produced; throws std::runtime_error on malformed input.
    bool next(std::vector<char> &buf, std::vector<RecordRef> &recs, size_t target);

    bool bgzf() const { return bgzf_data_ != nullptr; } // blocked gzip (bgzip): blocks are inflated in parallel

private:
    size_t fill(char *dst, size_t cap);
    size_t fill_bgzf(char *dst, size_t cap);
    int fd_{-1};
    gzFile gz_{nullptr};
    bool eof_{false};
    std::vector<char> carry_;
    // BGZF: the compressed file is mapped; every block states its compressed size in the 'BC' extra field and its
    // inflated size in its trailer, so a run of blocks can be inflated by all cores straight into the caller's buffer
    const unsigned char *bgzf_data_{nullptr};
    size_t bgzf_size_{0}, bgzf_pos_{0};
    std::vector<char> bgzf_rest_; // tail of a block that did not fit the caller's buffer
    size_t bgzf_rest_pos_{0};
    // bzip2 (the reference reads it through SeqAn3, taxor_search.cpp:181-182): libbz2.so.1.0 is bound at run time (dlopen) --
    // the image has the library but not its header, so the three entry points and the stream struct are declared here
    size_t fill_bz2(char *dst, size_t cap);
    void *bz_{nullptr};           // BzState
    std::string open_error_;
    // single-stream gzip in a regular file: mapped and decoded by GzipStream (inflate_fast.hpp); zlib's gzread (gz_) is kept for
    // input that cannot be mapped and behind TAXOR_GZIP=zlib
    const unsigned char *gzmap_{nullptr};
    size_t gzmap_size_{0};
    void *gzs_{nullptr}; // GzipStream, or (several threads, file of some size) ParallelGzip when gz_parallel_ is set
    bool gz_parallel_{false};
};

// ---- plain (not gzip) files: mapped, cut into byte segments, segments scanned in parallel ----
class MappedFile
{
public:
    explicit MappedFile(const std::string &path);
    ~MappedFile();
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
    bool ok() const { return ok_; }
    bool gzip() const { return gzip_; } // starts with the gzip or bzip2 magic: use RecordScanner instead
    const char *data() const { return data_; }
    size_t size() const { return size_; }

private:
    const char *data_{nullptr};
    size_t size_{0};
    bool ok_{false}, gzip_{false};
};

// A position >= from where a record PROBABLY starts (data[pos] == marker at a line start; for FASTQ also: the line
// two below starts with '+' and lines 1 and 3 have equal lengths -- the 4-line layout).  `size` when there is none.
// Only a guess: the caller checks it against the exact scan of the preceding segment and rescans on disagreement.
size_t guess_record_start(const char *data, size_t size, size_t from, char marker);

// Exact scan (scan semantics of RecordScanner) of the records that start in [begin, end_hint), where `begin` is a
// record start (or blank lines before one).  Descriptors are relative to data + begin.  Returns the position after
// the last scanned record; throws std::runtime_error on malformed input.
size_t scan_segment(const char *data, size_t size, size_t begin, size_t end_hint, std::vector<RecordRef> &recs);

// One byte range [lo, hi) of a mapped file, as the driver's scan jobs and its in-order consumer use it.
struct SegmentScan
{
    size_t begin{0}, end{0};     // first record start (guessed unless lo == 0) / position after the last record scanned
    std::vector<RecordRef> recs; // relative to data + begin
    std::string error;
};
// the job: guess the first record start at or after lo, scan the records that start before hi (never throws)
void scan_byte_range(const char *data, size_t size, size_t first_record, char marker, size_t lo, size_t hi, SegmentScan &out);
// the consumer, called in range order with `expected` = where the next record must start according to the exact
// scan of everything before: keeps `sg` if its guess agrees, rescans it exactly otherwise (throws on malformed input);
// a range lying wholly inside the record before yields no records.  Returns the new `expected`.
size_t accept_byte_range(const char *data, size_t size, size_t expected, size_t hi, SegmentScan &sg);

// The sequence of `r` with its line breaks removed (multi-line records); single-line records need no copy.
void join_record(const char *raw, const RecordRef &r, std::string &out);
// The sequence as SeqAn3's readers deliver it: line breaks dropped, and so are blanks (both formats) and digits (FASTA only)
// inside sequence lines (format_fasta.hpp / format_fastq.hpp filter them before the alphabet check).  Slow path, taken only
// for records whose fast pack hit such a character.
void clean_record(const char *raw, const RecordRef &r, std::string &out);
} // namespace txr
