// seqio.hpp -- FASTA / FASTQ (optionally gzip) record reader for the `taxor search` driver.  Replaces
// seqan3::sequence_file_input<dna4_traits, fields<id, seq>> (src/main/taxor_search.cpp:181-182): id = the header
// line without its '>' / '@', sequence = all sequence lines joined; the dna4 conversion happens in txr_pack_2bit.
#pragma once
#include <string>
#include <zlib.h>

namespace txr
{
class SeqReader
{
public:
    explicit SeqReader(const std::string &path);
    ~SeqReader();
    bool ok() const { return gz_ != nullptr; }
    // false at end of file; throws std::runtime_error on malformed input
    bool next(std::string &id, std::string &seq);

private:
    bool getline(std::string &line);
    int peek();
    gzFile gz_{nullptr};
    std::string buf_;
    size_t pos_{0};
    bool eof_{false};
    std::string line_;
};
} // namespace txr
