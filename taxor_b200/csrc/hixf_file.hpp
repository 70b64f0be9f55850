// hixf_file.hpp -- reader / writer of the `.hixf` index file (cereal BinaryOutputArchive of
// taxor_index<hixf_t>, version 1; src/main/index.hpp:208-244, store_index.hpp:20-27, load_index.hpp:27-38)
// without cereal: a little-endian stream parser following SURVEY Appendix A.
//
// *** PARITY UNPINNED at one record ***: the field order of seqan3::interleaved_xor_filter<uint8_t>::serialize
// lives in the un-vendored SeqAn3 fork.  That sub-record is therefore described by DATA (IxfRecordSpec: an ordered
// list of u64 scalars followed by one length-prefixed fingerprint vector); the reader can try several candidate
// orders and accepts one only if every IXF is self-consistent and the whole file tiles exactly.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace txr
{
struct SpeciesRecord // src/taxonomy/Species.hpp:14-21, serialised fields :43-49
{
    std::string organism_name, accession_id, taxid, taxnames_string, taxid_string;
    uint64_t user_bin{0}, seq_len{0};
};

struct IxfRecord
{
    uint64_t seed{0}, bins{0}, tbins{0}, seg_len{0}, max_elems{0}, ftype{8};
    uint64_t rows{0};           // slots per bin (0: 3 * seg_len)
    const uint8_t *fp{nullptr}; // fp[slot * tbins + bin], `rows` slots: points into `owned` or into the mapped file
    uint64_t fp_len{0};
    std::vector<uint8_t> owned;
};

struct FileMapping; // read-only mmap of an index file, kept alive by the TaxorIndexFile that points into it

struct TaxorIndexFile
{
    // index.hpp:32-43 in serialisation order (:211-232)
    uint32_t version{1};
    uint64_t window_size{20};
    uint64_t shape_size{0}, shape_bits{0}; // seqan3::shape = dynamic_bitset<58>: size, bits
    uint8_t kmer_size{0}, syncmer_size{0}, t_syncmer{0}, parts{1};
    bool use_syncmer{true};
    uint16_t scaling{1};
    bool compressed{false};
    std::vector<std::vector<std::string>> bin_path;
    std::vector<SpeciesRecord> species;
    // hixf.hpp:152-158
    std::vector<IxfRecord> ixf;
    std::vector<std::vector<int64_t>> next_ixf_id;
    // hixf.hpp:277-282
    std::vector<std::string> user_bin_filenames;
    std::vector<std::vector<int64_t>> ixf_bin_to_filename_position;
    std::shared_ptr<FileMapping> mapping;
};

// Scalar names: bins, tbins, slots (= 3*seg_len), seg_len, bin_words (= tbins/64), max_elems, seed, ftype, skip.
struct IxfRecordSpec
{
    std::vector<std::string> scalars;
    static IxfRecordSpec parse(const std::string &csv);
    static const std::vector<IxfRecordSpec> &candidates(); // first entry = the order the writer uses
    std::string str() const;
};

// The filter arithmetic the records are checked against (ixf_arith.cuh; mirrors txr_ixf_scheme of the C ABI).
// Text form: "xor3" | "fuse3", then optional ":key=value,..." with mix=add|xor, fp=fold32|low8|high8, rot=R1/R2, layout=slot|bin.
struct IxfSchemeSpec
{
    uint32_t slots{0}, mix{0}, fingerprint{0}, rot1{21}, rot2{42}, layout{0};
    static bool parse(const std::string &text, IxfSchemeSpec &out, std::string &error);
    std::string str() const;
};

// What the reader found out about the record order (the free scalars of a candidate order cannot all be told apart by
// "the file tiles": `seed` and `max_elems` are both arbitrary u64s).
struct HixfReadReport
{
    IxfRecordSpec used;              // the accepted order
    unsigned tiling_candidates{0};   // candidate orders under which every record is self-consistent and the file tiles
    unsigned capacity_consistent{0}; // ... of which also satisfy rows == geometry(max_elems) in every IXF (0 if none carries max_elems)
    std::string note;                // human-readable summary of the above (empty when the choice was unambiguous)
};

// Both return an empty string on success, else a description of what failed.
std::string write_hixf(const std::string &path, const TaxorIndexFile &idx, const IxfRecordSpec &spec);
// spec == nullptr: try every candidate; `used` (optional) receives the accepted order
std::string read_hixf(const std::string &path, TaxorIndexFile &idx, const IxfRecordSpec *spec, IxfRecordSpec *used,
                      const IxfSchemeSpec *scheme = nullptr, HixfReadReport *report = nullptr);
} // namespace txr
