// hixf_tools.cpp -- C entry points of libtaxor_tools.so around hixf_file.cpp, for tests and benchmarks:
// write a synthetic index as a real `.hixf` file and read one back into plain arrays.
#include "hixf_file.hpp"
#include "ingest.hpp"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace txr;

namespace
{
thread_local std::string g_err;
}

extern "C" {

const char *txs_last_error(void) { return g_err.c_str(); }

// species arrays: n_species entries; strings are NUL-terminated.  ixf arrays as in txr_hixf_view.
int txs_hixf_write(const char *path, const char *record_spec, uint64_t window_size, uint8_t k, uint8_t s, uint8_t t,
                   uint8_t use_syncmer, uint16_t scaling, uint64_t n_ixf, const uint64_t *seed, const uint64_t *bins,
                   const uint64_t *tbins, const uint64_t *seg_len, const uint8_t *const *data, const uint64_t *bin_off,
                   const int64_t *next_ixf_id, const int64_t *bin_to_ub, uint64_t n_user_bins, const char *const *ub_filenames,
                   uint64_t n_species, const char *const *organism, const char *const *accession, const char *const *taxid,
                   const char *const *taxnames, const char *const *taxids, const uint64_t *sp_user_bin, const uint64_t *sp_seq_len,
                   const uint64_t *rows, const uint64_t *max_elems) // rows / max_elems may be null (3*seg_len / 0)
{
    TaxorIndexFile f;
    f.window_size = window_size;
    f.shape_size = k;
    f.shape_bits = k >= 64 ? ~0ull : ((1ull << k) - 1); // ungapped shape: k ones
    f.kmer_size = k;
    f.syncmer_size = s;
    f.t_syncmer = t;
    f.use_syncmer = use_syncmer != 0;
    f.scaling = scaling;
    for (uint64_t i = 0; i < n_user_bins; ++i)
    {
        f.user_bin_filenames.emplace_back(ub_filenames[i]);
        f.bin_path.push_back({std::string(ub_filenames[i])}); // taxor_build.cpp:515
    }
    for (uint64_t i = 0; i < n_species; ++i)
    {
        SpeciesRecord r;
        r.organism_name = organism[i];
        r.accession_id = accession[i];
        r.taxid = taxid[i];
        r.taxnames_string = taxnames[i];
        r.taxid_string = taxids[i];
        r.user_bin = sp_user_bin[i];
        r.seq_len = sp_seq_len[i];
        f.species.push_back(std::move(r));
    }
    for (uint64_t i = 0; i < n_ixf; ++i)
    {
        IxfRecord x;
        x.seed = seed[i];
        x.bins = bins[i];
        x.tbins = tbins[i];
        x.seg_len = seg_len[i];
        x.fp = data[i];
        x.rows = rows ? rows[i] : 3 * seg_len[i];
        x.max_elems = max_elems ? max_elems[i] : 0;
        x.fp_len = x.rows * tbins[i];
        f.ixf.push_back(std::move(x));
        f.next_ixf_id.emplace_back(next_ixf_id + bin_off[i], next_ixf_id + bin_off[i + 1]);
        f.ixf_bin_to_filename_position.emplace_back(bin_to_ub + bin_off[i], bin_to_ub + bin_off[i + 1]);
    }
    const IxfRecordSpec spec = record_spec && *record_spec ? IxfRecordSpec::parse(record_spec) : IxfRecordSpec::candidates()[0];
    g_err = write_hixf(path, f, spec);
    return g_err.empty() ? 0 : -1;
}

static std::string g_note;
void *txs_hixf_open2(const char *path, const char *record_spec, const char *scheme_text)
{
    auto *f = new TaxorIndexFile;
    IxfRecordSpec used;
    IxfSchemeSpec scheme;
    HixfReadReport rep;
    g_note.clear();
    if (scheme_text && *scheme_text && !IxfSchemeSpec::parse(scheme_text, scheme, g_err))
    {
        delete f;
        return nullptr;
    }
    if (record_spec && *record_spec)
    {
        const IxfRecordSpec spec = IxfRecordSpec::parse(record_spec);
        g_err = read_hixf(path, *f, &spec, &used, &scheme, &rep);
    }
    else
        g_err = read_hixf(path, *f, nullptr, &used, &scheme, &rep);
    if (!g_err.empty())
    {
        delete f;
        return nullptr;
    }
    g_err = used.str();
    g_note = rep.note;
    return f;
}
void *txs_hixf_open(const char *path, const char *record_spec) { return txs_hixf_open2(path, record_spec, nullptr); }
// the reader's remark about the last successful open (ambiguous record orders); empty when there was nothing to say
const char *txs_hixf_open_note(void) { return g_note.c_str(); }
uint64_t txs_hixf_ixf_rows(void *p, uint64_t i)
{
    const IxfRecord &x = static_cast<TaxorIndexFile *>(p)->ixf[i];
    return x.rows ? x.rows : 3 * x.seg_len;
}
void txs_hixf_close(void *p) { delete static_cast<TaxorIndexFile *>(p); }

// scalars: [version, window, shape_size, shape_bits, k, s, t, parts, use_syncmer, scaling, compressed, n_ixf, n_user_bins, n_species]
void txs_hixf_info(void *p, uint64_t *out14)
{
    auto *f = static_cast<TaxorIndexFile *>(p);
    const uint64_t v[14] = {f->version, f->window_size, f->shape_size, f->shape_bits, f->kmer_size, f->syncmer_size, f->t_syncmer,
                            f->parts, f->use_syncmer, f->scaling, f->compressed, f->ixf.size(), f->user_bin_filenames.size(),
                            f->species.size()};
    memcpy(out14, v, sizeof v);
}
void txs_hixf_ixf(void *p, uint64_t i, uint64_t *seed, uint64_t *bins, uint64_t *tbins, uint64_t *seg_len, const uint8_t **fp,
                  const int64_t **next, const int64_t **ub)
{
    auto *f = static_cast<TaxorIndexFile *>(p);
    const IxfRecord &x = f->ixf[i];
    *seed = x.seed;
    *bins = x.bins;
    *tbins = x.tbins;
    *seg_len = x.seg_len;
    *fp = x.fp;
    *next = f->next_ixf_id[i].data();
    *ub = f->ixf_bin_to_filename_position[i].data();
}
const char *txs_hixf_species_field(void *p, uint64_t i, int field, uint64_t *user_bin, uint64_t *seq_len)
{
    auto *f = static_cast<TaxorIndexFile *>(p);
    const SpeciesRecord &s = f->species[i];
    *user_bin = s.user_bin;
    *seq_len = s.seq_len;
    switch (field)
    {
    case 0: return s.organism_name.c_str();
    case 1: return s.accession_id.c_str();
    case 2: return s.taxid.c_str();
    case 3: return s.taxnames_string.c_str();
    default: return s.taxid_string.c_str();
    }
}

// Test hook for the CLI's parallel ingest (ingest.cpp): scans `path` with raw buffers of `target` bytes and writes one
// "id<TAB>sequence" line per record to `out_path`.  Returns the number of records, -1 on error (txs_last_error).
int64_t txs_ingest_dump(const char *path, uint64_t target, const char *out_path)
{
    try
    {
        txr::RecordScanner scan(path);
        if (!scan.ok())
        {
            g_err = std::string("cannot open ") + path;
            return -1;
        }
        FILE *f = fopen(out_path, "wb");
        if (!f)
        {
            g_err = std::string("cannot write ") + out_path;
            return -1;
        }
        std::vector<char> buf;
        std::vector<txr::RecordRef> recs;
        std::string joined;
        int64_t n = 0;
        while (scan.next(buf, recs, target))
            for (const auto &r : recs)
            {
                fwrite(buf.data() + r.id_off, 1, r.id_len, f);
                fputc('\t', f);
                if (r.single_line)
                    fwrite(buf.data() + r.seq_off, 1, r.seq_len, f);
                else
                {
                    txr::join_record(buf.data(), r, joined);
                    fwrite(joined.data(), 1, joined.size(), f);
                }
                fputc('\n', f);
                ++n;
            }
        fclose(f);
        return n;
    }
    catch (std::exception const &e)
    {
        g_err = e.what();
        return -1;
    }
}

// The mapped-file path of the same ingest: byte ranges of `seg_bytes`, each guessed + scanned on its own, then
// accepted or rescanned in order -- exactly what the driver's jobs and assembler do, single threaded.
// `n_rescans` receives how many guesses were wrong.
int64_t txs_ingest_dump_mapped(const char *path, uint64_t seg_bytes, const char *out_path, uint64_t *n_rescans)
{
    try
    {
        txr::MappedFile mf(path);
        if (!mf.ok())
        {
            g_err = std::string("cannot open ") + path;
            return -1;
        }
        FILE *f = fopen(out_path, "wb");
        if (!f)
        {
            g_err = std::string("cannot write ") + out_path;
            return -1;
        }
        const char *data = mf.data();
        const size_t size = mf.size();
        size_t first = 0;
        while (first < size && (data[first] == '\n' || data[first] == '\r'))
            ++first;
        int64_t n = 0;
        uint64_t rescans = 0;
        if (first < size)
        {
            if (data[first] != '>' && data[first] != '@')
                throw std::runtime_error("sequence file: record does not start with '>' or '@'");
            const char marker = data[first];
            const size_t n_seg = (size + seg_bytes - 1) / seg_bytes;
            std::vector<txr::SegmentScan> segs(n_seg);
            for (size_t k = 0; k < n_seg; ++k) // "parallel" phase
                txr::scan_byte_range(data, size, first, marker, k * seg_bytes, std::min<size_t>(size, (k + 1) * seg_bytes), segs[k]);
            size_t expected = first;
            std::string joined;
            for (size_t k = 0; k < n_seg; ++k)
            {
                const size_t hi = std::min<size_t>(size, (k + 1) * seg_bytes);
                const bool inside = expected >= hi;
                const bool agreed = segs[k].error.empty() && segs[k].begin == expected;
                expected = txr::accept_byte_range(data, size, expected, hi, segs[k]);
                if (!inside && !agreed)
                    ++rescans;
                const char *base = data + segs[k].begin;
                for (const auto &r : segs[k].recs)
                {
                    fwrite(base + r.id_off, 1, r.id_len, f);
                    fputc('\t', f);
                    if (r.single_line)
                        fwrite(base + r.seq_off, 1, r.seq_len, f);
                    else
                    {
                        txr::join_record(base, r, joined);
                        fwrite(joined.data(), 1, joined.size(), f);
                    }
                    fputc('\n', f);
                    ++n;
                }
            }
        }
        fclose(f);
        if (n_rescans)
            *n_rescans = rescans;
        return n;
    }
    catch (std::exception const &e)
    {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
