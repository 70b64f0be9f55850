// hixf_file.cpp -- see hixf_file.hpp.  cereal's portable-less binary archive (RECALLED, cereal 1.3.x): arithmetic
// types raw little-endian, bool one byte, std::string / std::vector prefixed by a u64 element count, vectors of
// arithmetic types as one raw block, classes with serialize() without any header.
#include "hixf_file.hpp"
#include "ixf_arith.cuh"

#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sstream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace txr
{
struct FileMapping
{
    const uint8_t *base{nullptr};
    size_t size{0};
    ~FileMapping()
    {
        if (base)
            munmap(const_cast<uint8_t *>(base), size);
    }
};

IxfRecordSpec IxfRecordSpec::parse(const std::string &csv)
{
    IxfRecordSpec s;
    std::stringstream ss(csv);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty())
            s.scalars.push_back(tok);
    return s;
}

std::string IxfRecordSpec::str() const
{
    std::string r;
    for (auto &x : scalars)
        r += (r.empty() ? "" : ",") + x;
    return r;
}

bool IxfSchemeSpec::parse(const std::string &text, IxfSchemeSpec &out, std::string &error)
{
    out = IxfSchemeSpec{};
    const size_t colon = text.find(':');
    const std::string head = text.substr(0, colon);
    if (head == "xor3" || head.empty())
        out.slots = kIxfSlotsXor3;
    else if (head == "fuse3")
        out.slots = kIxfSlotsFuse3;
    else
    {
        error = "unknown filter scheme '" + head + "' (xor3 | fuse3)";
        return false;
    }
    if (colon == std::string::npos)
        return true;
    std::stringstream ss(text.substr(colon + 1));
    std::string kv;
    while (std::getline(ss, kv, ','))
    {
        const size_t eq = kv.find('=');
        const std::string k = kv.substr(0, eq), v = eq == std::string::npos ? "" : kv.substr(eq + 1);
        if (k == "mix" && (v == "add" || v == "xor"))
            out.mix = v == "xor" ? kIxfMixXorSeed : kIxfMixAddSeed;
        else if (k == "fp" && (v == "fold32" || v == "low8" || v == "high8"))
            out.fingerprint = v == "fold32" ? kIxfFpFold32 : v == "low8" ? kIxfFpLow8 : kIxfFpHigh8;
        else if (k == "rot" && sscanf(v.c_str(), "%u/%u", &out.rot1, &out.rot2) == 2 && out.rot1 < 64 && out.rot2 < 64)
            ;
        else if (k == "layout" && (v == "slot" || v == "bin"))
            out.layout = v == "bin" ? 1u : 0u;
        else
        {
            error = "bad filter scheme option '" + kv + "' (mix=add|xor, fp=fold32|low8|high8, rot=R1/R2, layout=slot|bin)";
            return false;
        }
    }
    return true;
}

std::string IxfSchemeSpec::str() const
{
    std::string r = slots == kIxfSlotsFuse3 ? "fuse3" : "xor3";
    r += std::string(":mix=") + (mix == kIxfMixXorSeed ? "xor" : "add");
    r += std::string(",fp=") + (fingerprint == kIxfFpFold32 ? "fold32" : fingerprint == kIxfFpLow8 ? "low8" : "high8");
    r += ",rot=" + std::to_string(rot1) + "/" + std::to_string(rot2);
    r += std::string(",layout=") + (layout ? "bin" : "slot");
    return r;
}

const std::vector<IxfRecordSpec> &IxfRecordSpec::candidates()
{
    // By analogy with upstream seqan3::interleaved_bloom_filter (bins, technical_bins, bin_size_, hash_shift,
    // bin_words, hash_funs, data).  The first entry is what write_hixf() emits; the others are plausible orders of
    // the fork that the reader probes.  Extend this table (or pass --ixf-record) once the fork header is visible.
    static const std::vector<IxfRecordSpec> c = {
        parse("bins,tbins,slots,bin_words,max_elems,seed"),
        parse("bins,tbins,slots,bin_words,seed"),
        parse("bins,tbins,slots,max_elems,bin_words,seed"),
        parse("bins,tbins,slots,bin_words,max_elems,seed,ftype"),
        parse("bins,tbins,slots,bin_words,max_elems,seg_len,seed"),
        parse("bins,tbins,max_elems,slots,bin_words,seed"),
        parse("bins,tbins,seg_len,bin_words,max_elems,seed"),
        parse("bins,tbins,slots,seed"),
        parse("seed,bins,tbins,slots,bin_words,max_elems"),
    };
    return c;
}

namespace
{
struct Writer
{
    FILE *f;
    bool ok{true};
    void raw(const void *p, size_t n)
    {
        if (ok && n && fwrite(p, 1, n, f) != n)
            ok = false;
    }
    template <typename T> void pod(T v) { raw(&v, sizeof v); }
    void str(const std::string &s)
    {
        pod<uint64_t>(s.size());
        raw(s.data(), s.size());
    }
    void vec_i64(const std::vector<int64_t> &v)
    {
        pod<uint64_t>(v.size());
        raw(v.data(), v.size() * 8);
    }
};

struct Reader
{
    const uint8_t *p, *end;
    bool ok{true};
    bool need(size_t n)
    {
        if (!ok || (size_t)(end - p) < n)
            ok = false;
        return ok;
    }
    template <typename T> T pod()
    {
        T v{};
        if (need(sizeof v))
        {
            memcpy(&v, p, sizeof v);
            p += sizeof v;
        }
        return v;
    }
    uint64_t count(size_t elem_bytes) // a u64 length that must fit into what is left of the file
    {
        const uint64_t n = pod<uint64_t>();
        if (ok && elem_bytes && n > (uint64_t)(end - p) / elem_bytes)
            ok = false;
        return ok ? n : 0;
    }
    std::string str()
    {
        const uint64_t n = count(1);
        std::string s;
        if (need(n))
        {
            s.assign(reinterpret_cast<const char *>(p), n);
            p += n;
        }
        return s;
    }
    std::vector<int64_t> vec_i64()
    {
        const uint64_t n = count(8);
        std::vector<int64_t> v;
        if (need(n * 8))
        {
            v.resize(n);
            if (n)
                memcpy(v.data(), p, n * 8);
            p += n * 8;
        }
        return v;
    }
};

// parses everything from the IXF vector to the end of the file with one candidate record order.
// capacity_ok (optional): every IXF carries a non-zero max_elems and its geometry is what the scheme derives from it
std::string parse_hixf_tail(Reader rd, TaxorIndexFile &idx, const IxfRecordSpec &spec, const IxfSchemeSpec &sc, bool *capacity_ok)
{
    const IxfScheme scheme{sc.slots, sc.mix, sc.fingerprint, sc.rot1, sc.rot2};
    bool cap_ok = true;
    // every record is at least its scalars and the length of its fingerprint vector: a count beyond that is damage,
    // not a reason to allocate
    const uint64_t n_ixf = rd.count(8 * (spec.scalars.size() + 1));
    if (!rd.ok || n_ixf == 0)
        return "IXF vector length unreadable";
    idx.ixf.assign(n_ixf, IxfRecord{});
    for (uint64_t i = 0; i < n_ixf; ++i)
    {
        IxfRecord &x = idx.ixf[i];
        uint64_t slots = 0, bin_words = 0;
        bool have_slots = false, have_words = false, have_seg = false, have_ftype = false;
        for (auto &name : spec.scalars)
        {
            const uint64_t v = rd.pod<uint64_t>();
            if (name == "bins") x.bins = v;
            else if (name == "tbins") x.tbins = v;
            else if (name == "slots") { slots = v; have_slots = true; }
            else if (name == "seg_len") { x.seg_len = v; have_seg = true; }
            else if (name == "bin_words") { bin_words = v; have_words = true; }
            else if (name == "max_elems") x.max_elems = v;
            else if (name == "seed") x.seed = v;
            else if (name == "ftype") { x.ftype = v; have_ftype = true; }
            else if (name != "skip") return "unknown scalar '" + name + "' in the IXF record order";
        }
        const uint64_t len = rd.count(1);
        if (!rd.need(len))
            return "IXF " + std::to_string(i) + ": fingerprint vector runs past the end of the file";
        x.fp = rd.p;
        x.fp_len = len;
        rd.p += len;
        if (have_slots && !have_seg)
        {
            if (scheme.slots != kIxfSlotsXor3)
                return "a binary-fuse record needs both 'seg_len' and 'slots' in the record order";
            x.seg_len = slots / 3;
        }
        if (!have_slots)
            slots = 3 * x.seg_len;
        x.rows = slots;
        if (x.bins == 0 || x.tbins < x.bins || x.tbins % 64 != 0 || x.tbins > (1u << 20))
            return "IXF " + std::to_string(i) + ": implausible bins/technical bins";
        uint64_t count_len = 0;
        if (!ixf_geometry_ok(scheme, x.seg_len, x.rows, count_len))
            return "IXF " + std::to_string(i) + (scheme.slots == kIxfSlotsXor3 ? ": slot count is not three equal segments"
                                                                                : ": slots are not whole power-of-two segments");
        if (x.max_elems == 0)
            cap_ok = false;
        else
        {
            const IxfGeometry g = ixf_geometry_for(scheme, x.max_elems);
            cap_ok = cap_ok && g.seg_len == x.seg_len && g.rows == x.rows;
        }
        if (have_ftype && x.ftype != 8 && x.ftype != 16)
            return "IXF " + std::to_string(i) + ": fingerprint width is neither 8 nor 16";
        if (have_words && bin_words * 64 != x.tbins)
            return "IXF " + std::to_string(i) + ": bin_words does not match technical bins";
        if (len != slots * x.tbins)
            return "IXF " + std::to_string(i) + ": fingerprint bytes != slots x technical bins";
    }
    const uint64_t n_next = rd.count(8);
    if (!rd.ok || n_next != n_ixf)
        return "next_ixf_id does not have one vector per IXF";
    idx.next_ixf_id.clear();
    for (uint64_t i = 0; i < n_next; ++i)
        idx.next_ixf_id.push_back(rd.vec_i64());
    const uint64_t n_names = rd.count(8);
    idx.user_bin_filenames.clear();
    for (uint64_t i = 0; i < n_names && rd.ok; ++i)
        idx.user_bin_filenames.push_back(rd.str());
    const uint64_t n_pos = rd.count(8);
    if (!rd.ok || n_pos != n_ixf)
        return "ixf_bin_to_filename_position does not have one vector per IXF";
    idx.ixf_bin_to_filename_position.clear();
    for (uint64_t i = 0; i < n_pos; ++i)
        idx.ixf_bin_to_filename_position.push_back(rd.vec_i64());
    if (!rd.ok)
        return "truncated file";
    if (rd.p != rd.end)
        return std::to_string((size_t)(rd.end - rd.p)) + " trailing bytes: the file does not tile";
    for (uint64_t i = 0; i < n_ixf; ++i)
    {
        if (idx.next_ixf_id[i].size() != idx.ixf[i].bins || idx.ixf_bin_to_filename_position[i].size() != idx.ixf[i].bins)
            return "IXF " + std::to_string(i) + ": per-bin vectors do not match the bin count";
        for (int64_t ub : idx.ixf_bin_to_filename_position[i])
            if (ub >= (int64_t)n_names)
                return "user bin id beyond user_bin_filenames";
    }
    if (capacity_ok)
        *capacity_ok = cap_ok;
    return "";
}
} // namespace

std::string write_hixf(const std::string &path, const TaxorIndexFile &idx, const IxfRecordSpec &spec)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f)
        return "cannot open " + path + " for writing";
    Writer w{f};
    w.pod<uint32_t>(idx.version);                 // index.hpp:211-212
    w.pod<uint64_t>(idx.window_size);             // :217
    w.pod<uint64_t>(idx.shape_size);              // :218 seqan3::shape (dynamic_bitset<58>): size, bits
    w.pod<uint64_t>(idx.shape_bits);
    w.pod<uint8_t>(idx.kmer_size);                // :219-222
    w.pod<uint8_t>(idx.syncmer_size);
    w.pod<uint8_t>(idx.t_syncmer);
    w.pod<uint8_t>(idx.parts);
    w.pod<uint8_t>(idx.use_syncmer ? 1 : 0);      // :223
    w.pod<uint16_t>(idx.scaling);                 // :224
    w.pod<uint8_t>(idx.compressed ? 1 : 0);       // :225
    w.pod<uint64_t>(idx.bin_path.size());         // :231
    for (auto &v : idx.bin_path)
    {
        w.pod<uint64_t>(v.size());
        for (auto &s : v)
            w.str(s);
    }
    w.pod<uint64_t>(idx.species.size());          // :232, Species.hpp:43-49
    for (auto &s : idx.species)
    {
        w.str(s.organism_name);
        w.str(s.accession_id);
        w.str(s.taxid);
        w.str(s.taxnames_string);
        w.str(s.taxid_string);
        w.pod<uint64_t>(s.user_bin);
        w.pod<uint64_t>(s.seq_len);
    }
    w.pod<uint64_t>(idx.ixf.size());              // hixf.hpp:155
    for (auto &x : idx.ixf)
    {
        for (auto &name : spec.scalars)
        {
            uint64_t v = 0;
            if (name == "bins") v = x.bins;
            else if (name == "tbins") v = x.tbins;
            else if (name == "slots") v = x.rows ? x.rows : 3 * x.seg_len;
            else if (name == "seg_len") v = x.seg_len;
            else if (name == "bin_words") v = x.tbins / 64;
            else if (name == "max_elems") v = x.max_elems;
            else if (name == "seed") v = x.seed;
            else if (name == "ftype") v = x.ftype;
            w.pod<uint64_t>(v);
        }
        w.pod<uint64_t>(x.fp_len);
        w.raw(x.fp, x.fp_len);
    }
    w.pod<uint64_t>(idx.next_ixf_id.size());      // hixf.hpp:156
    for (auto &v : idx.next_ixf_id)
        w.vec_i64(v);
    w.pod<uint64_t>(idx.user_bin_filenames.size()); // hixf.hpp:280-281
    for (auto &s : idx.user_bin_filenames)
        w.str(s);
    w.pod<uint64_t>(idx.ixf_bin_to_filename_position.size());
    for (auto &v : idx.ixf_bin_to_filename_position)
        w.vec_i64(v);
    const bool ok = w.ok;
    if (fclose(f) != 0 || !ok)
        return "write error on " + path;
    return "";
}

std::string read_hixf(const std::string &path, TaxorIndexFile &idx, const IxfRecordSpec *spec, IxfRecordSpec *used,
                      const IxfSchemeSpec *scheme_in, HixfReadReport *report)
{
    const IxfSchemeSpec scheme = scheme_in ? *scheme_in : IxfSchemeSpec{};
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0)
        return "cannot open " + path;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 32)
    {
        close(fd);
        return path + " is too small to be a taxor index";
    }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED)
        return "cannot map " + path;
    auto map = std::make_shared<FileMapping>();
    map->base = static_cast<const uint8_t *>(m);
    map->size = (size_t)st.st_size;
    idx = TaxorIndexFile{};
    idx.mapping = map;
    Reader rd{map->base, map->base + map->size};
    idx.version = rd.pod<uint32_t>();
    if (idx.version != 1)
        return "unsupported index version " + std::to_string(idx.version) + " (index.hpp:48 expects 1)";
    idx.window_size = rd.pod<uint64_t>();
    idx.shape_size = rd.pod<uint64_t>();
    idx.shape_bits = rd.pod<uint64_t>();
    idx.kmer_size = rd.pod<uint8_t>();
    idx.syncmer_size = rd.pod<uint8_t>();
    idx.t_syncmer = rd.pod<uint8_t>();
    idx.parts = rd.pod<uint8_t>();
    idx.use_syncmer = rd.pod<uint8_t>() != 0;
    idx.scaling = rd.pod<uint16_t>();
    idx.compressed = rd.pod<uint8_t>() != 0;
    const uint64_t n_bp = rd.count(8);
    for (uint64_t i = 0; i < n_bp && rd.ok; ++i)
    {
        const uint64_t k = rd.count(8);
        std::vector<std::string> v;
        for (uint64_t j = 0; j < k && rd.ok; ++j)
            v.push_back(rd.str());
        idx.bin_path.push_back(std::move(v));
    }
    const uint64_t n_sp = rd.count(8 * 7);
    for (uint64_t i = 0; i < n_sp && rd.ok; ++i)
    {
        SpeciesRecord s;
        s.organism_name = rd.str();
        s.accession_id = rd.str();
        s.taxid = rd.str();
        s.taxnames_string = rd.str();
        s.taxid_string = rd.str();
        s.user_bin = rd.pod<uint64_t>();
        s.seq_len = rd.pod<uint64_t>();
        idx.species.push_back(std::move(s));
    }
    if (!rd.ok)
        return "truncated header (parameters / bin paths / species table)";
    if (idx.kmer_size == 0 || idx.kmer_size > 32)
        return "implausible k-mer size " + std::to_string(idx.kmer_size);
    // Every candidate order is tried.  "The file tiles" cannot tell orders apart that only permute free u64 scalars (seed
    // vs. max_elems), so orders whose max_elems reproduces the stored geometry through the scheme's capacity formula
    // (rows == geometry(max_elems), ixf_arith.cuh) win; what is left ambiguous is reported, not hidden.
    std::string errors;
    const std::vector<IxfRecordSpec> one = spec ? std::vector<IxfRecordSpec>{*spec} : std::vector<IxfRecordSpec>{};
    const std::vector<IxfRecordSpec> &cands = spec ? one : IxfRecordSpec::candidates();
    int first_tiling = -1, first_consistent = -1;
    unsigned n_tiling = 0, n_consistent = 0;
    std::string tiling_list;
    for (size_t ci = 0; ci < cands.size(); ++ci)
    {
        bool cap_ok = false;
        TaxorIndexFile probe;
        const std::string e = parse_hixf_tail(rd, probe, cands[ci], scheme, &cap_ok);
        if (e.empty())
        {
            ++n_tiling;
            tiling_list += (tiling_list.empty() ? "" : " | ") + cands[ci].str();
            if (first_tiling < 0)
                first_tiling = (int)ci;
            if (cap_ok)
            {
                ++n_consistent;
                if (first_consistent < 0)
                    first_consistent = (int)ci;
            }
        }
        else
            errors += "\n  [" + cands[ci].str() + "] " + e;
    }
    if (first_tiling >= 0)
    {
        const int pick = first_consistent >= 0 ? first_consistent : first_tiling;
        parse_hixf_tail(rd, idx, cands[(size_t)pick], scheme, nullptr);
        if (used)
            *used = cands[(size_t)pick];
        if (report)
        {
            report->used = cands[(size_t)pick];
            report->tiling_candidates = n_tiling;
            report->capacity_consistent = n_consistent;
            const unsigned ambiguous = first_consistent >= 0 ? n_consistent : n_tiling;
            if (ambiguous > 1)
                report->note = std::to_string(ambiguous) + " record orders fit this file equally well (" + tiling_list +
                               "); using [" + cands[(size_t)pick].str() + "] -- seed and capacity may be swapped, pass --ixf-record to pin the order";
            else if (first_consistent < 0 && !spec)
                report->note = "record order [" + cands[(size_t)pick].str() + "] accepted on tiling alone (no max_elems/geometry cross-check possible)";
        }
        return "";
    }
    return "the interleaved-XOR-filter records of " + path + " do not parse with any known field order "
           "(the order is defined by the SeqAn3 fork; pass --ixf-record to override):" + errors;
}
} // namespace txr
