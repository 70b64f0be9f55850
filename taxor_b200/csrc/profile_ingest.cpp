// profile_ingest.cpp -- SURVEY 8(f) rank 3: the head of `taxor profile`, fed directly from in-memory search results.
//
// The reference writes every hit to a TSV (taxor_search.cpp:268-306) only for `taxor profile` to parse it back
// (taxor_profile.cpp:93-163) into a std::map keyed by read id and to run three rounds of reference filtering
// (tax_profile, :796-825) before its EM loop.  Here the same table is filled straight from txr_result batches (or, for
// files written earlier, from the TSV), and the three rounds run on it.  Host code only -- strings and small maps; the EM
// loop, the taxonomy roll-up and the CAMI writers stay with the reference binary (out of scope), which is why the table can
// be exported in a line format (txr_profile_text) and as flat arrays (txr_profile_view).
//
// Bit-exactness notes (all reproduced, none "fixed"):
//   * the key is the read id up to its first space (:125-126); reads that share an id share one entry, in arrival order;
//   * a no-hit line is dropped only if the entry already holds something (:156-159), so "no hit" followed by hits of a
//     second read with the same id leaves a "-" element in front of real ones, and later rounds count "-" like a reference;
//   * round 2 compares in float: unique / (unique + ambiguous) >= 0.01f with at least 3 unique reads (:276-279, :819);
//   * round 3 uses unsigned arithmetic (all - shared < u64(0.05 * all)) and, when it re-assigns a hit to the reference
//     that "explains" it, changes accession and reference length but not the tax id (:436-445).
#include "../../include/taxor_b200.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

namespace
{
struct Hit // taxonomy::Search_Result (src/taxonomy/search_results.hpp:8-18) without the read id (it is the key)
{
    std::string accession; // "-" = unclassified
    std::string tax_id;
    uint64_t ref_len{0}, query_len{0}, hash_count{0}, hash_match{0};
};
using ReadTable = std::map<std::string, std::vector<Hit>>; // ordered by read id: the order every round iterates in

std::string key_of(const char *id, size_t n)
{
    const void *sp = memchr(id, ' ', n);
    return std::string(id, sp ? (size_t)(static_cast<const char *>(sp) - id) : n);
}
} // namespace

struct txr_profile
{
    ReadTable reads;
    std::map<std::string, std::pair<std::string, std::string>> taxpath; // accession -> (tax id string, tax name string), first seen
    std::map<std::string, uint64_t> taxa;                               // accession -> reference length, after round 3
    int rounds_done{0};
    std::string text;                                                   // txr_profile_text buffer
    // flat export
    std::vector<const char *> v_read, v_acc, v_tax;
    std::vector<uint64_t> v_begin, v_ref_len, v_query_len, v_hash_count, v_hash_match;

    void add(const std::string &key, Hit &&h) // taxor_profile.cpp:149-161
    {
        std::vector<Hit> &v = reads[key];
        if (!v.empty() && h.accession == "-")
            return; // a null result never joins an entry that already holds something
        v.push_back(std::move(h));
    }
};

namespace
{
thread_local std::string g_profile_error;
int fail(int code, const std::string &msg)
{
    g_profile_error = msg;
    return code;
}

// round 1 helper and body: references that some read maps to uniquely (:166-181); every ambiguous read that touches at
// least one accepted reference keeps only its accepted references (:186-234)
void keep_only_accepted(ReadTable &reads, const std::unordered_set<std::string> &accepted)
{
    for (auto &entry : reads)
    {
        std::vector<Hit> &v = entry.second;
        if (v.size() <= 1)
            continue;
        bool touches = false;
        uint64_t query_len = 0;
        for (const Hit &h : v)
        {
            query_len = h.query_len;
            if (accepted.count(h.accession))
            {
                touches = true;
                break;
            }
        }
        if (touches)
        {
            std::vector<Hit> kept;
            for (Hit &h : v)
            {
                query_len = h.query_len;
                if (accepted.count(h.accession))
                    kept.push_back(std::move(h));
            }
            v.swap(kept);
        }
        if (v.empty()) // cannot happen after `touches`, kept for the reference's shape (:228-232)
        {
            Hit none;
            none.accession = "-";
            none.query_len = query_len;
            v.push_back(std::move(none));
        }
    }
}

void round_unique_refs(ReadTable &reads)
{
    std::unordered_set<std::string> accepted;
    for (auto &entry : reads)
        if (entry.second.size() == 1 && entry.second[0].accession != "-")
            accepted.insert(entry.second[0].accession);
    keep_only_accepted(reads, accepted);
}

// round 2 (:237-282): per reference, reads mapping only to it vs. reads it shares; low-confidence references go
void round_low_confidence(ReadTable &reads, unsigned min_unique, float min_fraction)
{
    std::map<std::string, std::pair<uint64_t, uint64_t>> counts; // accession -> (unique, ambiguous)
    for (auto &entry : reads)
    {
        const std::vector<Hit> &v = entry.second;
        if (v.size() == 1)
        {
            if (v[0].accession != "-")
                counts[v[0].accession].first += 1;
        }
        else
            for (const Hit &h : v)
                counts[h.accession].second += 1; // "-" elements are counted like any reference, as in the reference
    }
    std::unordered_set<std::string> accepted;
    for (auto &c : counts)
        if (c.second.first >= min_unique &&
            static_cast<float>(c.second.first) / static_cast<float>(c.second.first + c.second.second) >= min_fraction)
            accepted.insert(c.first);
    keep_only_accepted(reads, accepted);
}

// round 3 (:289-462): references whose reads are (almost) all shared with a stronger reference are "explained" by it
void round_associations(ReadTable &reads, std::map<std::string, uint64_t> &taxa)
{
    struct Info
    {
        uint64_t unique{0}, all{0};
        std::map<std::string, uint64_t> shared; // other reference -> reads mapped to both
    };
    std::map<std::string, Info> refs;
    taxa.clear();
    for (auto &entry : reads)
    {
        const std::vector<Hit> &v = entry.second;
        if (v.empty())
            continue;
        if (v.size() == 1)
        {
            if (v[0].accession == "-")
                continue;
            Info &i = refs[v[0].accession];
            i.unique += 1;
            i.all += 1;
            taxa.emplace(v[0].accession, v[0].ref_len);
            continue;
        }
        for (const Hit &h : v)
        {
            refs[h.accession].all += 1;
            taxa.emplace(h.accession, h.ref_len);
        }
        for (const Hit &a : v)
            for (const Hit &b : v)
                if (a.accession != b.accession)
                    refs[a.accession].shared[b.accession] += 1;
    }
    std::map<std::string, std::string> explained; // first is explained by second; the first insertion for a key stays
    for (auto &r : refs)
    {
        const Info &me = r.second;
        for (auto &s : me.shared)
        {
            const Info &other = refs.at(s.first);
            if (me.unique > other.unique || me.all > other.all)
            {
                if (me.all - s.second < static_cast<uint64_t>(0.05 * static_cast<double>(me.all)))
                    explained.emplace(r.first, s.first);
            }
            else if (other.all - other.shared.at(r.first) < static_cast<uint64_t>(0.05 * static_cast<double>(other.all)))
                explained.emplace(s.first, r.first);
        }
    }
    for (bool changed = true; changed;) // follow chains (:389-403)
    {
        changed = false;
        for (auto &e : explained)
        {
            auto next = explained.find(e.second);
            if (next != explained.end() && e.first != next->second)
            {
                e.second = next->second;
                changed = true;
            }
        }
    }
    for (auto &entry : reads)
    {
        std::vector<Hit> &v = entry.second;
        if (v.size() <= 1)
            continue;
        std::set<std::string> present;
        for (const Hit &h : v)
            present.insert(h.accession);
        std::vector<Hit> kept;
        for (Hit &h : v)
        {
            auto e = explained.find(h.accession);
            if (e != explained.end())
            {
                if (present.count(e->second))
                    continue; // the explaining reference is among this read's hits already: drop this one
                h.accession = e->second;
                h.ref_len = taxa.at(h.accession); // the tax id keeps its old value (:443-444)
            }
            kept.push_back(std::move(h));
        }
        v.swap(kept);
    }
    for (auto it = taxa.begin(); it != taxa.end();)
        it = explained.count(it->first) ? taxa.erase(it) : std::next(it);
}
} // namespace

extern "C" {

const char *txr_profile_last_error(void) { return g_profile_error.c_str(); }

int txr_profile_create(txr_profile **out)
{
    if (!out)
        return fail(TXR_ERR_ARG, "out is null");
    *out = new txr_profile;
    return TXR_OK;
}
void txr_profile_destroy(txr_profile *p) { delete p; }

int txr_profile_add_batch(txr_profile *p, const txr_result *res, const char *const *read_ids, const uint32_t *read_len,
                          const txr_profile_species *species, uint64_t n_species)
{
    if (!p || !res || !read_ids || !read_len || (!species && n_species))
        return fail(TXR_ERR_ARG, "null argument");
    if (p->rounds_done)
        return fail(TXR_ERR_STATE, "the table has been filtered already");
    // user bin -> species row: the first row that names a user bin wins (std::map::emplace, taxor_search.cpp:172-178)
    std::map<uint64_t, uint64_t> by_bin;
    for (uint64_t i = 0; i < n_species; ++i)
        by_bin.emplace(species[i].user_bin, i);
    for (uint64_t r = 0; r < res->n_reads; ++r)
    {
        if (!read_ids[r])
            return fail(TXR_ERR_ARG, "null read id");
        const std::string key = key_of(read_ids[r], strlen(read_ids[r]));
        bool any = false;
        for (uint64_t i = res->hit_begin[r]; i < res->hit_begin[r + 1]; ++i)
        {
            if (!res->keep[i]) // the result file only has the hits that pass the 0.8 * max filter (taxor_search.cpp:285-286)
                continue;
            auto it = by_bin.find((uint64_t)res->user_bin[i]);
            if (n_species == 0)
                return fail(TXR_ERR_ARG, "no species table");
            const txr_profile_species &sp = species[it == by_bin.end() ? 0 : it->second];
            Hit h;
            h.accession = sp.accession_id ? sp.accession_id : "";
            h.tax_id = sp.taxid ? sp.taxid : "";
            h.ref_len = sp.seq_len;
            h.query_len = read_len[r];
            h.hash_count = res->hash_count[r];
            h.hash_match = res->count[i];
            p->taxpath.emplace(h.accession, std::make_pair(std::string(sp.taxid_string ? sp.taxid_string : ""),
                                                           std::string(sp.taxnames_string ? sp.taxnames_string : "")));
            p->add(key, std::move(h));
            any = true;
        }
        // a read whose hits were all there but none kept cannot occur (the maximum always passes); no hits at all: the
        // 6-column line (taxor_search.cpp:268-273)
        if (!any && res->hit_begin[r] == res->hit_begin[r + 1])
        {
            Hit none;
            none.accession = "-";
            none.query_len = read_len[r];
            p->add(key, std::move(none));
        }
    }
    return TXR_OK;
}

int txr_profile_add_file(txr_profile *p, const char *search_file)
{
    if (!p || !search_file)
        return fail(TXR_ERR_ARG, "null argument");
    if (p->rounds_done)
        return fail(TXR_ERR_STATE, "the table has been filtered already");
    std::ifstream in(search_file);
    if (in.fail())
        return fail(TXR_ERR_IO, std::string("Could not open search results file: ") + search_file);
    std::string line;
    bool header = true;
    uint64_t line_no = 0;
    while (std::getline(in, line))
    {
        ++line_no;
        if (header) // the first line is skipped whatever it holds (:121-122)
        {
            header = false;
            continue;
        }
        std::vector<std::string> col;
        {
            std::stringstream ss(line);
            std::string f;
            while (std::getline(ss, f, '\t'))
                col.push_back(f);
        }
        if (col.size() < 6 || (col[1] != "-" && col.size() < 10))
            return fail(TXR_ERR_FORMAT, std::string(search_file) + ": line " + std::to_string(line_no) + " has too few columns");
        Hit h;
        try
        {
            if (col[1] == "-")
            {
                h.accession = "-";
                h.query_len = std::stoull(col[5]);
            }
            else
            {
                h.accession = col[1];
                h.tax_id = col[3];
                h.ref_len = std::stoull(col[4]);
                h.query_len = std::stoull(col[5]);
                h.hash_count = std::stoull(col[6]);
                h.hash_match = std::stoull(col[7]);
                p->taxpath.emplace(h.accession, std::make_pair(col[9], col[8]));
            }
        }
        catch (std::exception const &)
        {
            return fail(TXR_ERR_FORMAT, std::string(search_file) + ": line " + std::to_string(line_no) + " has a non-numeric length or count");
        }
        p->add(key_of(col[0].data(), col[0].size()), std::move(h));
    }
    return TXR_OK;
}

int txr_profile_filter(txr_profile *p, int rounds)
{
    if (!p || rounds < 0 || rounds > 3)
        return fail(TXR_ERR_ARG, "bad argument");
    for (; p->rounds_done < rounds; ++p->rounds_done)
    {
        if (p->rounds_done == 0)
            round_unique_refs(p->reads);
        else if (p->rounds_done == 1)
            round_low_confidence(p->reads, 3, 0.01f); // tax_profile's arguments (:819)
        else
            round_associations(p->reads, p->taxa);
    }
    return TXR_OK;
}

int txr_profile_text(txr_profile *p, const char **text, uint64_t *len)
{
    if (!p || !text)
        return fail(TXR_ERR_ARG, "null argument");
    std::string &s = p->text;
    s.clear();
    for (auto &e : p->reads)
    {
        s += "R\t" + e.first + "\t" + std::to_string(e.second.size()) + "\n";
        for (const Hit &h : e.second)
            s += "H\t" + h.accession + "\t" + h.tax_id + "\t" + std::to_string(h.ref_len) + "\t" + std::to_string(h.query_len) + "\t" +
                 std::to_string(h.hash_count) + "\t" + std::to_string(h.hash_match) + "\n";
    }
    for (auto &t : p->taxa)
        s += "T\t" + t.first + "\t" + std::to_string(t.second) + "\n";
    for (auto &t : p->taxpath)
        s += "P\t" + t.first + "\t" + t.second.first + "\t" + t.second.second + "\n";
    *text = s.c_str();
    if (len)
        *len = s.size();
    return TXR_OK;
}

int txr_profile_get(txr_profile *p, txr_profile_view *out)
{
    if (!p || !out)
        return fail(TXR_ERR_ARG, "null argument");
    p->v_read.clear();
    p->v_begin.assign(1, 0);
    p->v_acc.clear();
    p->v_tax.clear();
    p->v_ref_len.clear();
    p->v_query_len.clear();
    p->v_hash_count.clear();
    p->v_hash_match.clear();
    for (auto &e : p->reads)
    {
        p->v_read.push_back(e.first.c_str());
        for (const Hit &h : e.second)
        {
            p->v_acc.push_back(h.accession.c_str());
            p->v_tax.push_back(h.tax_id.c_str());
            p->v_ref_len.push_back(h.ref_len);
            p->v_query_len.push_back(h.query_len);
            p->v_hash_count.push_back(h.hash_count);
            p->v_hash_match.push_back(h.hash_match);
        }
        p->v_begin.push_back(p->v_acc.size());
    }
    out->n_reads = p->v_read.size();
    out->read_id = p->v_read.data();
    out->hit_begin = p->v_begin.data();
    out->accession_id = p->v_acc.data();
    out->tax_id = p->v_tax.data();
    out->ref_len = p->v_ref_len.data();
    out->query_len = p->v_query_len.data();
    out->query_hash_count = p->v_hash_count.data();
    out->query_hash_match = p->v_hash_match.data();
    out->n_taxa = p->taxa.size();
    return TXR_OK;
}

} // extern "C"
