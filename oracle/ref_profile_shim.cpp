/*
 * ref_profile_shim.cpp -- C-ABI shim around the REFERENCE'S OWN src/main/taxor_profile.cpp, compiled where it lies under
 * /root/reference (never copied).  TEST INFRASTRUCTURE: outputs go to oracle/_ref/ only.
 *
 * Exposes the head of tax_profile() (src/main/taxor_profile.cpp:796-824): parse_search_results (:93-163) and the three
 * reference-filter rounds -- remove_matches_to_nonunique_refs (:186-234), remove_low_confidence_references (:269-282, with
 * tax_profile's arguments 3 and 0.01), filter_ref_associations (:289-462) -- and serialises what they leave behind, so that
 * the product's txr_profile_* entry points (fed from in-memory hits) can be byte-compared with the reference fed from the
 * result file.  Stops before the EM loop.
 * The SeqAn3 headers taxor_profile.cpp includes are absent from /root/reference; oracle/stubs/seqan3/{utility,argument_parser}
 * hold empty stand-ins (nothing of them is used by the functions called here).
 */
#include <seqan3/utility/views/chunk.hpp> // the stand-in: brings the std headers the reference's own headers rely on
#include <ankerl/unordered_dense.h>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "taxor_profile_configuration.hpp" // /root/reference/src/main
#include "taxor_profile.hpp"
#include "search_results.hpp"              // /root/reference/src/taxonomy

namespace taxor::profile
{
// defined in /root/reference/src/main/taxor_profile.cpp (external linkage, no header declares them)
std::map<std::string, std::vector<taxonomy::Search_Result>> parse_search_results(std::string const filepath,
                                                                                 std::map<std::string, std::pair<std::string, std::string>> & taxpath);
ankerl::unordered_dense::set<std::string> get_refs_with_uniquely_mapping_reads(std::map<std::string, std::vector<taxonomy::Search_Result>> & search_results);
void remove_matches_to_nonunique_refs(std::map<std::string, std::vector<taxonomy::Search_Result>> & search_results,
                                      ankerl::unordered_dense::set<std::string> & ref_unique_mappings);
std::map<std::string, std::pair<uint64_t, uint64_t>> count_unique_ambiguous_mappings_per_reference(
    std::map<std::string, std::vector<taxonomy::Search_Result>> & search_results);
void remove_low_confidence_references(std::map<std::string, std::vector<taxonomy::Search_Result>> & search_results,
                                      std::map<std::string, std::pair<uint64_t, uint64_t>> & map_counts, uint8_t min_unique_mappings,
                                      float min_fraction_unique);
std::map<std::string, size_t> filter_ref_associations(std::map<std::string, std::vector<taxonomy::Search_Result>> & search_results, uint8_t threads);
} // namespace taxor::profile

extern "C" {
/* stage: 0 = after parsing, 1 = after the first filter round, 2 = after the second, 3 = after all three (+ the taxa table).
 * Returns a malloc'ed, NUL-terminated text (caller frees with ref_profile_free); NULL if the file cannot be opened. */
char *ref_profile_prefilter(const char *search_file, int stage)
{
    using namespace taxor::profile;
    std::map<std::string, std::pair<std::string, std::string>> taxpath{};
    std::map<std::string, std::vector<taxor::taxonomy::Search_Result>> results;
    std::map<std::string, size_t> found_taxa;
    try
    {
        results = parse_search_results(search_file, taxpath);                     // :801
        if (stage >= 1)
        {
            auto uniq = get_refs_with_uniquely_mapping_reads(results);            // :807-809
            remove_matches_to_nonunique_refs(results, uniq);
        }
        if (stage >= 2)
        {
            auto counts = count_unique_ambiguous_mappings_per_reference(results); // :815-819
            remove_low_confidence_references(results, counts, 3, 0.01);
        }
        if (stage >= 3)
            found_taxa = filter_ref_associations(results, 1);                     // :825
    }
    catch (std::exception const &)
    {
        return nullptr;
    }
    std::ostringstream os;
    for (auto & pr : results)
    {
        os << "R\t" << pr.first << '\t' << pr.second.size() << '\n';
        for (auto & h : pr.second)
            os << "H\t" << h.accession_id << '\t' << h.tax_id << '\t' << h.ref_len << '\t' << h.query_len << '\t' << h.query_hash_count
               << '\t' << h.query_hash_match << '\n';
    }
    for (auto & t : found_taxa)
        os << "T\t" << t.first << '\t' << t.second << '\n';
    for (auto & p : taxpath)
        os << "P\t" << p.first << '\t' << p.second.first << '\t' << p.second.second << '\n';
    const std::string s = os.str();
    char *out = static_cast<char *>(malloc(s.size() + 1));
    if (out)
        memcpy(out, s.c_str(), s.size() + 1);
    return out;
}
void ref_profile_free(char *p) { free(p); }
}
