/* STAND-IN for the SeqAn3 nucleotide alphabets (fork JensUweUlrich/seqan3 absent from /root/reference).
 * TEST INFRASTRUCTURE: the reference's syncmer.cpp only needs dna5_vector, size(), operator[] and to_char(). */
#pragma once
#include <vector>
namespace seqan3
{
struct dna5
{
    char c{'A'};
    constexpr char to_char() const noexcept { return c; }
};
using dna5_vector = std::vector<dna5>;
} // namespace seqan3
