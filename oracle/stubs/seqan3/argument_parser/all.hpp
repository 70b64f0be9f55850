/* STAND-IN for seqan3::argument_parser (SeqAn3 fork, absent from /root/reference).  TEST INFRASTRUCTURE: only so that the
 * reference's own src/main/taxor_profile.cpp compiles in place for oracle/_ref; the shim calls its parse/filter functions
 * directly and never parses a command line, so every member here is an empty shell with the signature the reference uses. */
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
namespace seqan3
{
enum class option_spec { standard, required, advanced, hidden };
struct argument_parser_error : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
template <typename T>
struct arithmetic_range_validator
{
    T lo, hi;
};
template <typename T>
arithmetic_range_validator(T, T) -> arithmetic_range_validator<T>;
class argument_parser
{
public:
    struct info_t
    {
        std::string version, author, email, short_description;
        std::vector<std::string> description;
    } info;
    void add_subsection(std::string const &) {}
    template <typename T, typename... Rest>
    void add_option(T &, char, std::string const &, std::string const &, Rest &&...)
    {}
    template <typename... Rest>
    void add_flag(bool &, char, std::string const &, std::string const &, Rest &&...)
    {}
    void parse() {}
};
} // namespace seqan3
