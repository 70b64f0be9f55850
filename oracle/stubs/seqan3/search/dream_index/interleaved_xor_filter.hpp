/* STAND-IN for seqan3::interleaved_xor_filter<uint8_t> of the fork JensUweUlrich/seqan3 (absent from
 * /root/reference).  TEST INFRASTRUCTURE so that the reference's own
 * src/hixf/build/hierarchical_interleaved_xor_filter.hpp (membership_agent DFS) compiles in place.
 * *** PARITY UNPINNED ***: bulk_count uses the oracle's restated probe arithmetic (oracle/ixf_ref.h);
 * only the surface the reference calls is provided (counting_agent<value_t>().bulk_count(range),
 * result.size(), result[bin]; see hixf.hpp:307-317). */
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <seqan3/core/concept/cereal.hpp>
#include "../../../../ixf_ref.h"
namespace seqan3
{
template <typename value_t>
class counting_vector : public std::vector<value_t>
{
public:
    using std::vector<value_t>::vector;
};

template <typename FingerprintType = uint8_t>
class interleaved_xor_filter
{
public:
    size_t bins_{}, tbins_{}, seg_len_{}, seed_{};
    FingerprintType const * data_{nullptr};

    interleaved_xor_filter() = default;
    interleaved_xor_filter(size_t bins, size_t tbins, size_t seg_len, size_t seed, FingerprintType const * data) :
        bins_{bins}, tbins_{tbins}, seg_len_{seg_len}, seed_{seed}, data_{data}
    {}
    size_t bin_count() const noexcept { return bins_; }

    template <typename value_t>
    class counting_agent_type
    {
        interleaved_xor_filter const * ixf{nullptr};
        counting_vector<value_t> result_buffer;
    public:
        using counting_vector = seqan3::counting_vector<value_t>;
        counting_agent_type() = default;
        explicit counting_agent_type(interleaved_xor_filter const & f) : ixf{&f}, result_buffer(f.bins_) {}
        template <typename range_t>
        [[nodiscard]] seqan3::counting_vector<value_t> const & bulk_count(range_t && values) & noexcept
        {
            for (auto & c : result_buffer)
                c = 0;
            for (auto && v : values)
            {
                uint64_t const h = ixfref_mix(static_cast<uint64_t>(v), ixf->seed_);
                FingerprintType const f = static_cast<FingerprintType>(ixfref_fingerprint(h));
                FingerprintType const * r0 = ixf->data_ + ixfref_slot(h, 0, ixf->seg_len_) * ixf->tbins_;
                FingerprintType const * r1 = ixf->data_ + ixfref_slot(h, 1, ixf->seg_len_) * ixf->tbins_;
                FingerprintType const * r2 = ixf->data_ + ixfref_slot(h, 2, ixf->seg_len_) * ixf->tbins_;
                for (size_t b = 0; b < ixf->bins_; ++b)
                    result_buffer[b] += static_cast<FingerprintType>(r0[b] ^ r1[b] ^ r2[b]) == f;
            }
            return result_buffer;
        }
    };

    template <typename value_t = uint16_t>
    counting_agent_type<value_t> counting_agent() const
    {
        return counting_agent_type<value_t>{*this};
    }
};
} // namespace seqan3
