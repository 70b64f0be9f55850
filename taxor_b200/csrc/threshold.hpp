// threshold.hpp -- host-side mirror of hixf::threshold::threshold (src/hixf/search/threshold.hpp:22-81) and its
// models.  Double arithmetic stays on the host (libm, same expressions as the reference) and is shipped to the
// GPU as an integer lookup table indexed by hash_count.
#pragma once
#include <cstddef>
#include <cstdint>

namespace txr
{
enum class ThresholdKind { fracminhash, percentage, kmer_model, syncmer_model };

class Thresholder
{
public:
    Thresholder() = default;
    // threshold.hpp:22-49
    Thresholder(uint32_t window_size, uint8_t kmer_size, double percentage, double error_rate, bool use_syncmer,
                bool fracminhash = false);
    // threshold.hpp:51-81
    size_t get(size_t minimiser_count, double scaling_factor) const noexcept;
    ThresholdKind kind() const { return kind_; }
    const char *banner() const; // the line the reference prints when the thresholder is built (:32-47)
    double percentage() const { return percentage_; }

private:
    ThresholdKind kind_{ThresholdKind::percentage};
    uint8_t kmer_size_{};
    double percentage_{};
    double error_rate_{};
};

double syncmer_match_ratio(size_t kmer_size, double error_rate);          // syncmer_model.hpp:38-50
double normal_cdf_inverse(double p);                                      // gaussian_inverse.cpp:28-52
void nmut_kmer_ci(double r, size_t k, size_t n, double confidence, size_t &low, size_t &high); // kmer_model.cpp:10-26
} // namespace txr
