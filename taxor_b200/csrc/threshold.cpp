// threshold.cpp -- see threshold.hpp.  Expressions are kept term-for-term as in the reference so that the same
// libm calls see the same operands (the result is truncated to size_t, so every ulp matters).
#include "threshold.hpp"

#include <cmath>

namespace txr
{
namespace
{
// syncmer_model.hpp:14-36: minimum matching-syncmer ratio by read accuracy (rows 80..100 %) and k (10..30, even)
const double kMatchingRatios[21][10] = {
    {0.552077, 0.195989, 0.151428, 0.118475, 0.0946177, 0.0797244, 0.0604658, 0.0480255, 0.0367569, 0.0252911},
    {0.552385, 0.207533, 0.161204, 0.127368, 0.103704, 0.0881939, 0.0689396, 0.0556991, 0.044185, 0.0298818},
    {0.552239, 0.220393, 0.17382, 0.139866, 0.113736, 0.0966358, 0.0783558, 0.0639223, 0.0523452, 0.0389549},
    {0.552682, 0.236329, 0.188152, 0.152267, 0.126191, 0.106106, 0.0876917, 0.0730642, 0.0621864, 0.0489249},
    {0.553172, 0.254091, 0.202686, 0.165344, 0.137087, 0.116649, 0.098822, 0.0831266, 0.0703342, 0.0582562},
    {0.553716, 0.271183, 0.219848, 0.181959, 0.152163, 0.130048, 0.110622, 0.0942414, 0.0810792, 0.0688187},
    {0.554532, 0.292154, 0.240059, 0.199738, 0.168952, 0.144956, 0.122726, 0.105878, 0.0940805, 0.0777557},
    {0.557957, 0.313553, 0.260912, 0.220014, 0.186567, 0.16101, 0.137399, 0.119867, 0.10453, 0.0900014},
    {0.563925, 0.338316, 0.283689, 0.2401, 0.206963, 0.179541, 0.155347, 0.135128, 0.121575, 0.104741},
    {0.568519, 0.364594, 0.310373, 0.267578, 0.231083, 0.20088, 0.174376, 0.153111, 0.139339, 0.120042},
    {0.579726, 0.395595, 0.338947, 0.295287, 0.258713, 0.22876, 0.200759, 0.175309, 0.161306, 0.139616},
    {0.599258, 0.430241, 0.371291, 0.325596, 0.289651, 0.257329, 0.228011, 0.201799, 0.186956, 0.164794},
    {0.611572, 0.468953, 0.410482, 0.363923, 0.325828, 0.293046, 0.26167, 0.235216, 0.216716, 0.192162},
    {0.624341, 0.510411, 0.452122, 0.407016, 0.370022, 0.334601, 0.303413, 0.275232, 0.254563, 0.227871},
    {0.655724, 0.555245, 0.498564, 0.453201, 0.416285, 0.381883, 0.352291, 0.322556, 0.299739, 0.271481},
    {0.694872, 0.608367, 0.552085, 0.509395, 0.471692, 0.437803, 0.405938, 0.377117, 0.354352, 0.325132},
    {0.742071, 0.669034, 0.613738, 0.57366, 0.539215, 0.50832, 0.476855, 0.449152, 0.42683, 0.397277},
    {0.795543, 0.733694, 0.68341, 0.647737, 0.617382, 0.588448, 0.56083, 0.533714, 0.514757, 0.486399},
    {0.853121, 0.802585, 0.763169, 0.733734, 0.708902, 0.684331, 0.660171, 0.637633, 0.621567, 0.596993},
    {0.918163, 0.882314, 0.854479, 0.835831, 0.819643, 0.804269, 0.788526, 0.771895, 0.763059, 0.742114},
    {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0}};

double nmut_q(double r, size_t k) { return 1.0 - pow(1.0 - r, k); }
double nmut_expected(double r, size_t k, size_t n) { return n * nmut_q(r, k); }       // kmer_model.cpp:28-32
double nmut_variance(double r, size_t k, size_t n)                                    // kmer_model.cpp:34-41
{
    const double q = nmut_q(r, k);
    const double kd = (double)k;
    return (double)n * (1.0 - q) * (q * (2.0 * kd + (2.0 / r) - 1.0) - 2.0 * kd)
         + kd * (kd - 1.0) * pow((1.0 - q), 2.0)
         + (2.0 * (1.0 - q) / (pow(r, 2.0))) * ((1.0 + (kd - 1.0) * (1.0 - q)) * r - q);
}
double nmut_expected_sq(double r, size_t k, size_t n)                                 // kmer_model.cpp:43-46
{
    return pow(nmut_expected(r, k, n), 2) + nmut_variance(r, k, n);
}
double containment_expected(double r, size_t k) { return pow((1.0 - r), k); }         // fracminhash_model.cpp:9-12
double containment_variance(double r, size_t k, size_t n, double sf)                  // fracminhash_model.cpp:14-21
{
    const double term3 = nmut_variance(r, k, n) / pow(n, 2);
    const double term2 = n * nmut_expected(r, k, n) - nmut_expected_sq(r, k, n);
    const double denominator = sf * pow(n, 3) * pow(1.0 - pow(1.0 - sf, n), 2);
    const double term1 = (1.0 - sf) / denominator;
    return term1 * term2 + term3;
}
} // namespace

double normal_cdf_inverse(double p)
{
    // Abramowitz-Stegun rational approximation, gaussian_inverse.cpp:12-23
    auto rational = [](double t)
    {
        const double c[] = {2.515517, 0.802853, 0.010328};
        const double d[] = {1.432788, 0.189269, 0.001308};
        return t - ((c[2] * t + c[1]) * t + c[0]) / (((d[2] * t + d[1]) * t + d[0]) * t + 1.0);
    };
    if (p <= 0.0 || p >= 1.0)
        return NAN; // the reference throws; unreachable with confidence 0.95
    return p < 0.5 ? -rational(sqrt(-2.0 * log(p))) : rational(sqrt(-2.0 * log(1.0 - p)));
}

void nmut_kmer_ci(double r, size_t k, size_t n, double confidence, size_t &low, size_t &high)
{
    const double q = nmut_q(r, k);
    const double varN = nmut_variance(r, k, n);
    const double alpha = 1 - confidence;
    const double z = normal_cdf_inverse(1.0 - alpha / 2.0);
    low = static_cast<size_t>(floor(n * q - z * sqrt(varN)));
    high = static_cast<size_t>(ceil(n * q + z * sqrt(varN)));
}

double syncmer_match_ratio(size_t kmer_size, double error_rate)
{
    const size_t row = ceil((1.0 - error_rate) * 100.0 - 80.0);
    const size_t col = kmer_size - 10 - ((kmer_size - 10) / 2) - 1;
    if (row > 20 || col > 9)
        return NAN; // outside the table: the reference's asserts are compiled out and it reads out of bounds
    return kMatchingRatios[row][col];
}

Thresholder::Thresholder(uint32_t window_size, uint8_t kmer_size, double percentage, double error_rate,
                         bool use_syncmer, bool fracminhash)
    : kmer_size_{kmer_size}, error_rate_{error_rate}
{
    const size_t kmers_per_window = (size_t)window_size - kmer_size + 1;
    if (percentage > 0.0 && percentage <= 1.0)
    {
        kind_ = ThresholdKind::percentage;
        percentage_ = percentage;
    }
    else if (use_syncmer)
        kind_ = ThresholdKind::syncmer_model;
    else if (kmers_per_window == 1 && !fracminhash)
        kind_ = ThresholdKind::kmer_model;
    else
        kind_ = ThresholdKind::fracminhash;
}

const char *Thresholder::banner() const
{
    switch (kind_)
    {
    case ThresholdKind::percentage: return "use percentage-model";
    case ThresholdKind::syncmer_model: return "use syncmer model";
    case ThresholdKind::kmer_model: return "use kmer-model";
    default: return "use frac minhash";
    }
}

size_t Thresholder::get(size_t minimiser_count, double scaling_factor) const noexcept
{
    const size_t fp_correction = minimiser_count * 0.0039;
    switch (kind_)
    {
    case ThresholdKind::syncmer_model:
        return static_cast<size_t>(minimiser_count * syncmer_match_ratio(kmer_size_, error_rate_));
    case ThresholdKind::kmer_model:
    {
        size_t lo, hi;
        nmut_kmer_ci(error_rate_, (size_t)kmer_size_, minimiser_count, 0.95, lo, hi);
        return minimiser_count - hi - fp_correction;
    }
    case ThresholdKind::fracminhash:
    {
        const double z = normal_cdf_inverse(1.0 - (1.0 - 0.95) / 2.0);
        const double clow = containment_expected(error_rate_, kmer_size_)
                          - z * sqrt(containment_variance(error_rate_, kmer_size_, minimiser_count, scaling_factor));
        return static_cast<size_t>(clow * minimiser_count) - fp_correction;
    }
    default:
        return static_cast<size_t>(minimiser_count * percentage_);
    }
}
} // namespace txr
