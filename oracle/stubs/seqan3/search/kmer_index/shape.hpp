/* STAND-IN for seqan3::shape: threshold_parameters.hpp stores one by value and never reads it. */
#pragma once
#include <cstdint>
namespace seqan3
{
struct shape
{
};
} // namespace seqan3
