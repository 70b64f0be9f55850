/* STAND-IN (test infrastructure): src/main/taxor_profile.cpp includes <seqan3/utility/views/chunk.hpp> but uses nothing of it
 * -- except the standard headers the real SeqAn3 header pulls in, which the reference's own headers rely on. */
#pragma once
#include <cstdint>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <vector>
