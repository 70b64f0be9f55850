/* STAND-IN (test infrastructure): src/main/taxor_profile.cpp includes <seqan3/utility/range/to.hpp> but uses nothing of it. */
#pragma once
