/* STAND-IN (empty): seqan3::debug_stream is only used in commented-out code of the reference headers we compile. */
#pragma once
#include <algorithm>
#include <iostream>
#include <memory>
#include <ranges>
#include <cassert>
#include <concepts>
#include <string>
#include <vector>
