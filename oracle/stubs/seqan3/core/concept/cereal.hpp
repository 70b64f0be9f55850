/* STAND-IN for seqan3/core/concept/cereal.hpp: the concepts only constrain serialize() templates that the
 * oracle/_ref shim never instantiates. */
#pragma once
#ifndef CEREAL_SERIALIZE_FUNCTION_NAME
#define CEREAL_SERIALIZE_FUNCTION_NAME serialize
#endif
namespace seqan3
{
template <typename t> concept cereal_archive = true;
template <typename t> concept cereal_input_archive = true;
template <typename t> concept cereal_output_archive = true;
} // namespace seqan3
