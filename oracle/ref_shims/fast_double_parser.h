// Shim for the un-vendored submodule external_libs/fast_double_parser (empty in /root/reference).
// Returning nullptr makes LightGBM's Common::AtofPrecise fall back to strtod
// (reference include/LightGBM/utils/common.h:359-374): text parsing speed only, no numeric effect.
#pragma once
namespace fast_double_parser {
inline const char* parse_number(const char*, double*) { return nullptr; }
}  // namespace fast_double_parser
