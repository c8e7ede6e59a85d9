// Test-infrastructure shim for "xxh3.h" (src/Common/XXHasher.hpp:7). The reference links xxHash
// (find_package(xxHash 0.7), CMakeLists.txt:14), which is not vendored and has no dev header in
// this image; pyarrow ships the single-header xxHash 0.8.3, whose XXH3 output is the frozen
// (>= 0.8.0) format. PHJ_XXHASH_HEADER is set by oracle/Makefile.
#pragma once
#define XXH_INLINE_ALL
#include PHJ_XXHASH_HEADER
