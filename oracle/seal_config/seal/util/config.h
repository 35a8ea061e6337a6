// TEST / BASELINE INFRASTRUCTURE ONLY.
// Hand-written stand-in for the header SEAL's CMake generates (seal/util/config.h), so that the SEAL 4.0 sources vendored by
// the reference (extern/SEAL, v4.0.0) compile with plain g++ where they lie (oracle/Makefile target `seal`), without CMake and
// without the optional dependencies (GSL, zlib, zstd, HEXL all off). Values are what SEAL's CMake picks for gcc on x86-64.
#pragma once

#define SEAL_VERSION "4.0.0"
#define SEAL_VERSION_MAJOR 4
#define SEAL_VERSION_MINOR 0
#define SEAL_VERSION_PATCH 0

#define SEAL_USE_STD_BYTE
#define SEAL_USE_ALIGNED_ALLOC
#define SEAL_USE_SHARED_MUTEX
#define SEAL_USE_IF_CONSTEXPR
#define SEAL_USE_MAYBE_UNUSED
#define SEAL_USE_NODISCARD
#define SEAL_USE_STD_FOR_EACH_N

#define SEAL_DEFAULT_PRNG Blake2xb

#define SEAL_USE_INTRIN
#define SEAL_USE___BUILTIN_CLZLL
#define SEAL_USE___INT128
#define SEAL_USE__ADDCARRY_U64
#define SEAL_USE__SUBBORROW_U64

#define SEAL_USE_EXPLICIT_BZERO
