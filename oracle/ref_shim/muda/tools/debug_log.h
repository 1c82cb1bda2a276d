// Stand-in for muda/tools/debug_log.h (TEST INFRASTRUCTURE): the two diagnostics macros the reference's distance headers use.
#pragma once
#include <cstdio>
#include <cstdlib>
#define MUDA_ASSERT(cond, ...) do { if (!(cond)) { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); std::abort(); } } while (0)
#define MUDA_ERROR_WITH_LOCATION(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); std::abort(); } while (0)
