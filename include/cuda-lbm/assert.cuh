// assert.cuh — the reference's start-up checks (src/assert.cuh, used by src/core/init/init.cuh:62-64)
#ifndef LBM_ASSERT_H
#define LBM_ASSERT_H
#include <cstdio>
#include <cstdlib>
#define LBM_ASSERT(cond, msg) do { if (!(cond)) { std::fprintf(stderr, "LBM assertion failed: %s (%s:%d)\n", msg, __FILE__, __LINE__); std::exit(EXIT_FAILURE); } } while (0)
#define LBM_DEVICE_ASSERT(cond, msg) LBM_ASSERT(cond, msg)
#endif
