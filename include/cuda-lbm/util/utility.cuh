// util/utility.cuh — host/device helpers scenario code may call (reference src/util/utility.cuh, src/util/utility.cu:4-12).
#pragma once
#ifndef UTILITY_H
#define UTILITY_H
#include <curand_kernel.h>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <type_traits>

// uniform deviate in (0, 1] / in (min, max] from a cuRAND state owned by the caller
__device__ inline float gpu_rand(curandState& state) { return curand_uniform(&state); }
__device__ inline float gpu_rand(curandState& state, float min, float max) { return min + (max - min) * curand_uniform(&state); }
inline double random_double() { return rand() / (RAND_MAX + 1.0); }

// The reference treats every CUDA error as fatal: message, cudaDeviceReset, exit(99) (src/util/utility.cu:4-12).
// Kept for code that calls the CUDA runtime directly; the engine itself reports errors through the C ABI's return codes.
inline void check_cuda(cudaError_t result, char const* const func, const char* const file, int const line) {
    if (result != cudaSuccess) {
        std::cerr << "CUDA error = " << static_cast<unsigned int>(result) << " (" << cudaGetErrorString(result) << ") at " << file << ":" << line
                  << " '" << func << "'\n";
        cudaDeviceReset();
        std::exit(99);
    }
}
#define checkCudaErrors(val) check_cuda((val), #val, __FILE__, __LINE__)

#ifdef DEBUG_KERNEL
#define DPRINTF(fmt, ...) printf(fmt, __VA_ARGS__)
#else
#define DPRINTF(fmt, ...)
#endif

#endif  // UTILITY_H
