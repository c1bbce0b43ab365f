// kernel_dmma.cuh -- placeholder until the n = 8 FP64 tensor-pipe kernel lands.
#pragma once
#include "common.cuh"
#include <atomic>
namespace kron
{
template<typename T>
static cudaError_t run_dmma(int, int, int, const T *const *, int, T *const *, T *const *, int, cudaStream_t,
                            std::atomic<long long> &, const char *&)
{
    return cudaErrorNotSupported;
}
} // namespace kron
