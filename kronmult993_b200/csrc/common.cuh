// common.cuh -- small device/host helpers shared by every kernel family.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

namespace kron
{

// Kernel families (values are the public `kronmult_b200_force_path` codes).
enum Path : int
{
    PATH_AUTO    = 0,
    PATH_GENERIC = 1, // shared-memory tiles, runtime d, any n <= 32, multi-pass for large n^d (fallback beyond n = 10, d = 6)
    PATH_TINY    = 2, // one thread per item, everything in registers (n^d * sizeof(T) <= 512 bytes)
    PATH_REGTILE = 3, // register-tiled in-place mode products, compile-time (n,d), n in {3,4,5,6}
    PATH_DMMA    = 4, // n = 8 on the FP64 tensor pipe (mma.sync m8n8k4), double only
    PATH_WSPEC   = 5, // n = 4, d = 5,6: warp-specialised two-phase kernel with 64-value register tiles
    PATH_WSPEC5  = 6, // n = 4, d = 5: two items per step, split rows, double-buffered exchange
    PATH_PAIRTILE = 7, // compile-time (n, d), n x n register tiles, two factors per shared-memory round trip
    PATH_SYM5    = 8, // n = 4, d = 5: symmetric single-role kernel, one warp per item stream, in-place phases
    PATH_LAST    = PATH_SYM5,
};

__host__ __device__ constexpr int ipow(int b, int e)
{
    int v = 1;
    for (int i = 0; i < e; ++i) v *= b;
    return v;
}

// Thread-safe `*addr += v` without a return value: lowers to RED.E.ADD.{F32,F64} on sm_100a.
// Every final accumulation in this library goes through an atomic-class add so that ANY aliasing of
// output pointers (equal pointers, partial overlap, non-adjacent repeats) stays correct, like the
// reference's element-wise atomicAdd (kronmult_gpu/kronmult.cu:126-129).
// Output vectors always live in the global address space (device or managed memory, kronmult.cuh:22);
// saying so in PTX yields REDG instead of a generic ATOM with a shared-memory CAS fallback branch.
__device__ __forceinline__ void red_add(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// Per-device, per-kernel launch state: the opt-in to more than 48 KiB of dynamic shared memory
// (cudaFuncAttributeMaxDynamicSharedMemorySize) and the occupancy of a kernel belong to a (function, device) pair,
// not to the process -- a host thread that drives several GPUs (kronmult_batched_host_*(..., device)) needs both on
// each of them.  `smem` is the largest dynamic shared-memory size the kernel is ever launched with on this device;
// the attribute is only ever raised.  Thread-safe.
struct KernelState
{
    int smem = -1; // opt-in already granted on this device (bytes), -1: never set
    int occ  = 0;  // resident CTAs per SM for (threads, smem) of the last query
    int occ_threads = 0, occ_smem = -1;
};
inline cudaError_t kernel_setup_any(const void *fn, int threads, int smem, int *ctas_per_sm)
{
    static std::mutex mtx;
    static std::map<std::pair<const void *, int>, KernelState> states;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mtx);
    KernelState &ks = states[std::make_pair(fn, dev)];
    if (smem > ks.smem)
    {
        if (smem > 48 * 1024 || ks.smem > 48 * 1024)
        {
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
        }
        ks.smem = smem;
    }
    if (ctas_per_sm)
    {
        if (ks.occ_threads != threads || ks.occ_smem != smem)
        {
            int occ = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, (size_t)smem);
            if (e != cudaSuccess) return e;
            ks.occ = occ > 0 ? occ : 1;
            ks.occ_threads = threads;
            ks.occ_smem    = smem;
        }
        *ctas_per_sm = ks.occ;
    }
    return cudaSuccess;
}
template<typename K>
inline cudaError_t kernel_setup(K kfn, int threads, int smem, int &ctas_per_sm)
{
    return kernel_setup_any(reinterpret_cast<const void *>(kfn), threads, smem, &ctas_per_sm);
}
template<typename K>
inline cudaError_t kernel_setup(K kfn, int smem)
{
    return kernel_setup_any(reinterpret_cast<const void *>(kfn), 0, smem, nullptr);
}

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace kron
