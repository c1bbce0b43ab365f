// common.cuh -- small device/host helpers shared by every kernel family.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
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
    PATH_SYM4    = 9, // n = 4, d = 4: the same with a half-warp per item (two item streams per warp)
    PATH_LAST    = PATH_SYM4,
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

// ---- L2-resident chunking of the multi-pass routes (vectors that do not fit one CTA's shared memory) ----------
// A multi-pass route writes an intermediate vector after its first pass and reads it back in the next one.  Run
// over the whole batch, pass by pass, that intermediate makes a round trip through HBM (3x the algorithmic bytes
// for two passes, 5x for three).  Instead the batch is cut into chunks of items whose vectors fit the 126 MB L2
// together, and ALL passes of a chunk run before the next chunk starts: the later passes read the intermediate
// from L2, and once the last pass has consumed it, the (dirty, never needed again) lines are dropped from L2 with
// discard.global.L2 before they can be written back -- input and workspace may be clobbered (kronmult.cuh:23), so
// their contents after the call are unspecified anyway.
// Measured on B200 (profiles/multipass_chunks_r02.md): a chunk's kernels are short (tens of microseconds), so the
// chunks have to overlap on several streams to pay; then the three-pass route (n = 10, d = 6, fp64) gains 29 %,
// the two-pass pairtile routes 0-7 %, and the HBM-bound DMMA route (n = 8, d = 6) loses 25 % -- its two long
// kernels stream at 4.2 TB/s, which the short ones do not reach.  The discard costs an extra launch per chunk and
// bought nothing measurable.  Hence the default: chunk only routes of three or more passes, no discard.
//   knob 6: chunk size in MiB (0 = whole batch at once, the round-1 behaviour; -1 = automatic, the default)
//   knob 7: 1 = discard the intermediate after the last pass of a chunk (default 0)
inline std::atomic<int> &multipass_chunk_mib() { static std::atomic<int> v{-1}; return v; }
inline std::atomic<int> &multipass_discard() { static std::atomic<int> v{0}; return v; }

// items per chunk for vectors of `bytes_item` bytes (always a multiple of `quantum` items when possible, so that
// runs of equal output pointers are not cut more often than necessary)
inline long long multipass_chunk_items(long long nb, long long bytes_item, int passes, long long quantum = 32)
{
    long long mib = multipass_chunk_mib().load(std::memory_order_relaxed);
    if (mib < 0) mib = passes >= 3 ? 32 : 0;
    if (mib <= 0) return nb;
    long long cb = (mib << 20) / (bytes_item > 0 ? bytes_item : 1);
    if (cb >= quantum) cb = cb / quantum * quantum;
    if (cb < 1) cb = 1;
    return cb < nb ? cb : nb;
}

// Chunks run round-robin on a small pool of internal streams (forked from / joined into the caller's stream with
// events): the first pass of chunk c+1 fills the SMs that the last pass of chunk c leaves idle while it drains, and
// the launch gaps disappear -- a chunk's kernels are too short (tens of microseconds) to run back to back alone.
//   knob 8: number of pool streams (1 = everything on the caller's stream)
inline std::atomic<int> &multipass_streams() { static std::atomic<int> v{3}; return v; }
struct ChunkStreams
{
    static constexpr int MAXS = 4;
    cudaStream_t caller = nullptr;
    cudaStream_t s[MAXS] = {};
    int ns = 1, next = 0;
    bool forked = false;
    cudaError_t begin(cudaStream_t st, long long nchunks)
    {
        caller = st;
        ns     = multipass_streams().load(std::memory_order_relaxed);
        if (ns > MAXS) ns = MAXS;
        if (ns < 2 || nchunks < 2) { ns = 1; return cudaSuccess; }
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        static std::mutex mtx;
        static cudaStream_t pool[64][MAXS] = {};
        {
            std::lock_guard<std::mutex> lk(mtx);
            if (dev < 0 || dev >= 64) { ns = 1; return cudaSuccess; }
            for (int i = 0; i < ns; ++i)
            {
                if (!pool[dev][i])
                {
                    e = cudaStreamCreateWithFlags(&pool[dev][i], cudaStreamNonBlocking);
                    if (e != cudaSuccess) return e;
                }
                s[i] = pool[dev][i];
            }
        }
        cudaEvent_t fork;
        e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(fork, caller);
        for (int i = 0; i < ns && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(s[i], fork, 0);
        cudaEventDestroy(fork); // released once the waits have consumed it
        forked = (e == cudaSuccess);
        return e;
    }
    cudaStream_t pick() { return ns == 1 ? caller : s[next++ % ns]; }
    cudaError_t end()
    {
        if (!forked) return cudaSuccess;
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < ns && e == cudaSuccess; ++i)
        {
            cudaEvent_t join;
            e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
            if (e != cudaSuccess) break;
            e = cudaEventRecord(join, s[i]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(caller, join, 0);
            cudaEventDestroy(join);
        }
        forked = false;
        return e;
    }
};

// drops the 128-byte lines that lie entirely inside [v[k], v[k] + N) for every item of the chunk from L2
template<typename T>
__global__ void discard_vectors_kernel(T *const *__restrict__ v, int cnt, long long N)
{
    const long long lines_max = (N * (long long)sizeof(T)) / 128 + 1;
    for (int k = blockIdx.y; k < cnt; k += gridDim.y)
    {
        const uintptr_t b = reinterpret_cast<uintptr_t>(v[k]);
        const uintptr_t lo = (b + 127) & ~uintptr_t(127), hi = (b + (uintptr_t)N * sizeof(T)) & ~uintptr_t(127);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < lines_max; i += (long long)gridDim.x * blockDim.x)
        {
            const uintptr_t a = lo + (uintptr_t)i * 128;
            if (a + 128 <= hi) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
        }
    }
}
template<typename T>
inline cudaError_t launch_discard(T *const *v, int cnt, long long N, cudaStream_t st)
{
    if (cnt <= 0 || !multipass_discard().load(std::memory_order_relaxed)) return cudaSuccess;
    const long long lines = (N * (long long)sizeof(T)) / 128 + 1;
    int gx = (int)((lines + 255) / 256);
    if (gx > 64) gx = 64;
    int gy = cnt < 65535 ? cnt : 65535;
    while ((long long)gx * gy > 148LL * 32 && gy > 1) gy = (gy + 1) / 2;
    discard_vectors_kernel<T><<<dim3(gx, gy), 256, 0, st>>>(v, cnt, N);
    return cudaGetLastError();
}

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace kron
