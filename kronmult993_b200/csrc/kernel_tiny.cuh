// kernel_tiny.cuh -- one thread per batch item, everything in registers ("tiny" path, n^d * sizeof(T) <= 512).
//
// Replaces cuda_kronmult_batchelement + cuda_kronmult (kronmult_gpu/kronmult.cu:139-167, :95-130)
// for the latency/launch-bound shapes (BASELINE config 2: n = 2, d = 2; the reference's `toy` and
// `small` cases).  The reference gives each such item its own n^d-thread block (4 threads for n^d = 4,
// kronmult.cu:188) plus a device-heap allocation (kronmult.cu:156); here a warp covers 32 consecutive
// items, pointer arrays are read coalesced, vectors and factors are fetched with 128-bit loads when
// their addresses allow it, and the d mode products are fully unrolled in registers.  The path is
// HBM-bound (192 B/item of compulsory traffic at n = 2, d = 2, fp64).
#pragma once
#include "common.cuh"

namespace kron
{

template<typename T, int COUNT>
__device__ __forceinline__ void load_contig(const T *__restrict__ p, T (&v)[COUNT])
{
    constexpr int VEC = 16 / sizeof(T);
    if constexpr (COUNT % VEC == 0)
    {
        if (aligned16(p))
        {
            const int4 *q = reinterpret_cast<const int4 *>(p);
#pragma unroll
            for (int i = 0; i < COUNT / VEC; ++i)
            {
                const int4 w = __ldg(q + i);
                if constexpr (sizeof(T) == 8)
                {
                    v[2 * i]     = __hiloint2double(w.y, w.x);
                    v[2 * i + 1] = __hiloint2double(w.w, w.z);
                }
                else
                {
                    v[4 * i]     = __int_as_float(w.x);
                    v[4 * i + 1] = __int_as_float(w.y);
                    v[4 * i + 2] = __int_as_float(w.z);
                    v[4 * i + 3] = __int_as_float(w.w);
                }
            }
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < COUNT; ++i) v[i] = __ldg(p + i);
}

template<typename T, int n, int d>
__global__ void __launch_bounds__(128) kron_tiny_kernel(const T *const *__restrict__ A, T *const *__restrict__ in,
                                                        T *const *__restrict__ out, int lda, int nb)
{
    constexpr int N = ipow(n, d);
    long long k      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < nb;
    if (!valid) k = nb - 1; // computes a duplicate of the last item and drops it at the flush (the warp stays converged)

    T v[N];
    load_contig<T, N>(in[k], v);

#pragma unroll
    for (int j = d - 1; j >= 0; --j)
    {
        const T *__restrict__ M = A[k * d + j];
        T m[n * n]; // m[c*n + r] = M(r, c)
        if (lda == n) { load_contig<T, n * n>(M, m); }
        else
        {
#pragma unroll
            for (int c = 0; c < n; ++c)
#pragma unroll
                for (int r = 0; r < n; ++r) m[c * n + r] = __ldg(M + r + (long long)c * lda);
        }
        constexpr int dummy = 0;
        (void)dummy;
        const int S = ipow(n, d - 1 - j); // stride of the index factor j acts on
#pragma unroll
        for (int f = 0; f < N / n; ++f)
        {
            const int hi   = f / S;
            const int lo   = f - hi * S;
            const int base = hi * S * n + lo;
            T x[n];
#pragma unroll
            for (int kk = 0; kk < n; ++kk) x[kk] = v[base + kk * S];
#pragma unroll
            for (int i = 0; i < n; ++i)
            {
                T dot = T(0);
#pragma unroll
                for (int kk = 0; kk < n; ++kk) dot += x[kk] * m[kk * n + i];
                v[base + i * S] = dot;
            }
        }
    }

    // 32-byte items with distinct outputs: the REDG of a warp already fill whole sectors.  When neighbouring items
    // share an output pointer (aliased batches) the adds are summed per run in shared memory first, like the
    // longer items below.
    bool direct = false;
    if constexpr (N <= 4)
    {
        T *o            = out[k];
        const T *o_next = reinterpret_cast<const T *>(__shfl_down_sync(0xffffffffu, reinterpret_cast<unsigned long long>(o), 1));
        direct          = !__any_sync(0xffffffffu, (threadIdx.x & 31) < 31 && o_next == o);
        if (direct && valid)
        {
#pragma unroll
            for (int i = 0; i < N; ++i) red_add(o + i, v[i]);
        }
    }
    if (!direct)
    {
        // Flush through shared memory: a thread holds a whole item, so REDG straight from registers would hit 32
        // different lines per instruction with 8 (4) bytes each (measured 47 G RED/s against 330 G/s for
        // sector-complete ones).  The warp's 32 results are transposed so that N consecutive lanes add one item's
        // N consecutive elements; each lane group walks a segment of consecutive items and sums runs of equal
        // output pointers before adding.  The kernel has no early exit above, so every lane reaches __syncwarp.
        // Items longer than 32 elements go through the buffer in rounds of 32.
        constexpr int CH     = N < 32 ? N : 32;      // elements per round
        constexpr int ROUNDS = (N + CH - 1) / CH;
        constexpr int PITCH  = CH | 1;               // odd pitch: conflict-free rows
        constexpr int G      = 32 / CH;              // lane groups
        constexpr int SEG    = (32 + G - 1) / G;     // consecutive items per group
        __shared__ T s_val[4][32 * PITCH];
        __shared__ T *s_out[4][32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        T *sv = s_val[warp];
        s_out[warp][lane] = valid ? out[k] : nullptr;
        const int g = lane / CH, i = lane - g * CH;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r)
        {
            if (r > 0) __syncwarp();
#pragma unroll
            for (int e = 0; e < CH; ++e)
                if (r * CH + e < N) sv[lane * PITCH + e] = v[r * CH + e];
            __syncwarp();
            if (g < G && r * CH + i < N)
            {
                T sum  = T(0);
                T *cur = nullptr;
#pragma unroll 4
                for (int t = 0; t < SEG; ++t)
                {
                    const int item = g * SEG + t;
                    if (item >= 32) break;
                    T *o = s_out[warp][item];
                    if (o != cur)
                    {
                        if (cur) red_add(cur + r * CH + i, sum);
                        cur = o;
                        sum = T(0);
                    }
                    sum += sv[item * PITCH + i];
                }
                if (cur) red_add(cur + r * CH + i, sum);
            }
        }
    }
}

} // namespace kron
