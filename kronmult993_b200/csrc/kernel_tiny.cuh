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

// ---------------------------------------------------------------------------------------------------------------
// Staged variant for items of 128 bytes and more (n^d * sizeof(T) in [128, 512]): BASELINE config 1 (n = 4, d = 3),
// the reference's `small`-like shapes, n = 2 with d >= 4 ...
// ncu on the kernel above (profiles/ncu_pairtile_tiny_r01.md, n = 4, d = 3): the L1 data pipe is 64 % busy at 32 %
// DRAM throughput -- a thread that reads ITS item with 128-bit loads makes every warp instruction touch 32
// different lines.  Here a warp brings its 32 items (and, when they are dense and aligned, their d factors) into
// shared memory with fully coalesced 16-byte cp.async copies -- one warp instruction = 512 contiguous bytes of one
// or more items, nothing held in registers while in flight -- and each lane then reads its own item row.  All
// rows are a multiple of 128 bytes apart, so the lane visits the 16-byte chunks of its row at c ^ (lane & 7):
// the eight lanes of a quarter-warp always hit eight different bank groups (the trick of kernel_sym5.cuh).
template<typename T, int n, int d>
struct TinyStaged
{
    static constexpr int N     = ipow(n, d);
    static constexpr int S     = (int)sizeof(T);
    static constexpr int VB    = N * S;                       // bytes of a vector
    static constexpr int FB    = n * n * S;                   // bytes of a dense factor
    static constexpr int VP    = (VB + 127) / 128 * 128;      // row pitches in shared memory
    static constexpr int FP    = (d * FB + 127) / 128 * 128;
    static constexpr int WARPS = 2;
    static constexpr bool OK   = (VB == 128 || VB == 256 || VB == 512); // measured: 288-byte items (n = 6, d = 2) lose 16 %
    static constexpr bool FOK  = (FB % 16 == 0);              // factors can be staged in 16-byte chunks
    static constexpr int PTRS  = 32 * d * 8;                  // factor pointer table of a warp
    static constexpr int WBYTES = 32 * VP + (FOK ? 32 * FP : 0) + PTRS;
    static constexpr int SMEM  = WARPS * WBYTES;
};

template<typename T, int n, int d>
__global__ void __launch_bounds__(TinyStaged<T, n, d>::WARPS * 32)
kron_tiny_staged_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out, int lda,
                        int nb)
{
    using C = TinyStaged<T, n, d>;
    constexpr int N = C::N, S = C::S, CPI = C::VB / 16, CPF = C::FOK ? C::FB / 16 : 1, VEC = 16 / S;
    extern __shared__ __align__(128) unsigned char tiny_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wb   = tiny_smem + warp * C::WBYTES;
    unsigned char *vbuf = wb;                                   // [32][VP]
    unsigned char *fbuf = wb + 32 * C::VP;                      // [32][FP]
    const T **ptab      = reinterpret_cast<const T **>(wb + 32 * C::VP + (C::FOK ? 32 * C::FP : 0)); // [32 * d]

    const long long k0 = ((long long)blockIdx.x * C::WARPS + warp) * 32;
    if (k0 >= nb) return;                                       // whole warp
    long long k      = k0 + lane;
    const bool valid = k < nb;
    if (!valid) k = nb - 1; // a duplicate of the last item, dropped at the flush (the warp stays converged)

    const T *ip   = in[k];
    const bool va = __all_sync(0xffffffffu, aligned16(ip));
    // factor pointers of the warp's items: entries [k0*d, (k0+32)*d) of A, read coalesced
    bool fa = C::FOK && (lda == n);
#pragma unroll
    for (int i = 0; i < d; ++i)
    {
        const long long e = k0 * d + lane + 32 * i;
        const T *fp       = A[e < (long long)nb * d ? e : (long long)nb * d - 1];
        ptab[lane + 32 * i] = fp;
        fa = fa && aligned16(fp);
    }
    fa = __all_sync(0xffffffffu, fa);
    __syncwarp();
    if (va)
    {
        const unsigned vs = (unsigned)__cvta_generic_to_shared(vbuf);
#pragma unroll
        for (int it = 0; it < CPI; ++it)
        {
            const int q = it * 32 + lane, j = q / CPI, c = q - j * CPI;
            const T *src = reinterpret_cast<const T *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ip), j));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vs + j * C::VP + ((c ^ (j & 7)) << 4)),
                         "l"(src + c * VEC) : "memory");
        }
    }
    if (fa)
    {
        const unsigned fs = (unsigned)__cvta_generic_to_shared(fbuf);
#pragma unroll
        for (int it = 0; it < d * CPF; ++it)
        {
            const int q = it * 32 + lane, j = q / (d * CPF), c = q - j * (d * CPF); // c = factor * CPF + chunk
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(fs + j * C::FP + ((c ^ (j & 7)) << 4)),
                         "l"(ptab[j * d + c / CPF] + (c % CPF) * VEC) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    T *o = out[k]; // in flight while the copies land
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    T v[N];
    if (va)
    {
        const unsigned char *row = vbuf + lane * C::VP;
#pragma unroll
        for (int c = 0; c < CPI; ++c)
        {
            const int4 w = *reinterpret_cast<const int4 *>(row + ((c ^ (lane & 7)) << 4));
            if constexpr (S == 8)
            {
                v[2 * c]     = __hiloint2double(w.y, w.x);
                v[2 * c + 1] = __hiloint2double(w.w, w.z);
            }
            else
            {
                v[4 * c] = __int_as_float(w.x); v[4 * c + 1] = __int_as_float(w.y);
                v[4 * c + 2] = __int_as_float(w.z); v[4 * c + 3] = __int_as_float(w.w);
            }
        }
    }
    else load_contig<T, N>(ip, v);

#pragma unroll
    for (int j = d - 1; j >= 0; --j)
    {
        T m[n * n]; // m[c*n + r] = M(r, c)
        if (fa)
        {
            const unsigned char *row = fbuf + lane * C::FP;
#pragma unroll
            for (int c = 0; c < CPF; ++c)
            {
                const int4 w = *reinterpret_cast<const int4 *>(row + (((j * CPF + c) ^ (lane & 7)) << 4));
                if constexpr (S == 8)
                {
                    m[2 * c]     = __hiloint2double(w.y, w.x);
                    m[2 * c + 1] = __hiloint2double(w.w, w.z);
                }
                else
                {
                    m[4 * c] = __int_as_float(w.x); m[4 * c + 1] = __int_as_float(w.y);
                    m[4 * c + 2] = __int_as_float(w.z); m[4 * c + 3] = __int_as_float(w.w);
                }
            }
        }
        else
        {
            const T *__restrict__ M = A[k * d + j];
            if (lda == n) { load_contig<T, n * n>(M, m); }
            else
            {
#pragma unroll
                for (int c = 0; c < n; ++c)
#pragma unroll
                    for (int r = 0; r < n; ++r) m[c * n + r] = __ldg(M + r + (long long)c * lda);
            }
        }
        const int St = ipow(n, d - 1 - j); // stride of the index factor j acts on
#pragma unroll
        for (int f = 0; f < N / n; ++f)
        {
            const int hi   = f / St;
            const int lo   = f - hi * St;
            const int base = hi * St * n + lo;
            T x[n];
#pragma unroll
            for (int kk = 0; kk < n; ++kk) x[kk] = v[base + kk * St];
#pragma unroll
            for (int i = 0; i < n; ++i)
            {
                T dot = T(0);
#pragma unroll
                for (int kk = 0; kk < n; ++kk) dot += x[kk] * m[kk * n + i];
                v[base + i * St] = dot;
            }
        }
    }

    // flush: the warp's 32 results are transposed through the (now free) staging area so that N consecutive lanes
    // add one item's N consecutive elements (sector-complete REDG); each lane group walks a segment of consecutive
    // items and sums runs of equal output pointers first.  Same scheme as kron_tiny_kernel above.
    {
        constexpr int CH     = N < 32 ? N : 32;      // elements per round
        constexpr int ROUNDS = (N + CH - 1) / CH;
        constexpr int PITCH  = CH | 1;               // odd pitch: conflict-free rows
        constexpr int G      = 32 / CH;              // lane groups
        constexpr int SEG    = (32 + G - 1) / G;     // consecutive items per group
        static_assert(32 * PITCH * S + 8 + 32 * 8 <= C::WBYTES, "the flush buffer fits the warp's staging area");
        T *sv    = reinterpret_cast<T *>(vbuf);
        T **sout = reinterpret_cast<T **>(vbuf + 32 * PITCH * S + ((32 * PITCH * S) % 8 ? 4 : 0));
        __syncwarp(); // every lane has read its rows (vector and factors)
        sout[lane] = valid ? o : nullptr;
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
                    T *oo = sout[item];
                    if (oo != cur)
                    {
                        if (cur) red_add(cur + r * CH + i, sum);
                        cur = oo;
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
