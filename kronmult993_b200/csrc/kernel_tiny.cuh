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

// ---------------------------------------------------------------------------------------------------------------
// Element-wise staged variant for every other item size (n^d * sizeof(T) not 128 / 256 / 512 bytes: n = 3, 5, 6, 7, 9, 10
// ...): the same idea with 4- / 8-byte cp.async copies.  For item j of the warp the 32 lanes copy elements lane,
// lane + 32, ... (one warp instruction = up to 256 contiguous bytes of ONE item instead of 32 scattered pieces), the
// d factors of an item are copied as one flattened run of d*n*n elements (any lda).  Rows are an ODD number of
// elements apart, so a lane's reads of its own row are bank-conflict free without any permutation.
template<typename T, int n, int d>
struct TinyEStaged
{
    static constexpr int N     = ipow(n, d);
    static constexpr int S     = (int)sizeof(T);
    static constexpr int FE    = d * n * n;                   // factor elements per item
    static constexpr int VP    = N | 1;                       // row pitches in elements (odd)
    static constexpr int FP    = FE | 1;
    static constexpr int WARPS = 2;
    static constexpr int WBYTES = ((32 * (VP + FP) * S + 32 * d * 8 + 32 * 8) + 15) / 16 * 16;
    static constexpr int SMEM  = WARPS * WBYTES;
    // Measured on B200 (tools/tiny_session.sh, knob 9 = 1 vs 2): the many small copies pay only where the plain kernel's
    // per-thread loads were worst -- fp64 n = 8, 9 with d = 1 (0.70 -> 0.79, 0.52 -> 0.68: the factor is larger than the
    // vector), n = 3, d = 3 (0.64 -> 0.70), n = 5, d = 2 (0.60 -> 0.64), n = 3, d = 4; everywhere else (all of fp32,
    // n = 6, 7, 10) the plain kernel is faster by up to 2x, so those shapes keep it.
    static constexpr bool LISTED = sizeof(T) == 8 && ((d == 1 && (n == 8 || n == 9)) || (n == 3 && (d == 3 || d == 4)) || (n == 5 && d == 2));
    static constexpr bool OK   = LISTED && !TinyStaged<T, n, d>::OK && SMEM <= 200 * 1024;
};

template<typename T, int n, int d>
__global__ void __launch_bounds__(TinyEStaged<T, n, d>::WARPS * 32)
kron_tiny_estaged_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out, int lda,
                         int nb)
{
    using C = TinyEStaged<T, n, d>;
    constexpr int N = C::N, S = C::S, FE = C::FE, VP = C::VP, FP = C::FP;
    extern __shared__ __align__(16) unsigned char tiny_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wb = tiny_smem + warp * C::WBYTES;
    T *vbuf           = reinterpret_cast<T *>(wb);                         // [32][VP]
    T *fbuf           = vbuf + 32 * VP;                                    // [32][FP]
    const T **ptab    = reinterpret_cast<const T **>(wb + ((32 * (VP + FP) * S + 7) / 8) * 8); // [32 * d] factor pointers
    const T **itab    = ptab + 32 * d;                                     // [32] vector pointers

    const long long k0 = ((long long)blockIdx.x * C::WARPS + warp) * 32;
    if (k0 >= nb) return;                                                  // whole warp
    long long k      = k0 + lane;
    const bool valid = k < nb;
    if (!valid) k = nb - 1; // a duplicate of the last item, dropped at the flush (the warp stays converged)

    itab[lane] = in[k];
#pragma unroll
    for (int i = 0; i < d; ++i)
    {
        const long long e = k0 * d + lane + 32 * i;
        ptab[lane + 32 * i] = A[e < (long long)nb * d ? e : (long long)nb * d - 1];
    }
    __syncwarp();
    const unsigned vs = (unsigned)__cvta_generic_to_shared(vbuf), fs = (unsigned)__cvta_generic_to_shared(fbuf);
#pragma unroll 4
    for (int j = 0; j < 32; ++j)
    {
        const T *src = itab[j];
#pragma unroll
        for (int i = 0; i < (N + 31) / 32; ++i)
        {
            const int e = lane + 32 * i;
            if (e < N)
            {
                if constexpr (S == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(vs + (j * VP + e) * 8), "l"(src + e) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(vs + (j * VP + e) * 4), "l"(src + e) : "memory");
            }
        }
#pragma unroll
        for (int i = 0; i < (FE + 31) / 32; ++i)
        {
            const int e = lane + 32 * i; // element e = factor e / n^2, column (e % n^2) / n, row e % n
            if (e < FE)
            {
                const int f = e / (n * n), rc = e - f * (n * n), c = rc / n, r = rc - c * n;
                const T *g = ptab[j * d + f] + r + (long long)c * lda;
                if constexpr (S == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(fs + (j * FP + e) * 8), "l"(g) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(fs + (j * FP + e) * 4), "l"(g) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    T *o = out[k]; // in flight while the copies land
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    T v[N];
    {
        const T *row = vbuf + lane * VP;
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = row[e];
    }
#pragma unroll
    for (int j = d - 1; j >= 0; --j)
    {
        T m[n * n]; // m[c*n + r] = M(r, c)
        const T *row = fbuf + lane * FP + j * n * n;
#pragma unroll
        for (int e = 0; e < n * n; ++e) m[e] = row[e];
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

    // flush through the (now free) staging area: same scheme as kron_tiny_kernel
    {
        constexpr int CH     = N < 32 ? N : 32;      // elements per round
        constexpr int ROUNDS = (N + CH - 1) / CH;
        constexpr int PITCH  = CH | 1;               // odd pitch: conflict-free rows
        constexpr int G      = 32 / CH;              // lane groups
        constexpr int SEG    = (32 + G - 1) / G;     // consecutive items per group
        static_assert(32 * PITCH * S + 8 + 32 * 8 <= C::WBYTES, "the flush buffer fits the warp's staging area");
        T *sv    = reinterpret_cast<T *>(wb);
        T **sout = reinterpret_cast<T **>(wb + ((32 * PITCH * S + 7) / 8) * 8);
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
