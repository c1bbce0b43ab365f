// kernel_generic.cuh -- the shape-agnostic kernel family ("generic" path).
//
// Replaces, for every (n, d) the specialised kernels do not cover, the reference's
// cuda_kronmult_batchelement / cuda_kronmult / transpose / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :34-43, :54-78).
//
// Differences in design (not a port):
//  * The vector is staged once into shared memory and the d mode products are applied IN PLACE:
//    factor j acts on the index with stride n^(d-1-j), each thread owns whole length-n fibers, so
//    there is no transposition, no rotation of the index order and no global ping-pong between
//    `input` and `workspace` (the reference moves 2*N elements through global memory per factor,
//    kronmult.cu:112-121).  Per output element the sum still runs k ascending from 0, as in
//    multiply_transpose (kronmult.cu:66-70), so single-item results are bit-identical to an
//    FMA-contracted build of the reference.
//  * A CTA processes B item "streams" side by side (so tiny n^d still fills 256 threads, where the
//    reference launches n^d-thread blocks, kronmult.cu:188) and `chunk` consecutive items per stream;
//    results of consecutive items that share an output pointer are summed in shared memory and
//    added to global memory once per run, always with an atomic-class add.
//  * n^d too large for shared memory is handled by several passes over groups of factors, each
//    pass working on (n^G x LB)-element tiles in place in `input` (which the contract allows to be
//    clobbered, kronmult.cuh:23); the last pass accumulates into `output`.
#pragma once
#include "common.cuh"

namespace kron
{

template<typename T>
struct PassParams
{
    const T *const *A;
    T *const *in;
    T *const *dst;            // where a non-final pass stores its result: == in (in place) or the scratch vectors
    T *const *out;
    int d, n, lda, nb;
    int j0, G;                // this pass applies factors j0 .. j0+G-1 (highest index first)
    int Mext;                 // n^G: extent of the indices touched by this pass
    int LB;                   // tile width along the faster indices not touched (divides L)
    long long L;              // n^(d-j0-G)
    int lblocks;              // L / LB
    long long tiles_per_item; // (N / (Mext*L)) * lblocks
    int tile_elems;           // Mext * LB
    int fibers;               // tile_elems / n
    int final_pass;           // 1: add into out, 0: store back into in
    int use_acc;              // 1: sum runs of equal output pointers in shared memory first
    int B;                    // item streams per CTA
    int chunk;                // consecutive items per stream
    long long units;          // (#item groups) * tiles_per_item
    int off_acc, off_mats, off_ptrs; // byte offsets into dynamic shared memory
};

// y = M x along one fiber of stride S, in place.  M is row-major n x n in shared memory.
template<typename T, int NT>
__device__ __forceinline__ void apply_fiber(T *x, int S, const T *M, int n)
{
    if constexpr (NT > 0)
    {
        T v[NT];
#pragma unroll
        for (int k = 0; k < NT; ++k) v[k] = x[(size_t)k * S];
#pragma unroll
        for (int i = 0; i < NT; ++i)
        {
            T dot = T(0);
#pragma unroll
            for (int k = 0; k < NT; ++k) dot += v[k] * M[i * NT + k];
            x[(size_t)i * S] = dot;
        }
    }
    else
    {
        T v[32];
        for (int k = 0; k < n; ++k) v[k] = x[(size_t)k * S];
        for (int i = 0; i < n; ++i)
        {
            T dot = T(0);
            for (int k = 0; k < n; ++k) dot += v[k] * M[i * n + k];
            x[(size_t)i * S] = dot;
        }
    }
}

template<typename T, int NT>
__global__ void __launch_bounds__(256) kron_pass_kernel(const PassParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    T *buf         = reinterpret_cast<T *>(smem);
    T *acc         = reinterpret_cast<T *>(smem + p.off_acc);
    T *mats        = reinterpret_cast<T *>(smem + p.off_mats);
    T **s_in       = reinterpret_cast<T **>(smem + p.off_ptrs);
    T **s_out      = s_in + p.B;
    T **s_dst      = s_out + p.B;
    const T **s_A  = static_cast<const T **>(static_cast<void *>(s_dst + p.B));
    int *s_flag    = reinterpret_cast<int *>(s_A + p.B * p.G);
    const int n    = NT > 0 ? NT : p.n;
    const int nn   = n * n;
    const int tid  = threadIdx.x;
    const int nthr = blockDim.x;
    const int te   = p.tile_elems;
    const long long group_items = (long long)p.B * p.chunk;

    for (long long u = blockIdx.x; u < p.units; u += gridDim.x)
    {
        const long long g         = u / p.tiles_per_item;
        const long long tile      = u - g * p.tiles_per_item;
        const long long h         = tile / p.lblocks;
        const int lb              = (int)(tile - h * p.lblocks);
        const long long tile_base = h * (long long)p.Mext * p.L + (long long)lb * p.LB;
        const long long item0     = g * group_items;

        for (int t = 0; t < p.chunk; ++t)
        {
            __syncthreads(); // everyone is done with the previous step's pointers, matrices and tile
            for (int b = tid; b < p.B; b += nthr)
            {
                const long long k = item0 + (long long)b * p.chunk + t;
                int flag = 0;
                if (k < p.nb)
                {
                    T *o     = p.out[k];
                    s_in[b]  = p.in[k];
                    s_dst[b] = p.dst[k];
                    s_out[b] = o;
                    // bit1: first item of a run of equal output pointers (within this stream)
                    // bit2: last item of the run -> flush the shared-memory sum
                    const bool first = (t == 0) || (p.out[k - 1] != o);
                    const bool last  = (t == p.chunk - 1) || (k + 1 >= p.nb) || (p.out[k + 1] != o);
                    flag = 1 | (first ? 2 : 0) | (last ? 4 : 0);
                }
                s_flag[b] = flag;
            }
            for (int e = tid; e < p.B * p.G; e += nthr)
            {
                const int b       = e / p.G;
                const int jj      = e - b * p.G;
                const long long k = item0 + (long long)b * p.chunk + t;
                s_A[e] = (k < p.nb) ? p.A[k * p.d + p.j0 + jj] : nullptr;
            }
            __syncthreads();
            if (s_flag[0] == 0) break; // stream 0 holds the smallest item index: nothing left

            // factor matrices -> shared memory, row-major (global reads run down the columns)
            for (int e = tid; e < p.B * p.G * nn; e += nthr)
            {
                const int m  = e / nn;
                const int r  = e - m * nn;
                const int kk = r / n;
                const int i  = r - kk * n;
                const T *Ap  = s_A[m];
                if (Ap) mats[m * nn + i * n + kk] = Ap[i + (long long)kk * p.lda];
            }
            // the tile of every stream
            for (int e = tid; e < p.B * te; e += nthr)
            {
                const int b = e / te;
                if (!s_flag[b]) continue;
                const int r = e - b * te;
                const int m = r / p.LB;
                const int l = r - m * p.LB;
                buf[e] = s_in[b][tile_base + (long long)m * p.L + l];
            }
            __syncthreads();

            int S = p.LB;
            for (int jj = p.G - 1; jj >= 0; --jj)
            {
                for (int f = tid; f < p.B * p.fibers; f += nthr)
                {
                    const int b = f / p.fibers;
                    if (!s_flag[b]) continue;
                    const int ff = f - b * p.fibers;
                    const int hi = ff / S;
                    const int lo = ff - hi * S;
                    apply_fiber<T, NT>(buf + (size_t)b * te + (size_t)hi * S * n + lo, S,
                                       mats + (size_t)(b * p.G + jj) * nn, n);
                }
                __syncthreads();
                S *= n;
            }

            // epilogue: element e is handled by the same thread in every step, so the shared-memory
            // accumulator needs no synchronisation of its own
            for (int e = tid; e < p.B * te; e += nthr)
            {
                const int b    = e / te;
                const int flag = s_flag[b];
                if (!flag) continue;
                const int r = e - b * te;
                const int m = r / p.LB;
                const int l = r - m * p.LB;
                const long long gi = tile_base + (long long)m * p.L + l;
                const T v = buf[e];
                if (!p.final_pass) { s_dst[b][gi] = v; }
                else if (p.use_acc)
                {
                    const T a = (flag & 2) ? v : acc[e] + v;
                    if (flag & 4) red_add(s_out[b] + gi, a);
                    else acc[e] = a;
                }
                else { red_add(s_out[b] + gi, v); }
            }
        }
    }
}

} // namespace kron
