// kernel_dmma.cuh -- n = 8, d = 4, double precision on the FP64 tensor pipe ("dmma" path).
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78) for BASELINE config 4 (n = 8, d = 4).
//
// tcgen05 has no FP64 kind, so the FP64 tensor path on sm_100a is mma.sync m8n8k4 (SASS DMMA.8x8x4).
// One mode product with an 8x8 factor M over an 8x8 slice T (contracted index u, other index v) is
//     D[i][v] = sum_u M[i][u] T[u][v]        = two m8n8k4 steps over u.
// Fragment layout (lane = 4g+q):  A: M[g][k=q]   B: T[k=q][col=g]   C/D: D[g][2q], D[g][2q+1].
// Using the contraction order u = 2q+s in step s (A_s = M[g][2q+s], B_s = T[2q+s][g]), a lane's two B
// values are ADJACENT elements of T (one 128-bit load), and the C/D fragment of this product,
// D[i=g][v=2q+s], is exactly the B fragment needed to contract v next.  Two factors are therefore
// chained in registers with no data movement, and the result lands on the same (u in {2q,2q+1},
// v = g) positions the lane loaded from.  Consequences:
//   phase 1 (indices i3, i2): each warp loads its 8x8 slices straight from global memory as one
//           coalesced 512-byte request, chains 4 DMMAs, writes 128-bit to shared memory;
//   phase 2 (indices i1, i0): slices are strided in memory; a 128-bit shared load fetches the same
//           (u, v) element of two neighbouring slices, so the warp works on slice pairs; a 16-byte-
//           chunk XOR swizzle keyed on (i0, i1) makes both phases bank-conflict free;
//   the factor matrices never touch shared memory (2 doubles per lane per factor), and the last
//           DMMA accumulates directly onto the running sum of the current run of equal output
//           pointers (C operand), flushed with RED per element when the pointer changes.
// Shared-memory traffic is 2 x 32 KiB per item (the reference moves 8 x 32 KiB through GLOBAL memory
// for d = 4, kronmult.cu:112-121).  Summation order inside a dot product: u even/odd interleaved by
// the tensor pipe instead of k ascending -- within the 1e-12 relative-L2 tolerance of BASELINE.json.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh" // prefetch_l2
#include <atomic>

namespace kron
{

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// volatile twin: keeps the issue order written in the source (the chains below are interleaved on purpose; left to
// itself the compiler re-serialises them chain by chain to save registers)
__device__ __forceinline__ void dmma884v(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// Two chained mode products (factor U on the contracted index of the B fragment, then factor V on the other one) on NS
// independent 8 x 8 slices at once, stage by stage, so that ptxas can interleave the NS dependency chains: one chain is
// four back-to-back DEPENDENT DMMAs, and a warp that issues them slice after slice (load, 4 DMMAs, store, next slice)
// keeps the FP64 tensor pipe idle for most of each DMMA's latency -- ncu on the persistent n = 8, d = 6 kernel showed the
// pipe 38 % busy with exactly that instruction order (profiles/ncu_dmma_l2_r02.md).
template<int NS>
__device__ __forceinline__ void dmma_pair(double u0, double u1, double v0, double v1, const double (&x0)[NS],
                                          const double (&x1)[NS], double (&z0)[NS], double (&z1)[NS])
{
    double y0[NS], y1[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(y0[i], y1[i], u0, x0[i], 0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(y0[i], y1[i], u1, x1[i], y0[i], y1[i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(z0[i], z1[i], v0, y0[i], 0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(z0[i], z1[i], v1, y1[i], z0[i], z1[i]);
}
// the same with the second product accumulating onto (c0, c1) -- the running sum of a run of equal output pointers
template<int NS>
__device__ __forceinline__ void dmma_pair_acc(double u0, double u1, double v0, double v1, const double (&x0)[NS],
                                              const double (&x1)[NS], double (&c0)[NS], double (&c1)[NS])
{
    double y0[NS], y1[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(y0[i], y1[i], u0, x0[i], 0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(y0[i], y1[i], u1, x1[i], y0[i], y1[i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(c0[i], c1[i], v0, y0[i], c0[i], c1[i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) dmma884v(c0[i], c1[i], v1, y1[i], c0[i], c1[i]);
}

// 16-byte-chunk swizzle of the exchange buffer: slice h = i0*8+i1 holds 64 contiguous doubles
// (32 chunks); chunk index is XORed with ((i0&1)<<2 | i1>>1).
__device__ __forceinline__ int dmma_sigma(int h) { return (((h >> 3) & 1) << 2) | ((h >> 1) & 3); }

// Phase 2 of the n = 8 kernels on a chunk-swizzled 4096-element tile: warp w owns the P2 slice pairs j = w*P2 .. and
// contracts the two slow indices of the tile (rows h0 = 8g + 2q, h0 + 1), GJ slice pairs (2 GJ chains) at a time.
// acc[jj][s + 2t] = Out[i0 = g][i1 = 2q + s][f = 2j + t]: the products accumulate onto the run sums.  RS = row pitch of the
// buffer in elements (64: a whole 64-column tile; 32: the compact half tiles of kernel_dmma_l2.cuh).
template<int P2, int GJ, int RS = 64>
__device__ __forceinline__ void dmma_phase2_acc(const double *__restrict__ Ec, int w, int g, int q, double u0, double u1,
                                                double v0, double v1, double (&acc)[P2][4], int jbase = 0)
{
    const int h0 = g * 8 + 2 * q;
    const int sg = ((g & 1) << 2) | q; // dmma_sigma(h0) == dmma_sigma(h0 + 1)
#pragma unroll
    for (int j0 = 0; j0 < P2; j0 += GJ)
    {
        double x0[2 * GJ], x1[2 * GJ], c0[2 * GJ], c1[2 * GJ];
#pragma unroll
        for (int i = 0; i < GJ; ++i)
        {
            const int j = jbase + w * P2 + j0 + i; // slices f = 2j, 2j+1
            const double2 a0 = *reinterpret_cast<const double2 *>(Ec + h0 * RS + ((j ^ sg) << 1));
            const double2 a1 = *reinterpret_cast<const double2 *>(Ec + (h0 + 1) * RS + ((j ^ sg) << 1));
            x0[2 * i] = a0.x; x1[2 * i] = a1.x; x0[2 * i + 1] = a0.y; x1[2 * i + 1] = a1.y;
            c0[2 * i] = acc[j0 + i][0]; c1[2 * i] = acc[j0 + i][1]; c0[2 * i + 1] = acc[j0 + i][2]; c1[2 * i + 1] = acc[j0 + i][3];
        }
        dmma_pair_acc<2 * GJ>(u0, u1, v0, v1, x0, x1, c0, c1);
#pragma unroll
        for (int i = 0; i < GJ; ++i)
        {
            acc[j0 + i][0] = c0[2 * i]; acc[j0 + i][1] = c1[2 * i]; acc[j0 + i][2] = c0[2 * i + 1]; acc[j0 + i][3] = c1[2 * i + 1];
        }
    }
}
// the same in place (pass A of the multi-pass routes: the tile goes back to memory afterwards)
template<int P2, int GJ>
__device__ __forceinline__ void dmma_phase2_inplace(double *__restrict__ Ec, int w, int g, int q, double u0, double u1,
                                                    double v0, double v1)
{
    const int h0 = g * 8 + 2 * q;
    const int sg = ((g & 1) << 2) | q;
#pragma unroll
    for (int j0 = 0; j0 < P2; j0 += GJ)
    {
        double x0[2 * GJ], x1[2 * GJ], z0[2 * GJ], z1[2 * GJ];
#pragma unroll
        for (int i = 0; i < GJ; ++i)
        {
            const int j = w * P2 + j0 + i;
            const double2 a0 = *reinterpret_cast<const double2 *>(Ec + h0 * 64 + ((j ^ sg) << 1));
            const double2 a1 = *reinterpret_cast<const double2 *>(Ec + (h0 + 1) * 64 + ((j ^ sg) << 1));
            x0[2 * i] = a0.x; x1[2 * i] = a1.x; x0[2 * i + 1] = a0.y; x1[2 * i + 1] = a1.y;
        }
        dmma_pair<2 * GJ>(u0, u1, v0, v1, x0, x1, z0, z1);
        // a lane rewrites exactly the two chunks it read: no other lane touches them in this phase
#pragma unroll
        for (int i = 0; i < GJ; ++i)
        {
            const int j = w * P2 + j0 + i;
            *reinterpret_cast<double2 *>(Ec + h0 * 64 + ((j ^ sg) << 1))       = make_double2(z0[2 * i], z0[2 * i + 1]);
            *reinterpret_cast<double2 *>(Ec + (h0 + 1) * 64 + ((j ^ sg) << 1)) = make_double2(z1[2 * i], z1[2 * i + 1]);
        }
    }
}
// Phase 1 on a tile in shared memory: warp w owns the T1 slices h = w*T1 .., contracts their two fast indices G1 slices
// at a time and rewrites them chunk-swizzled in place.  `gsrc` != nullptr: the tile is read from global memory instead
// (vectors that are not 16-byte aligned cannot come by TMA).
template<int T1, int G1>
__device__ __forceinline__ void dmma_phase1_inplace(double *__restrict__ Ec, const double *__restrict__ gsrc, int w, int g,
                                                    int q, double u0, double u1, double v0, double v1)
{
#pragma unroll
    for (int t0 = 0; t0 < T1; t0 += G1)
    {
        double x0[G1], x1[G1], z0[G1], z1[G1];
#pragma unroll
        for (int i = 0; i < G1; ++i)
        {
            const int h = w * T1 + t0 + i;
            if (!gsrc)
            {
                const double2 v = *reinterpret_cast<const double2 *>(Ec + h * 64 + g * 8 + 2 * q);
                x0[i] = v.x; x1[i] = v.y;
            }
            else { x0[i] = __ldg(gsrc + h * 64 + g * 8 + 2 * q); x1[i] = __ldg(gsrc + h * 64 + g * 8 + 2 * q + 1); }
        }
        dmma_pair<G1>(u0, u1, v0, v1, x0, x1, z0, z1);
        __syncwarp(); // every lane has read its chunks of these slices before they are rewritten
#pragma unroll
        for (int i = 0; i < G1; ++i)
        {
            const int h = w * T1 + t0 + i;
            const int chunk16 = (4 * g + q) ^ dmma_sigma(h);
            *reinterpret_cast<double2 *>(Ec + h * 64 + chunk16 * 2) = make_double2(z0[i], z1[i]);
        }
    }
}

struct Dmma84
{
    static constexpr int N       = 4096;
    static constexpr int WARPS   = 4;               // warps cooperating on one item
    static constexpr int THREADS = WARPS * 32;
    static constexpr int T1      = 64 / WARPS;      // phase-1 slices per warp
    static constexpr int P2      = 32 / WARPS;      // phase-2 slice pairs per warp
    static constexpr int SMEM    = 2 * N * 8 + 32;  // two item slots + their mbarriers
    static constexpr int G1      = 8;               // phase-1 slices a warp works on at once (independent DMMA chains)
    static constexpr int G2      = 4;               // phase-2 slice pairs at once (two chains each)
};

// Items arrive by TMA: one elected thread fetches item k+1 with a single 32 KiB cp.async.bulk into the other slot
// right after the barrier of item k (every warp has then left item k-1, the slot's last user), and pulls item k+3
// into L2 with one cp.async.bulk.prefetch.L2.  Phase 1 reads its 512-byte slices from the slot and writes them
// back swizzled IN PLACE (a slice is read and rewritten by the same warp), phase 2 reads slice pairs, the flush
// reuses the slot.  One CTA barrier per item (two when a run ends).  ncu on the previous version, which loaded
// phase 1 straight from global memory, showed 7-12 long-scoreboard stall cycles per issued instruction.
__global__ void __launch_bounds__(Dmma84::THREADS, 3)
kron_dmma84_kernel(const double *const *__restrict__ A, double *const *__restrict__ in, double *const *__restrict__ out,
                   const int lda, const int nb, const long long items_per_cta)
{
    using C = Dmma84;
    constexpr int N = C::N, T1 = C::T1, P2 = C::P2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *E     = reinterpret_cast<double *>(smem_raw); // [2][4096]
    uint64_t *bar = reinterpret_cast<uint64_t *>(E + 2 * N);

    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int g = lane >> 2, q = lane & 3;

    const long long k0 = (long long)blockIdx.x * items_per_cta;
    long long kend     = k0 + items_per_cta;
    if (kend > nb) kend = nb;
    if (k0 >= kend) return;

    if (t == 0)
    {
        mbar_init(bar + 0, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[P2][4];
#pragma unroll
    for (int j = 0; j < P2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;

    // item k -> slot (k - k0) & 1, thread 0 only; vectors that are not 16-byte aligned are read from global memory
    auto issue = [&](long long k, const double *ip) {
        uint64_t *b = bar + ((k - k0) & 1);
        if (aligned16(ip))
        {
            fence_proxy_async(); // the slot was last WRITTEN through the generic proxy (in-place phase 1, flush)
            mbar_arrive_expect_tx(b, N * 8);
            tma_load_1d(E + ((k - k0) & 1) * N, ip, N * 8, b);
        }
        else mbar_arrive(b);
    };
    auto l2_pull = [&](const double *ip) {
        if (aligned16(ip)) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ip), "r"(N * 8) : "memory");
    };

    // software pipeline over items: factor fragments one item ahead, their pointers two ahead
    const int lane_off0 = g + (2 * q) * lda; // M[g][2q]; M[g][2q+1] is lda further
    const double *ap[4];
    double a_nxt[8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const double *p0 = A[k0 * 4 + j];
        a_nxt[2 * j]     = __ldg(p0 + lane_off0);
        a_nxt[2 * j + 1] = __ldg(p0 + lane_off0 + lda);
        ap[j]            = (k0 + 1 < kend) ? A[(k0 + 1) * 4 + j] : nullptr;
    }
    const double *ip_cur = in[k0];
    const double *ip_nxt = (k0 + 1 < kend) ? in[k0 + 1] : nullptr;
    if (t == 0)
    {
        issue(k0, ip_cur);
        if (ip_nxt) l2_pull(ip_nxt);
        if (k0 + 2 < kend) l2_pull(in[k0 + 2]);
    }
    double *o_cur = out[k0];

    for (long long k = k0; k < kend; ++k)
    {
        const int cur = (int)((k - k0) & 1);
        double *Ec    = E + cur * N;
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a_nxt[i];
        const double *ip_n2 = nullptr;
        if (k + 1 < kend)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                a_nxt[2 * j]     = __ldg(ap[j] + lane_off0);
                a_nxt[2 * j + 1] = __ldg(ap[j] + lane_off0 + lda);
            }
            if (k + 2 < kend)
            {
#pragma unroll
                for (int j = 0; j < 4; ++j) ap[j] = A[(k + 2) * 4 + j];
                ip_n2 = in[k + 2];
            }
        }

        // ---------------- phase 1: factors 3 (index i3 = u) and 2 (index i2 = v), slice by slice, in place
        const bool vec = aligned16(ip_cur);
        mbar_wait(bar + cur, (unsigned)((k - k0) >> 1) & 1u);
        dmma_phase1_inplace<T1, C::G1>(Ec, vec ? nullptr : ip_cur, w, g, q, a[6], a[7], a[4], a[5]);
        __syncthreads();
        // every warp has left item k-1 (phase 2 and flush read the other slot): refill it
        if (t == 0)
        {
            if (ip_nxt) issue(k + 1, ip_nxt);
            if (k + 3 < kend) l2_pull(in[k + 3]);
        }

        // ---------------- phase 2: factors 1 (index i1 = u) and 0 (index i0 = v), slice pairs
        dmma_phase2_acc<P2, C::G2>(Ec, w, g, q, a[2], a[3], a[0], a[1], acc);

        double *o_next = (k + 1 < kend) ? out[k + 1] : nullptr;
        if (o_next != o_cur) // uniform over the CTA
        {
            // A lane's sums sit at Out[i0 = g][i1 = 2q + s][f = 2j + t] (acc[jj][s + 2t]): addresses
            // 512 B apart across lanes.  128-byte-strided REDs are ~7x slower than coalesced ones
            // (profiles/microbench_r01.jsonl), so transpose through the slot first: each warp overwrites
            // exactly the elements it alone read in phase 2 (no barrier needed before), then the CTA
            // reads the item linearly and issues sector-complete REDs.
#pragma unroll
            for (int jj = 0; jj < P2; ++jj)
            {
                const int j  = w * P2 + jj;
                const int h0 = g * 8 + 2 * q;
                const int sg = ((g & 1) << 2) | q;
                *reinterpret_cast<double2 *>(Ec + h0 * 64 + ((j ^ sg) << 1))       = make_double2(acc[jj][0], acc[jj][2]);
                *reinterpret_cast<double2 *>(Ec + (h0 + 1) * 64 + ((j ^ sg) << 1)) = make_double2(acc[jj][1], acc[jj][3]);
                acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.0;
            }
            __syncthreads();
#pragma unroll 4
            for (int i = 0; i < N / 2 / C::THREADS; ++i)
            {
                const int c  = t + i * C::THREADS; // 16-byte chunk of the item, linear order
                const int h  = c >> 5;
                const double2 v = *reinterpret_cast<const double2 *>(Ec + h * 64 + (((c & 31) ^ dmma_sigma(h)) << 1));
                red_add(o_cur + 2 * c, v.x);
                red_add(o_cur + 2 * c + 1, v.y);
            }
        }
        o_cur  = o_next;
        ip_cur = ip_nxt;
        ip_nxt = ip_n2;
    }
}

static cudaError_t launch_dmma84(int sms, const double *const *A, int lda, double *const *in, double *const *out,
                                 int nb, cudaStream_t st, std::atomic<long long> &launches)
{
    using C = Dmma84;
    cudaError_t e = kernel_setup(kron_dmma84_kernel, C::SMEM); // per device (common.cuh)
    if (e != cudaSuccess) return e;
    // one contiguous range of items per CTA (runs of equal output pointers stay together), one wave of CTAs
    long long grid = (long long)sms * 3;
    long long ipc  = ((long long)nb + grid - 1) / grid;
    if (ipc > 64) ipc = (ipc + 31) / 32 * 32;
    grid = ((long long)nb + ipc - 1) / ipc;
    kron_dmma84_kernel<<<(int)grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb, ipc);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// n = 8, d = 5 and 6 (n^d = 32768 / 262144: the reference's `large`/`realistic` cases,
// tests/kronmult_bench_gpu.cpp:71-72).  The vector no longer fits on chip, so the factors are applied in two
// passes through global memory, in place in `input` (which kronmult.cuh:23 allows to be clobbered):
//   pass A (kron_dmma8_tile4_kernel): the four fastest factors on every contiguous 4096-element tile,
//          with exactly the two-phase DMMA scheme above, result written back over the tile;
//   pass B (kron_dmma8_rows2_kernel, d = 6): factors 0 and 1 on 64 x 64 tiles (64 rows = (i0, i1), 64
//          consecutive columns), fetched with 16-byte cp.async into the chunk-swizzled layout, DMMA phase 2,
//          accumulated over consecutive items of equal output pointer and flushed with coalesced REDG.
//          For d = 5 the remaining single factor goes through the generic pass kernel.
// The reference moves 2 x N x 8 bytes through global memory per FACTOR (kronmult.cu:112-121); this moves
// 3 x N x 8 bytes per ITEM.
// ------------------------------------------------------------------------------------------------
// Pass A is a pure streaming pass (read 32 KiB, 512 DMMAs, write 32 KiB per tile), so it is built as a
// 3-stage TMA ring: one elected thread fetches the tile two units ahead with a single cp.async.bulk; phase 1
// reads its 512-byte slices from the ring slot and writes them back swizzled IN PLACE (a slice is read and
// written by one warp instruction pair), phase 2 works in place too, and the write-back streams the slot
// linearly to global memory.  Factor fragments are fetched one unit ahead, their pointers two.
__global__ void __launch_bounds__(Dmma84::THREADS, 2)
kron_dmma8_tile4_kernel(const double *const *__restrict__ A, double *const *__restrict__ in,
                        double *const *__restrict__ dst, const int lda, const int d, const int tiles_per_item,
                        const long long total_units)
{
    using C = Dmma84;
    constexpr int N = C::N, T1 = C::T1, P2 = C::P2, NST = 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *R     = reinterpret_cast<double *>(smem_raw);            // [3][4096] ring of tiles
    uint64_t *bar = reinterpret_cast<uint64_t *>(R + NST * N);      // [3] "tile landed"
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int g = lane >> 2, q = lane & 3;
    const int lane_off0 = g + (2 * q) * lda;

    if (t == 0)
    {
        for (int i = 0; i < NST; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto unit_of  = [&](int it) { return (long long)blockIdx.x + (long long)it * gridDim.x; };
    auto tile_ptr = [&](long long u) -> double * {
        const long long k = u / tiles_per_item;
        return in[k] + (u - k * tiles_per_item) * N;
    };
    // fetch unit `it` into ring slot it % 3 (thread 0 only).  Unaligned tiles are loaded by everybody in fill().
    auto issue = [&](int it) {
        const long long u = unit_of(it);
        if (u >= total_units) return;
        double *src = tile_ptr(u);
        if (aligned16(src))
        {
            fence_proxy_async();
            mbar_arrive_expect_tx(bar + it % NST, N * 8);
            tma_load_1d(R + (it % NST) * N, src, N * 8, bar + it % NST);
        }
        else mbar_arrive(bar + it % NST);
    };
    auto load_frags = [&](long long u, double (&a)[8]) {
        const long long k = u / tiles_per_item;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const double *p0 = A[k * d + (d - 4) + j];
            a[2 * j]     = __ldg(p0 + lane_off0);
            a[2 * j + 1] = __ldg(p0 + lane_off0 + lda);
        }
    };

    if (t == 0) { issue(0); issue(1); }
    double a_nxt[8];
    if (unit_of(0) < total_units) load_frags(unit_of(0), a_nxt);
    unsigned parity = 0; // bit i = phase parity of ring slot i

    for (int it = 0; unit_of(it) < total_units; ++it)
    {
        const long long u = unit_of(it);
        const int slot    = it % NST;
        double *Ec        = R + slot * N;
        double *base      = tile_ptr(u);
        const bool vec    = aligned16(base);
        // results go back over the tile (dst == in: the reference's clobbering contract) or into the per-item
        // scratch vector of the read-only-input entry points
        double *wb        = dst[u / tiles_per_item] + (u % tiles_per_item) * N;
        const bool vecw   = aligned16(wb);
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a_nxt[i];
        if (unit_of(it + 1) < total_units) load_frags(unit_of(it + 1), a_nxt);

        mbar_wait(bar + slot, (parity >> slot) & 1u);
        parity ^= 1u << slot;
        // phase 1: the two fastest indices, slice by slice, in place
        dmma_phase1_inplace<T1, C::G1>(Ec, vec ? nullptr : base, w, g, q, a[6], a[7], a[4], a[5]);
        __syncthreads();
        // everyone has left the previous unit (its write-back read slot (it+2)%3): refill that slot
        if (t == 0) issue(it + 2);
        // phase 2: the next two indices, slice pairs, in place
        dmma_phase2_inplace<P2, C::G2>(Ec, w, g, q, a[2], a[3], a[0], a[1]);
        __syncthreads();
        // linear write-back of the tile (coalesced 128-bit stores)
#pragma unroll 4
        for (int i = 0; i < N / 2 / C::THREADS; ++i)
        {
            const int c  = t + i * C::THREADS;
            const int h  = c >> 5;
            const double2 v = *reinterpret_cast<const double2 *>(Ec + h * 64 + (((c & 31) ^ dmma_sigma(h)) << 1));
            if (vecw) *reinterpret_cast<double2 *>(wb + 2 * c) = v;
            else { wb[2 * c] = v.x; wb[2 * c + 1] = v.y; }
        }
    }
}

// pass B for d = 6: factors 0 and 1.  L = n^d / 64 columns per row, tile = 64 rows x 64 columns.
__global__ void __launch_bounds__(Dmma84::THREADS, 3)
kron_dmma8_rows2_kernel(const double *const *__restrict__ A, double *const *__restrict__ in, double *const *__restrict__ out,
                        const int lda, const int nb, const int d, const long long L, const int tiles, const int chunk,
                        const long long total_units)
{
    using C = Dmma84;
    constexpr int N = C::N, P2 = C::P2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *E = reinterpret_cast<double *>(smem_raw); // [2][4096]
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int g = lane >> 2, q = lane & 3;
    const int lane_off0 = g + (2 * q) * lda;

    double acc[P2][4];
#pragma unroll
    for (int j = 0; j < P2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;

    // tile of item k -> swizzled exchange buffer, 16-byte chunks (8-byte elements when unaligned)
    auto fetch = [&](long long k, int tile, double *Eb) {
        const double *src = in[k] + (long long)tile * 64;
        const bool vec    = aligned16(src) && ((L & 1) == 0);
#pragma unroll 4
        for (int i = 0; i < N / 2 / C::THREADS; ++i)
        {
            const int c = t + i * C::THREADS, h = c >> 5, ci = c & 31;
            double *dst = Eb + h * 64 + ((ci ^ dmma_sigma(h)) << 1);
            const double *s2 = src + (long long)h * L + 2 * ci;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            if (vec) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(s2) : "memory");
            else
            {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(s2) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + 8), "l"(s2 + 1) : "memory");
            }
        }
        cp_async_commit();
    };

    for (long long u = blockIdx.x; u < total_units; u += gridDim.x)
    {
        const long long c0   = u / tiles;            // item chunk
        const int tile       = (int)(u - c0 * tiles);
        const long long k0   = c0 * chunk;
        const long long kend = (k0 + chunk < nb) ? k0 + chunk : nb;
        __syncthreads(); // previous unit's readers are done with both buffers
        fetch(k0, tile, E);
        double *o_cur = out[k0];
        for (long long k = k0; k < kend; ++k)
        {
            double *Ec = E + (int)((k - k0) & 1) * N;
            double a[4];
            {
                const double *p0 = A[k * d + 0], *p1 = A[k * d + 1];
                a[0] = __ldg(p0 + lane_off0); a[1] = __ldg(p0 + lane_off0 + lda);
                a[2] = __ldg(p1 + lane_off0); a[3] = __ldg(p1 + lane_off0 + lda);
            }
            double *o_next = (k + 1 < kend) ? out[k + 1] : nullptr;
            cp_async_wait_all();
            __syncthreads(); // tile k is visible; everyone left iteration k-1, so the other buffer is free
            if (k + 1 < kend) fetch(k + 1, tile, E + (int)(((k - k0) & 1) ^ 1) * N);
            dmma_phase2_acc<P2, C::G2>(Ec, w, g, q, a[2], a[3], a[0], a[1], acc);
            if (o_next != o_cur) // uniform over the CTA
            {
#pragma unroll
                for (int jj = 0; jj < P2; ++jj)
                {
                    const int j  = w * P2 + jj;
                    const int h0 = g * 8 + 2 * q;
                    const int sg = ((g & 1) << 2) | q;
                    *reinterpret_cast<double2 *>(Ec + h0 * 64 + ((j ^ sg) << 1))       = make_double2(acc[jj][0], acc[jj][2]);
                    *reinterpret_cast<double2 *>(Ec + (h0 + 1) * 64 + ((j ^ sg) << 1)) = make_double2(acc[jj][1], acc[jj][3]);
                    acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.0;
                }
                __syncthreads();
                double *obase = o_cur + (long long)tile * 64;
#pragma unroll 4
                for (int i = 0; i < N / 2 / C::THREADS; ++i)
                {
                    const int c = t + i * C::THREADS, h = c >> 5, ci = c & 31;
                    const double2 v = *reinterpret_cast<const double2 *>(Ec + h * 64 + ((ci ^ dmma_sigma(h)) << 1));
                    red_add(obase + (long long)h * L + 2 * ci, v.x);
                    red_add(obase + (long long)h * L + 2 * ci + 1, v.y);
                }
            }
            o_cur = o_next;
        }
    }
}

// pass A for n = 8, d >= 5: the four fastest factors, in place
static constexpr int TILE4_SMEM = 3 * Dmma84::N * 8 + 64;
static cudaError_t launch_dmma8_tile4(int sms, int d, long long N, const double *const *A, int lda, double *const *in,
                                      double *const *dst,
                                      int nb, cudaStream_t st, std::atomic<long long> &launches)
{
    using C = Dmma84;
    cudaError_t e = kernel_setup(kron_dmma8_tile4_kernel, TILE4_SMEM);
    if (e != cudaSuccess) return e;
    const int tpi            = (int)(N / C::N);
    const long long units    = (long long)nb * tpi;
    const long long max_grid = (long long)sms * 2;
    const int grid           = (int)(units < max_grid ? units : max_grid);
    kron_dmma8_tile4_kernel<<<grid, C::THREADS, TILE4_SMEM, st>>>(A, in, dst, lda, d, tpi, units);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// pass B for n = 8, d = 6: factors 0 and 1 with accumulation into the outputs
static cudaError_t launch_dmma8_rows2(int sms, int d, long long N, const double *const *A, int lda, double *const *in,
                                      double *const *out, int nb, cudaStream_t st, std::atomic<long long> &launches)
{
    using C = Dmma84;
    cudaError_t e = kernel_setup(kron_dmma8_rows2_kernel, C::SMEM);
    if (e != cudaSuccess) return e;
    const long long L = N / 64;
    const int tiles   = (int)(L / 64);
    long long chunk   = ((long long)nb * tiles) / ((long long)sms * 12);
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    const long long units    = ((nb + chunk - 1) / chunk) * tiles;
    const long long max_grid = (long long)sms * 3;
    const int grid           = (int)(units < max_grid ? units : max_grid);
    kron_dmma8_rows2_kernel<<<grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb, d, L, tiles, (int)chunk, units);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

} // namespace kron
#include "kernel_dmma_l2.cuh" // n = 8, d = 6: both passes in one persistent kernel, intermediate resident in L2
namespace kron
{

// ------------------------------------------------------------------------------------------------
// n = 5 .. 8, d = 2 and 3, both precisions: one WARP per item, everything in registers ("dmma", warp per item).
// ncu on the pair-tile kernel these shapes used (profiles/ncu_pairtile_small_r02.md): n = 8, d = 3 is shared-memory-
// bound with a third of its wavefronts being bank-conflict replays, the d = 2 items are instruction-bound.  With
// the chained DMMA of the d = 4 kernel a warp needs no shared memory for the two fastest factors at all: a lane
// loads its two adjacent elements (row g, columns 2q, 2q+1) of every n x n slice straight from global memory (one
// coalesced request per slice), 4 DMMAs contract the slice's two indices and leave the result on the positions it
// was loaded from.  n < 8 runs on the same 8 x 8 x 4 tiles with the slices and factors ZERO-PADDED in registers (the
// lanes outside the n x n corner load nothing and add nothing): the tensor pipe has time to spare on these HBM-bound
// shapes.  Single precision is converted to double at the loads and back at the final adds (more accurate than the
// fp32 reference, within its 1e-5 tolerance a fortiori) -- there is no fp32 tensor path of sufficient precision.
// d = 3: the remaining factor combines the n slices with coefficients that are uniform over the warp -- its n^2
// entries travel through 512 bytes of shared memory per warp (cp.async one item ahead, broadcast 128-bit loads).
// Runs of equal output pointers are summed in the 2 (2n) result registers; the flush is two REDG per lane and slice on
// adjacent elements.  Data of item s+1 is in flight in registers while item s is computed; pointers two items ahead.
inline std::atomic<int> &dmma8s_enabled() { static std::atomic<int> v{1}; return v; } // knob 11

template<typename T, int NN, int D>
__global__ void __launch_bounds__(128, (D == 2) ? 6 : 3)
kron_dmma8s_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                   const int lda, const int nb, const long long items_per_warp)
{
    constexpr int SL  = (D == 2) ? 1 : NN;    // n x n slices per item
    constexpr int NSQ = NN * NN;
    __shared__ __align__(16) double F0s[4][2][64]; // d = 3: factor 0 of items s, s+1 per warp (column-major, pitch 8, fp64)
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int g = lane >> 2, q = lane & 3;
    // my two adjacent elements of a slice: row g, columns 2q and 2q+1 (outside the n x n corner: zero padding)
    const bool in0 = (g < NN) && (2 * q < NN), in1 = (g < NN) && (2 * q + 1 < NN);
    const int epos = g * NN + 2 * q;
    const long long lane_off0 = g + (long long)(2 * q) * lda; // factor fragment: M[g][2q], M[g][2q+1]

    const long long k0 = ((long long)blockIdx.x * 4 + w) * items_per_warp;
    if (k0 >= nb) return;
    const int cnt = (int)((k0 + items_per_warp <= nb) ? items_per_warp : (nb - k0));

    struct Ptrs { const T *ip; T *op; const T *ap[D]; };
    auto load_ptrs = [&](int s, Ptrs &p) {
        if (s >= cnt) { p.ip = nullptr; p.op = nullptr; return; }
        const long long k = k0 + s;
        p.ip = in[k]; p.op = out[k];
#pragma unroll
        for (int j = 0; j < D; ++j) p.ap[j] = A[k * D + j];
    };
    struct Data { double x[SL][2]; double a[4]; };
    auto load_data = [&](const Ptrs &p, Data &dt, int slot) {
        if (p.ip)
        {
            if constexpr (NN == 8 && sizeof(T) == 8)
            {
                if (aligned16(p.ip))
                {
#pragma unroll
                    for (int h = 0; h < SL; ++h)
                    {
                        const double2 v = __ldg(reinterpret_cast<const double2 *>(p.ip + h * NSQ + epos));
                        dt.x[h][0] = v.x; dt.x[h][1] = v.y;
                    }
                }
                else
                {
#pragma unroll
                    for (int h = 0; h < SL; ++h) { dt.x[h][0] = __ldg(p.ip + h * NSQ + epos); dt.x[h][1] = __ldg(p.ip + h * NSQ + epos + 1); }
                }
            }
            else
            {
#pragma unroll
                for (int h = 0; h < SL; ++h)
                {
                    dt.x[h][0] = in0 ? (double)__ldg(p.ip + h * NSQ + epos) : 0.0;
                    dt.x[h][1] = in1 ? (double)__ldg(p.ip + h * NSQ + epos + 1) : 0.0;
                }
            }
            dt.a[0] = in0 ? (double)__ldg(p.ap[D - 2] + lane_off0) : 0.0; dt.a[1] = in1 ? (double)__ldg(p.ap[D - 2] + lane_off0 + lda) : 0.0;
            dt.a[2] = in0 ? (double)__ldg(p.ap[D - 1] + lane_off0) : 0.0; dt.a[3] = in1 ? (double)__ldg(p.ap[D - 1] + lane_off0 + lda) : 0.0;
            if constexpr (D == 3)
            {
                // factor 0 -> shared memory as doubles, column-major with pitch 8: element (r, c) at c*8 + r
                if constexpr (sizeof(T) == 8)
                {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(&F0s[w][slot][0]);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                    {
                        const int e = lane + 32 * i, r = e % NN, c = e / NN;
                        if (e < NSQ)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + (c * 8 + r) * 8), "l"(p.ap[0] + r + (long long)c * lda) : "memory");
                    }
                }
                else
                {
                    // fp32 factors are converted on the way: plain loads + shared stores (the values are needed a full item later)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                    {
                        const int e = lane + 32 * i, r = e % NN, c = e / NN;
                        if (e < NSQ) F0s[w][slot][c * 8 + r] = (double)__ldg(p.ap[0] + r + (long long)c * lda);
                    }
                }
            }
        }
        // always a group, even an empty one past the last item: the consumer's `wait_group 1` counts groups, and without
        // it the last item's factor 0 would still be allowed in flight when it is read (found by racecheck, round 2)
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    Ptrs p_cur, p_nxt, p_nx2;
    Data d_cur, d_nxt;
    load_ptrs(0, p_cur);
    load_ptrs(1, p_nxt);
    load_data(p_cur, d_cur, 0);

    double acc[SL][2];
#pragma unroll
    for (int h = 0; h < SL; ++h) acc[h][0] = acc[h][1] = 0.0;

    for (int s = 0; s < cnt; ++s)
    {
        load_ptrs(s + 2, p_nx2);
        load_data(p_nxt, d_nxt, (s + 1) & 1); // item s+1 in flight while item s is computed
        if constexpr (D == 3)
        {
            asm volatile("cp.async.wait_group 1;" ::: "memory"); // factor 0 of item s has landed (item s+1's may be pending)
            __syncwarp();
        }
        double z[SL][2];
#pragma unroll
        for (int h = 0; h < SL; ++h)
        {
            double y0, y1;
            dmma884(y0, y1, d_cur.a[2], d_cur.x[h][0], 0.0, 0.0);
            dmma884(y0, y1, d_cur.a[3], d_cur.x[h][1], y0, y1);
            if constexpr (D == 2)
            {
                // the second product accumulates straight onto the run sum (C operand)
                dmma884(acc[0][0], acc[0][1], d_cur.a[0], y0, acc[0][0], acc[0][1]);
                dmma884(acc[0][0], acc[0][1], d_cur.a[1], y1, acc[0][0], acc[0][1]);
            }
            else
            {
                dmma884(z[h][0], z[h][1], d_cur.a[0], y0, 0.0, 0.0);
                dmma884(z[h][0], z[h][1], d_cur.a[1], y1, z[h][0], z[h][1]);
            }
        }
        if constexpr (D == 3)
        {
            const double *F = &F0s[w][s & 1][0];
#pragma unroll
            for (int h = 0; h < NN; ++h)
            {
                double f[8]; // column h of factor 0: F0(h', h), h' = 0..n-1 (pitch 8; the padding is never used)
#pragma unroll
                for (int c = 0; c < (NN + 1) / 2; ++c)
                {
                    const double2 v = *reinterpret_cast<const double2 *>(F + h * 8 + 2 * c);
                    f[2 * c] = v.x; f[2 * c + 1] = v.y;
                }
#pragma unroll
                for (int hp = 0; hp < NN; ++hp)
                {
                    acc[hp][0] = fma(f[hp], z[h][0], acc[hp][0]);
                    acc[hp][1] = fma(f[hp], z[h][1], acc[hp][1]);
                }
            }
            __syncwarp(); // every lane has read slot s & 1 before item s+2's factor is copied into it
        }
        if (p_nxt.op != p_cur.op) // end of the run of equal output pointers (also the last item: p_nxt.op is null)
        {
#pragma unroll
            for (int h = 0; h < SL; ++h)
            {
                if (in0) red_add(p_cur.op + h * NSQ + epos, (T)acc[h][0]);
                if (in1) red_add(p_cur.op + h * NSQ + epos + 1, (T)acc[h][1]);
                acc[h][0] = acc[h][1] = 0.0;
            }
        }
        p_cur = p_nxt; p_nxt = p_nx2;
        d_cur = d_nxt;
    }
}

template<typename T, int NN, int D>
static cudaError_t launch_dmma8s(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                 cudaStream_t st, std::atomic<long long> &launches)
{
    const long long warps = (long long)sms * ((D == 2) ? 6 : 3) * 4;
    long long ipw = ((long long)nb + warps - 1) / warps;
    if (ipw > 64) ipw = (ipw + 31) / 32 * 32;
    const long long nw   = ((long long)nb + ipw - 1) / ipw;
    const long long grid = (nw + 3) / 4;
    kron_dmma8s_kernel<T, NN, D><<<(int)grid, 128, 0, st>>>(A, in, out, lda, nb, ipw);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// which (T, n, d) the warp-per-item kernel takes: measured per shape against the kernels it replaces
// (tools/dmmaw_session.py, profiles/dmma_warp_per_item_r02.jsonl; fraction of the roofline, new vs old):
//   fp64  d = 2: n = 5 0.53 / 0.64, n = 6 0.64 / 0.69, n = 7 0.75 / 0.54, n = 8 0.87 / 0.43
//         d = 3: n = 5 0.49 / 0.37, n = 6 0.57 / 0.41, n = 7 0.65 / 0.32, n = 8 0.68 / 0.44
//   fp32 (computed in double: the FP64 pipe is the bound) loses everywhere but n = 8, d = 3 (0.39 / 0.35)
template<typename T>
static bool dmma8s_takes(int n, int d)
{
    const int mode = dmma8s_enabled().load(std::memory_order_relaxed);
    if (mode == 0 || n < 5 || n > 8 || (d != 2 && d != 3)) return false;
    if (mode == 2) return true; // everything it can do (development)
    if (sizeof(T) == 8) return n == 8 || d == 3 || n == 7;
    return n == 8 && d == 3;
}

template<typename T>
static cudaError_t run_dmma8s(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                              cudaStream_t st, std::atomic<long long> &launches)
{
#define KRON_D8S(NN, DD) if (n == NN && d == DD) return launch_dmma8s<T, NN, DD>(sms, A, lda, in, out, nb, st, launches);
    KRON_D8S(5, 2) KRON_D8S(6, 2) KRON_D8S(7, 2) KRON_D8S(8, 2) KRON_D8S(5, 3) KRON_D8S(6, 3) KRON_D8S(7, 3) KRON_D8S(8, 3)
#undef KRON_D8S
    return cudaErrorNotSupported;
}

// cudaErrorNotSupported when the shape or type is outside the family.
// d = 5 returns cudaSuccess after pass A with *remaining = 1: the caller applies factor 0 (generic pass).
template<typename T>
static cudaError_t run_dmma(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                            cudaStream_t st, std::atomic<long long> &launches, const char *&last_path, int *remaining,
                            T *const *scratch = nullptr)
{
    *remaining = 0;
    if (dmma8s_takes<T>(n, d))
    {
        last_path = "dmma";
        return run_dmma8s<T>(sms, d, n, A, lda, in, out, nb, st, launches);
    }
    if constexpr (sizeof(T) == 8)
    {
        if (n == 8 && d == 4)
        {
            last_path = "dmma";
            return launch_dmma84(sms, A, lda, in, out, nb, st, launches);
        }

        if (n == 8 && (d == 5 || d == 6))
        {
            const long long N = (d == 5) ? 32768 : 262144;
            last_path = "dmma-multipass";
            // scratch != nullptr: `in` is read-only, pass A writes into the scratch vectors and the rest works there
            T *const *work = scratch ? scratch : in;
            if (dmma86_l2_mode().load(std::memory_order_relaxed) > 0)
            {
                // one persistent kernel; the intermediate lives in a library-owned ring in L2, `in` is only read
                last_path = "dmma-l2";
                return d == 6 ? launch_dmma8_l2<6>(sms, A, lda, in, out, nb, st, launches)
                              : launch_dmma8_l2<5>(sms, A, lda, in, out, nb, st, launches);
            }
            if (d == 6)
            {
                // both passes chunk by chunk, so that pass B reads pass A's result from L2 (common.cuh)
                const long long cb = multipass_chunk_items(nb, N * 8, 2, 8);
                ChunkStreams cs;
                cudaError_t e = cs.begin(st, (nb + cb - 1) / cb);
                if (e != cudaSuccess) return e;
                for (long long k0 = 0; k0 < nb; k0 += cb)
                {
                    const int cnt     = (int)(nb - k0 < cb ? nb - k0 : cb);
                    cudaStream_t sc   = cs.pick();
                    e = launch_dmma8_tile4(sms, d, N, A + k0 * d, lda, in + k0, work + k0, cnt, sc, launches);
                    if (e == cudaSuccess) e = launch_dmma8_rows2(sms, d, N, A + k0 * d, lda, work + k0, out + k0, cnt, sc, launches);
                    if (e == cudaSuccess) e = launch_discard<T>(work + k0, cnt, N, sc);
                    if (e != cudaSuccess) break;
                }
                const cudaError_t j = cs.end();
                return e != cudaSuccess ? e : j;
            }
            cudaError_t e = launch_dmma8_tile4(sms, d, N, A, lda, in, work, nb, st, launches);
            if (e != cudaSuccess) return e;
            *remaining = 1;
            return cudaSuccess;
        }
    }
    return cudaErrorNotSupported;
}

} // namespace kron
