// kernel_regtile.cuh -- register-tiled in-place mode products for n = 4 ("regtile" path), d in {4,5,6}.
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78) for BASELINE configs 3 and 5 (n = 4, d = 6 / 5).
//
// Design (B200-first, not a port):
//  * A thread owns a 4x4 sub-tensor (two tensor indices, 16 values) in registers and applies two
//    factors to it before touching shared memory again, so a d = 6 item needs only two exchanges
//    through shared memory (4 x N x sizeof(T) bytes of shared-memory traffic) where a one-factor-
//    per-pass scheme needs 2d (the reference does those passes through GLOBAL memory,
//    kronmult.cu:112-121).
//      phase A: slowest index/indices, operands loaded straight from global memory (coalesced:
//               consecutive threads hold consecutive fast indices), result -> shared memory
//      phase B: the two fastest indices, row-wise (128-bit) in place in shared memory
//      phase C: the middle index/indices, column-wise; the result is added to 16 per-thread
//               accumulators whose global addresses are 128-byte contiguous per half-warp, so the
//               flush is made of sector-complete REDs (128-byte-strided REDs measured 7x slower)
//  * Items that share an output pointer and are consecutive in the batch (ASGarD-style runs) are
//    summed in those registers; global memory sees one RED per element per run instead of one
//    atomicAdd per element per item (kronmult.cu:126-129).  The flush is always atomic-class, so any
//    aliasing pattern remains correct.
//  * The next item is fetched while the current one is processed: one elected thread issues a 1-D TMA
//    bulk copy (cp.async.bulk, completion on an mbarrier) of the whole vector into the other item
//    buffer right after the single CTA barrier of the step; whole items are additionally pulled
//    into L2 three steps ahead (prefetch.global.L2).  No registers are held by data in flight.
//    Vectors that are not 16-byte aligned (the API only promises alignment to T) take an
//    element-wise cp.async (LDGSTS) route into the same slots.  Factor matrices use cp.async too.
//  * TMA writes the vector linearly; phase A reads it column-wise (conflict free), and writes its
//    result in the 128-byte XOR-swizzled layout IN PLACE: the swizzle only permutes slots among
//    the lanes of one warp, so a __syncwarp() between the warp's loads and stores suffices.  The
//    swizzle makes both the row-wise 128-bit accesses of phase B and the column-wise accesses of
//    phase C bank-conflict free.
//  * After phase A the item decomposes into independent 256-element slices (fixed slow indices),
//    each owned by one aligned 16-thread group in both phase B and phase C: the B -> C exchange
//    needs only __syncwarp().  One CTA barrier per item remains (after phase A).
//  * The FP64 pipe is the roofline for d = 6: 24 DFMA per element, ~0.4 other instructions per DFMA.
// Per output element the products are still accumulated k ascending from 0 as in
// multiply_transpose (kronmult.cu:66-70); only the order in which the d factors are applied
// (slowest index first here, fastest first in the reference) and the summation over a run differ,
// which is covered by the 1e-12 / 1e-5 relative-L2 tolerance of BASELINE.json.
#pragma once
#include "common.cuh"
#include <atomic>

namespace kron
{

template<typename T>
__device__ __forceinline__ int swz(int idx)
{
    // 128-byte XOR swizzle on an element index: 16-byte chunk ^= (128-byte line & 7)
    if constexpr (sizeof(T) == 8) return idx ^ (((idx >> 4) & 7) << 1);
    else return idx ^ (((idx >> 5) & 7) << 2);
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// 16 contiguous, 16-byte aligned values from shared memory
template<typename T>
__device__ __forceinline__ void lds16(const T *p, T (&m)[16])
{
    if constexpr (sizeof(T) == 8)
    {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const double2 w = q[i];
            m[2 * i] = w.x; m[2 * i + 1] = w.y;
        }
    }
    else
    {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const float4 w = q[i];
            m[4 * i] = w.x; m[4 * i + 1] = w.y; m[4 * i + 2] = w.z; m[4 * i + 3] = w.w;
        }
    }
}

// x is a 4x4 register tile indexed hi*4+lo.  STRIDE = 1: apply M along lo; STRIDE = 4: along hi.
// Ms: row-major 4x4 factor in shared memory (Ms[i*4+k] = M(i,k)).
template<typename T, int STRIDE>
__device__ __forceinline__ void tile16_apply(T (&x)[16], const T *Ms)
{
    T m[16];
    lds16<T>(Ms, m);
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        const int base = (STRIDE == 1) ? f * 4 : f;
        const T a0 = x[base], a1 = x[base + STRIDE], a2 = x[base + 2 * STRIDE], a3 = x[base + 3 * STRIDE];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            T dot = a0 * m[i * 4];
            dot += a1 * m[i * 4 + 1];
            dot += a2 * m[i * 4 + 2];
            dot += a3 * m[i * 4 + 3];
            x[base + i * STRIDE] = dot;
        }
    }
}

template<typename T, int D>
struct Regtile4
{
    static constexpr int N       = ipow(4, D);
    static constexpr int TPI     = N / 16;                       // threads per item
    static constexpr int THREADS = (D == 6) ? 256 : 64;          // CTA size
    static constexpr int MINB    = (D == 6) ? 2 : 8;             // resident CTAs per SM aimed at
    static constexpr int B       = THREADS / TPI;                // items side by side in one CTA
    static constexpr int AD      = (D == 6) ? 2 : 1;             // factors applied in phase A
    static constexpr int CD      = D - AD - 2;                   // factors applied in phase C
    static constexpr int AT      = ipow(4, AD);                  // phase-A fibre bundle: 4 or 16 values
    static constexpr int FA      = 16 / AT;                      // bundles per thread in phase A
    static constexpr int R       = N / AT;                       // element stride of the phase-A index
    static constexpr int MSTR    = D * 16 + 16 / (int)sizeof(T); // per-item factor block, padded by 16 B
    static constexpr int LD      = (D * 16 + TPI - 1) / TPI;     // factor elements staged per thread
    static constexpr int LPS     = B * N * (int)sizeof(T) / 128; // 128-byte lines per step
    static constexpr int NMB     = 3;                            // factor buffers (steps s, s+1, s+2)
    static constexpr int SMEM    = (2 * B * N + NMB * B * MSTR) * (int)sizeof(T) + 16;
    static_assert(D >= 4 && D <= 6, "regtile covers d = 4..6");
    static_assert(LPS <= THREADS, "one L2 prefetch line per thread");
};

template<typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// 1-D TMA: global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// like tile16_apply, but the result of the factor is added onto acc instead of replacing x
template<typename T, int STRIDE>
__device__ __forceinline__ void tile16_apply_acc(const T (&x)[16], const T *Ms, T (&acc)[16])
{
    T m[16];
    lds16<T>(Ms, m);
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        const int base = (STRIDE == 1) ? f * 4 : f;
        const T a0 = x[base], a1 = x[base + STRIDE], a2 = x[base + 2 * STRIDE], a3 = x[base + 3 * STRIDE];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            T dot = acc[base + i * STRIDE];
            dot += a0 * m[i * 4];
            dot += a1 * m[i * 4 + 1];
            dot += a2 * m[i * 4 + 2];
            dot += a3 * m[i * 4 + 3];
            acc[base + i * STRIDE] = dot;
        }
    }
}

// STAGE = 0: operands of the next step arrive by TMA (or element-wise cp.async) in the other item buffer.
// STAGE = 1: they are prefetched into L1 (prefetch.global.L1) and phase A loads them from global.
template<typename T, int D, int STAGE>
__global__ void __launch_bounds__(Regtile4<T, D>::THREADS, Regtile4<T, D>::MINB)
kron_regtile4_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                     const int lda, const int nb, const int chunk, const long long ngroups)
{
    using C = Regtile4<T, D>;
    constexpr int N = C::N, TPI = C::TPI, B = C::B, AD = C::AD, AT = C::AT, FA = C::FA, R = C::R;
    constexpr int MSTR = C::MSTR, LD = C::LD, NMB = C::NMB;
    constexpr int PFD = 3; // L2 prefetch distance in items
    constexpr unsigned ITEM_BYTES = N * sizeof(T);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *S          = reinterpret_cast<T *>(smem_raw);       // [2][B*N]     item buffers
    T *MS         = S + 2 * B * N;                         // [3][B*MSTR]  factor matrices, row-major 4x4 blocks
    uint64_t *bar = reinterpret_cast<uint64_t *>(MS + NMB * B * MSTR); // [2] one mbarrier per item buffer

    const int t  = threadIdx.x;
    const int b  = t / TPI; // stream (item slot) of this thread
    const int tl = t % TPI;

    if (t == 0)
    {
        mbar_init(bar + 0, B);
        mbar_init(bar + 1, B);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned parity0 = 0, parity1 = 0;

    // factor elements this thread stages for ITS OWN item each step (global col-major -> row-major)
    int l_ok[LD], l_dst[LD], l_src[LD], l_j[LD];
#pragma unroll
    for (int q = 0; q < LD; ++q)
    {
        const int e  = tl + q * TPI;
        const int ej = e / 16;
        const int ec = (e % 16) / 4; // column
        const int er = e % 4;        // row (contiguous in memory)
        l_ok[q]  = e < D * 16;
        l_j[q]   = ej;
        l_dst[q] = b * MSTR + ej * 16 + er * 4 + ec;
        l_src[q] = er + ec * lda;
    }
    // L2-prefetch role: one 128-byte line of the items PFD steps ahead
    constexpr int EPL = 128 / (int)sizeof(T);
    const int p_b  = (t * EPL) / N;
    const int p_lo = (t * EPL) % N;

    T acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = T(0);

    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x)
    {
        const long long g0   = g * (long long)(B * chunk);
        const long long k0   = g0 + (long long)b * chunk;
        const long long kend = (k0 + chunk < nb) ? k0 + chunk : nb;

        // does step s exist for this CTA?  (uniform: stream 0 holds the smallest item index)
        auto step_exists = [&](int s) { return s < chunk && g0 + s < nb; };
        // asynchronous fetch of the operands of step s (vector pointer ip, may be null) into item buffer s&1
        auto stage_data = [&](int s, const T *ip) {
            if (!step_exists(s)) return;
            const bool v = ip != nullptr;
            if constexpr (STAGE == 1)
            {
                if (v)
                {
#pragma unroll
                    for (int i = 0; i < (N * (int)sizeof(T) / 128 + TPI - 1) / TPI; ++i)
                    {
                        const int line = tl + i * TPI;
                        if (line < N * (int)sizeof(T) / 128)
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(ip + line * (128 / (int)sizeof(T))));
                    }
                }
                return;
            }
            const bool tma = v && aligned16(ip);
            T *dst         = S + (s & 1) * (B * N) + b * N;
            if (tl == 0)
            {
                if (tma)
                {
                    fence_proxy_async(); // generic-proxy accesses of this buffer (ordered by the barrier) first
                    mbar_arrive_expect_tx(bar + (s & 1), ITEM_BYTES);
                    tma_load_1d(dst, ip, ITEM_BYTES, bar + (s & 1));
                }
                else mbar_arrive(bar + (s & 1));
            }
            if (v && !tma)
            {
#pragma unroll
                for (int q = 0; q < FA; ++q)
#pragma unroll
                    for (int h = 0; h < AT; ++h) cp_async_elem<T>(dst + h * R + tl + q * TPI, ip + h * R + tl + q * TPI);
            }
        };
        // factor elements of step s: ap[q] = pointer to the factor this thread copies from (may be null)
        auto stage_mats = [&](int s, const T *const (&ap)[LD]) {
            if (!step_exists(s)) return;
#pragma unroll
            for (int q = 0; q < LD; ++q)
                if (l_ok[q] && ap[q]) cp_async_elem<T>(MS + (s % NMB) * (B * MSTR) + l_dst[q], ap[q] + l_src[q]);
        };
        auto load_in_ptr = [&](int s) -> const T * {
            const long long k = k0 + s;
            return (s < chunk && k < kend) ? in[k] : nullptr;
        };
        auto load_mat_ptrs = [&](int s, const T *(&ap)[LD]) {
            const long long k = k0 + s;
            const bool v      = s < chunk && k < kend;
#pragma unroll
            for (int q = 0; q < LD; ++q) ap[q] = (v && l_ok[q]) ? A[k * D + l_j[q]] : nullptr;
        };
        auto l2_prefetch = [&](int s) {
            if (t < C::LPS && s < chunk)
            {
                const long long kk = g0 + (long long)p_b * chunk + s;
                if (kk < nb && kk < g0 + (long long)(p_b + 1) * chunk) prefetch_l2(in[kk] + p_lo);
            }
        };

        __syncthreads(); // the previous group's readers are done with S and MS; mbarriers initialised
#pragma unroll
        for (int a = 0; a < PFD; ++a) l2_prefetch(a);
        const T *ap_n[LD];
        const T *ip_cur = load_in_ptr(0);
        stage_data(0, ip_cur);
        load_mat_ptrs(0, ap_n);
        stage_mats(0, ap_n);
        load_mat_ptrs(1, ap_n);
        stage_mats(1, ap_n);
        cp_async_commit();
        T *o_cur = (k0 < kend) ? out[k0] : nullptr;
        cp_async_wait_all();
        __syncthreads(); // factor matrices of steps 0 and 1 are visible to every thread

        for (int s = 0; s < chunk; ++s)
        {
            if (g0 + s >= nb) break; // uniform
            const int cur     = s & 1;
            const long long k = k0 + s;
            const bool valid  = k < kend;
            T *Sc             = S + cur * (B * N);
            const T *Mc       = MS + (s % NMB) * (B * MSTR) + b * MSTR;
            l2_prefetch(s + PFD);
            // pointers needed after the barrier are fetched now, so their latency hides behind phase A
            const T *ip_n = load_in_ptr(s + 1);
            load_mat_ptrs(s + 2, ap_n);
            T *o_next = (valid && k + 1 < kend) ? out[k + 1] : nullptr;

            // ---------------- phase A: slowest index/indices; linear in, swizzled out, in place per warp
            cp_async_wait_all();                 // element-wise copies issued by this thread during step s-1
            T x[16];
            if constexpr (STAGE == 0)
            {
                if (cur == 0) { mbar_wait(bar + 0, parity0); parity0 ^= 1; }
                else          { mbar_wait(bar + 1, parity1); parity1 ^= 1; }
#pragma unroll
                for (int q = 0; q < FA; ++q)
#pragma unroll
                    for (int h = 0; h < AT; ++h) x[q * AT + h] = Sc[b * N + h * R + tl + q * TPI];
            }
            else
            {
#pragma unroll
                for (int q = 0; q < FA; ++q)
#pragma unroll
                    for (int h = 0; h < AT; ++h) x[q * AT + h] = valid ? ip_cur[h * R + tl + q * TPI] : T(0);
            }
            if constexpr (AD == 2)
            {
                tile16_apply<T, 1>(x, Mc + 1 * 16); // factor 1 acts on the low two bits of h
                tile16_apply<T, 4>(x, Mc + 0 * 16); // factor 0 on the high two bits
            }
            else { tile16_apply<T, 1>(x, Mc); } // four independent fibres of factor 0
            __syncwarp(); // every lane has read its linear slots: the swizzle permutes slots within a warp only
#pragma unroll
            for (int q = 0; q < FA; ++q)
#pragma unroll
                for (int h = 0; h < AT; ++h) Sc[swz<T>(b * N + h * R + tl + q * TPI)] = x[q * AT + h];
            // The item's threads must all have written before rows are read.  Past this barrier every
            // thread has also left phase C of step s-1: the other item buffer and factor buffer
            // (s+2)%3 are free, and the factor copies awaited above are visible to the whole item.
            if constexpr (TPI <= 16) __syncwarp();
            else __syncthreads();

            stage_data(s + 1, ip_n);
            stage_mats(s + 2, ap_n);
            cp_async_commit();
            ip_cur = ip_n;

            // ---------------- phase B: two fastest indices, row-wise 128-bit, in place
            {
                constexpr int VE = 16 / (int)sizeof(T);
                const int e0 = b * N + tl * 16;
#pragma unroll
                for (int c = 0; c < 16 / VE; ++c)
                {
                    const T *src = Sc + swz<T>(e0 + c * VE);
                    if constexpr (sizeof(T) == 8)
                    {
                        const double2 w = *reinterpret_cast<const double2 *>(src);
                        x[2 * c] = w.x; x[2 * c + 1] = w.y;
                    }
                    else
                    {
                        const float4 w = *reinterpret_cast<const float4 *>(src);
                        x[4 * c] = w.x; x[4 * c + 1] = w.y; x[4 * c + 2] = w.z; x[4 * c + 3] = w.w;
                    }
                }
                tile16_apply<T, 1>(x, Mc + (D - 1) * 16);
                tile16_apply<T, 4>(x, Mc + (D - 2) * 16);
#pragma unroll
                for (int c = 0; c < 16 / VE; ++c)
                {
                    T *dst = Sc + swz<T>(e0 + c * VE);
                    if constexpr (sizeof(T) == 8) *reinterpret_cast<double2 *>(dst) = make_double2(x[2 * c], x[2 * c + 1]);
                    else *reinterpret_cast<float4 *>(dst) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
                }
            }
            __syncwarp(); // rows h*16 .. h*16+15 are produced and consumed by the same aligned 16 threads

            // ---------------- phase C: middle index/indices, column-wise (lanes along the fast index)
            if (valid) // finished streams keep their accumulators clean (stale factor slots may hold NaNs)
            {
                const int eC = b * N + (tl / 16) * 256 + (tl % 16);
#pragma unroll
                for (int m = 0; m < 16; ++m) x[m] = Sc[swz<T>(eC + m * 16)];
                if constexpr (C::CD == 2)
                {
                    tile16_apply<T, 1>(x, Mc + (D - 3) * 16);
                    tile16_apply_acc<T, 4>(x, Mc + (D - 4) * 16, acc);
                }
                else { tile16_apply_acc<T, 1>(x, Mc + (D - 3) * 16, acc); }
            }
            if (valid)
            {
                if (o_next != o_cur)
                {
                    T *dst = o_cur + (tl / 16) * 256 + (tl % 16);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                    {
                        red_add(dst + i * 16, acc[i]);
                        acc[i] = T(0);
                    }
                }
                o_cur = o_next;
            }
        }
    }
}

template<typename T, int D, int STAGE>
static cudaError_t launch_regtile4(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                   cudaStream_t st, std::atomic<long long> &launches)
{
    using C  = Regtile4<T, D>;
    auto kfn = kron_regtile4_kernel<T, D, STAGE>;
    cudaError_t e = kernel_setup(kfn, C::SMEM); // per device (common.cuh)
    if (e != cudaSuccess) return e;
    // consecutive items per stream: long runs of equal outputs merge in registers
    long long chunk = nb / ((long long)C::B * sms * C::MINB * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    const long long per_group = (long long)C::B * chunk;
    const long long ngroups   = (nb + per_group - 1) / per_group;
    const long long max_grid  = (long long)sms * C::MINB;
    const int grid            = (int)(ngroups < max_grid ? ngroups : max_grid);
    kfn<<<grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb, (int)chunk, ngroups);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// experiment knob (kronmult_b200_set_tuning): operand staging mode of the regtile kernels
static std::atomic<int> g_regtile_stage{0};

// cudaErrorNotSupported when (n, d) is outside the family
template<typename T>
static cudaError_t run_regtile(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out,
                               int nb, cudaStream_t st, std::atomic<long long> &launches, const char *&last_path)
{
    if (n != 4 || d < 4 || d > 6) return cudaErrorNotSupported;
    cudaError_t e;
    const int stage = g_regtile_stage.load(std::memory_order_relaxed);
#define KRON_RT(DD)                                                                              \
    (stage == 1 ? launch_regtile4<T, DD, 1>(sms, A, lda, in, out, nb, st, launches)              \
                : launch_regtile4<T, DD, 0>(sms, A, lda, in, out, nb, st, launches))
    if (d == 4) e = KRON_RT(4);
    else if (d == 5) e = KRON_RT(5);
    else e = KRON_RT(6);
#undef KRON_RT
    last_path = "regtile";
    return e;
}

} // namespace kron
