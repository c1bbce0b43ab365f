// kernel_pairpass.cuh -- the multi-pass route for vectors that do not fit in shared memory, built from the register
// tiles of kernel_pairtile.cuh ("pairtile-multipass").
//
// Replaces, for n^d beyond shared memory (the top of the reference's sweep envelope,
// tests/kronmult_fullbench_gpu.cpp:70-74: n = 6, 7 with d = 6; n = 9, 10 with d = 5, 6), the reference's
// cuda_kronmult (kronmult_gpu/kronmult.cu:95-130), which streams 2 n^d elements through global memory per FACTOR.
// Here the d factors are applied in 2 (3) passes over the vector, each pass a kernel that applies a group of GF
// factors on tiles of n^(GF+Q) elements:
//   pass 1   the GF fastest factors on contiguous tiles (Q = 0), written back in place (or into the scratch vector
//            of the read-only-input entry points);
//   pass 2.. the next GF factors on tiles of n^GF rows x n^Q contiguous columns (rows lie L = n^(factors done) apart);
//            the last pass adds into `output`, summing runs of consecutive items with equal output pointers in shared
//            memory first, and flushes with REDG.
// A tile is treated exactly like a resident item of kernel_pairtile.cuh whose Q fastest indices carry no factor:
// two factors per shared-memory round trip on n x n register tiles, the next tile arrives by cp.async while the
// current one is worked on.  Units are ordered tile-major, item-minor, so that consecutive units of a CTA are the
// same tile of consecutive items (-> runs of equal outputs).
#pragma once
#include "kernel_pairtile.cuh"

namespace kron
{

template<typename T_, int n_, int GF_, int Q_, bool FINAL_>
struct PassCfg : PairBase<T_, n_, GF_ + Q_, GF_>
{
    using Base = PairBase<T_, n_, GF_ + Q_, GF_>;
    using Base::S; using Base::ITEMP; using Base::MATP; using Base::TPS; using Base::REGS;
    static constexpr bool TO_OUT   = FINAL_;
    static constexpr bool REGFLUSH = true;
    static constexpr int PSTRIDE   = GF_ + 3; // GF factors, source tile, output tile, destination tile
    static constexpr int POUT      = GF_ + 1;
    static constexpr int B         = 1;
    static constexpr int STAGES    = 2;
    static constexpr int ACC       = FINAL_ ? 1 : 0;
    static constexpr int TILES     = TPS;
    static constexpr int LANEMAP   = 0;
    static constexpr int cap()
    {
        int c = (65536 / REGS) / 32 * 32;
        return c > 256 ? 256 : (c < 32 ? 32 : c);
    }
    static constexpr int ITERS   = (TILES + cap() - 1) / cap();
    static constexpr int THREADS = ((TILES + ITERS - 1) / ITERS + 31) / 32 * 32;

    static constexpr int OFF_ACC  = STAGES * ITEMP * S;
    static constexpr int OFF_MAT  = OFF_ACC + ACC * ITEMP * S;
    static constexpr int OFF_PTR  = OFF_MAT + STAGES * MATP * S;
    static constexpr int OFF_FLAG = OFF_PTR + 3 * PSTRIDE * 8;
    static constexpr int SMEM     = OFF_FLAG + 16;
    static constexpr bool FITS    = SMEM <= 226 * 1024;
    static constexpr int minb()
    {
        int m = 227 * 1024 / (SMEM + 1024);
        const int r = 65536 / (THREADS * REGS);
        if (r < m) m = r;
        return m > 8 ? 8 : (m < 1 ? 1 : m);
    }
    static constexpr int MINB = minb();
};

struct PassArgs
{
    int lda, nb, d, j0;            // j0: index of the slowest factor of this pass's group
    long long L;                   // distance between the rows of a tile = n^(factors already applied)
    long long tiles_per_item;
    long long lblocks;             // column blocks per row group = L / n^Q
    long long units;               // tiles_per_item * nb
};

template<typename T, int n, int GF, int Q, bool FINAL>
__global__ void __launch_bounds__(PassCfg<T, n, GF, Q, FINAL>::THREADS, PassCfg<T, n, GF, Q, FINAL>::MINB)
    kron_pairpass_kernel(const T *const *__restrict__ A, T *const *__restrict__ src, T *const *__restrict__ dst,
                         T *const *__restrict__ out, const PassArgs p)
{
    using C = PassCfg<T, n, GF, Q, FINAL>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;

    const long long per = (p.units + gridDim.x - 1) / gridDim.x;
    const long long u0  = per * blockIdx.x;
    long long u1        = u0 + per;
    if (u1 > p.units) u1 = p.units;
    if (u0 >= u1) return;
    const int steps = (int)(u1 - u0);

    auto slot_ptrs = [&](int slot) { return reinterpret_cast<T **>(smem + C::OFF_PTR) + (size_t)slot * C::PSTRIDE; };
    auto slot_flag = [&](int slot) { return reinterpret_cast<int *>(smem + C::OFF_FLAG) + slot; };

    // unit -> (tile, item); pointers of the step: GF factors, source tile, output tile, destination tile
    auto fetch_ptrs = [&](int t) {
        if (tid >= C::PSTRIDE) return;
        const long long u    = u0 + t;
        const long long tile = u / p.nb;
        const long long k    = u - tile * p.nb;
        const long long h    = tile / p.lblocks;
        const long long lb   = tile - h * p.lblocks;
        const long long off  = h * (long long)ipow(n, GF) * p.L + lb * C::LBQ;
        T **sp               = slot_ptrs(t % 3);
        T *ptr               = nullptr;
        if (tid < GF) ptr = const_cast<T *>(A[k * p.d + p.j0 + tid]);
        else if (tid == GF) ptr = src[k] + off;
        else if (tid == GF + 2) ptr = FINAL ? nullptr : dst[k] + off;
        else
        {
            int flag = 1 | 2 | 4;
            if constexpr (FINAL)
            {
                T *o             = out[k];
                ptr              = o + off;
                const bool first = (t == 0) || (k == 0) || (out[k - 1] != o);
                const bool last  = (t + 1 == steps) || (k + 1 == p.nb) || (out[k + 1] != o);
                flag             = 1 | (first ? 2 : 0) | (last ? 4 : 0);
            }
            *slot_flag(t % 3) = flag;
        }
        sp[tid] = ptr;
    };

    auto issue_copies = [&](int t, int stage) {
        T *const *sp = slot_ptrs(t % 3);
        T *vec       = reinterpret_cast<T *>(smem) + (size_t)stage * C::ITEMP;
        T *mats      = reinterpret_cast<T *>(smem + C::OFF_MAT) + (size_t)stage * C::MATP;
        const T *s0  = sp[GF];
        if constexpr (Q == 0)
        {
            // one contiguous run of N elements
            if (aligned16(s0))
            {
                for (int q = tid; q < C::NCH; q += C::THREADS)
                {
                    int el = q * C::VEC;
                    if constexpr (C::PADE > 0) el += (q / (C::NSQ / C::VEC)) * C::PADE;
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(vec + el);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(s0 + q * C::VEC) : "memory");
                }
                for (int i = C::NCH * C::VEC + tid; i < C::N; i += C::THREADS) cp_async_elem<T>(vec + i, s0 + i);
            }
            else
            {
                for (int i = tid; i < C::N; i += C::THREADS)
                {
                    int el = i;
                    if constexpr (C::PADE > 0) el += (i / C::NSQ) * C::PADE;
                    cp_async_elem<T>(vec + el, s0 + i);
                }
            }
        }
        else
        {
            // n^GF rows of LBQ contiguous elements, L apart
            constexpr int LB = C::LBQ;
            const bool vecok = (LB % C::VEC == 0) && (p.L % C::VEC == 0) && aligned16(s0);
            if (vecok)
            {
                constexpr int CPR = LB / C::VEC > 0 ? LB / C::VEC : 1; // chunks per row
                for (int q = tid; q < C::N / C::VEC; q += C::THREADS)
                {
                    const int m = q / CPR;
                    const int c = q - m * CPR;
                    int el      = q * C::VEC;
                    if constexpr (C::PADE > 0) el += (el / C::NSQ) * C::PADE;
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(vec + el);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(s0 + m * p.L + c * C::VEC) : "memory");
                }
            }
            else
            {
                for (int i = tid; i < C::N; i += C::THREADS)
                {
                    const int m = i / LB;
                    const int l = i - m * LB;
                    int el      = i;
                    if constexpr (C::PADE > 0) el += (i / C::NSQ) * C::PADE;
                    cp_async_elem<T>(vec + el, s0 + m * p.L + l);
                }
            }
        }
        // the GF factors, column-major with pitch RP
        constexpr int CPC = (n % C::VEC == 0) ? n / C::VEC : 1;
        if ((n % C::VEC == 0) && (p.lda % C::VEC == 0))
        {
            for (int r = tid; r < GF * n * CPC; r += C::THREADS)
            {
                const int j  = r / (n * CPC);
                const int rc = r - j * (n * CPC);
                const int cc = rc / CPC;
                const int q  = rc - cc * CPC;
                const T *g   = sp[j] + (long long)cc * p.lda + q * C::VEC;
                T *s         = mats + j * n * C::RP + cc * C::RP + q * C::VEC;
                if (aligned16(g))
                {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < C::VEC; ++i) cp_async_elem<T>(s + i, g + i);
                }
            }
        }
        else
        {
            for (int r = tid; r < GF * C::NSQ; r += C::THREADS)
            {
                const int j  = r / C::NSQ;
                const int rc = r - j * C::NSQ;
                const int cc = rc / n;
                const int rr = rc - cc * n;
                cp_async_elem<T>(mats + j * n * C::RP + cc * C::RP + rr, sp[j] + rr + (long long)cc * p.lda);
            }
        }
        cp_async_commit();
    };

    fetch_ptrs(0);
    __syncthreads();
    issue_copies(0, 0);
    if (steps > 1) fetch_ptrs(1);

    for (int t = 0; t < steps; ++t)
    {
        const int stage = t & 1;
        cp_async_wait_all();
        __syncthreads(); // step t's tile and step t+1's pointers are visible; the other stage is free
        if (t + 1 < steps) issue_copies(t + 1, stage ^ 1);
        if (t + 2 < steps) fetch_ptrs(t + 2);

        pair_passes<C, 0>(smem, stage, slot_flag(t % 3), slot_ptrs(t % 3), p.L);

        if constexpr (!FINAL)
        {
            // the tile goes back to global memory row by row
            __syncthreads();
            const T *vec = reinterpret_cast<const T *>(smem) + (size_t)stage * C::ITEMP;
            T *g0        = slot_ptrs(t % 3)[GF + 2];
            constexpr int LB = C::LBQ;
            const bool vecok = (Q == 0 || ((LB % C::VEC == 0) && (p.L % C::VEC == 0))) && aligned16(g0) && C::TAIL == 0;
            if (vecok)
            {
                constexpr int CPR = (Q == 0) ? C::N / C::VEC : (LB / C::VEC > 0 ? LB / C::VEC : 1);
                for (int q = tid; q < C::N / C::VEC; q += C::THREADS)
                {
                    const int m = q / CPR;
                    const int c = q - m * CPR;
                    int el      = q * C::VEC;
                    if constexpr (C::PADE > 0) el += (el / C::NSQ) * C::PADE;
                    const int4 w = *reinterpret_cast<const int4 *>(vec + el);
                    T *g         = (Q == 0) ? g0 + q * C::VEC : g0 + m * p.L + c * C::VEC;
                    *reinterpret_cast<int4 *>(g) = w;
                }
            }
            else
            {
                for (int i = tid; i < C::N; i += C::THREADS)
                {
                    const int m = i / LB;
                    const int l = i - m * LB;
                    int el      = i;
                    if constexpr (C::PADE > 0) el += (i / C::NSQ) * C::PADE;
                    if constexpr (Q == 0) g0[i] = vec[el];
                    else g0[m * p.L + l] = vec[el];
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
template<typename T, int n, int GF, int Q, bool FINAL>
static cudaError_t launch_pairpass(int sms, const T *const *A, T *const *src, T *const *dst, T *const *out,
                                   const PassArgs &pa, cudaStream_t st, std::atomic<long long> &launches)
{
    using C = PassCfg<T, n, GF, Q, FINAL>;
    if constexpr (!C::FITS) { return cudaErrorNotSupported; }
    else
    {
        auto kfn = kron_pairpass_kernel<T, n, GF, Q, FINAL>;
        {
            cudaError_t e = kernel_setup(kfn, C::SMEM); // per device, cached (common.cuh)
            if (e != cudaSuccess) return e;
        }
        // contiguous unit ranges per CTA; at least ~8 units per CTA so that the pipeline has something to overlap
        long long grid = (long long)sms * C::MINB;
        if (grid > pa.units) grid = pa.units;
        kfn<<<(int)grid, C::THREADS, C::SMEM, st>>>(A, src, dst, out, pa);
        launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
}

// largest factor group whose tile n^g stays within TILE_MAX bytes
template<typename T>
constexpr int pairpass_gmax(int n)
{
    int g = 1;
    long long e = n;
    while (e * n * (long long)sizeof(T) <= 68 * 1024) { e *= n; ++g; }
    return g;
}

// cudaErrorNotSupported when (T, n, d) is outside the family; defined in pairpass_f64.cu / pairpass_f32.cu.
// scratch != nullptr: `in` is read-only, the first pass writes the scratch vectors and the rest works there.
template<typename T>
cudaError_t run_pairpass(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                         cudaStream_t st, std::atomic<long long> &launches, T *const *scratch);

template<typename T, int n>
static cudaError_t run_pairpass_chunk(int sms, int d, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                      cudaStream_t st, std::atomic<long long> &launches, T *const *scratch, long long N);

template<typename T, int n>
static cudaError_t run_pairpass_n(int sms, int d, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                  cudaStream_t st, std::atomic<long long> &launches, T *const *scratch)
{
    constexpr int GM = pairpass_gmax<T>(n);
    static_assert(GM >= 3, "tiles of at least three indices");
    if (d <= GM) return cudaErrorNotSupported;
    if (GM > 3 && d - GM > 2) return cudaErrorNotSupported; // non-final later passes are only built for GM = 3
    long long N = 1;
    for (int i = 0; i < d; ++i) N *= n;
    // chunks of items whose vectors stay in L2 between the passes (common.cuh); all passes of a chunk, then the next
    const int passes   = 1 + (d - GM + 1) / 2;
    const long long cb = multipass_chunk_items(nb, N * (long long)sizeof(T), passes);
    if (cb < nb)
    {
        ChunkStreams cs;
        cudaError_t e = cs.begin(st, (nb + cb - 1) / cb);
        if (e != cudaSuccess) return e;
        for (long long k0 = 0; k0 < nb; k0 += cb)
        {
            const int cnt = (int)(nb - k0 < cb ? nb - k0 : cb);
            e = run_pairpass_chunk<T, n>(sms, d, A + k0 * d, lda, in + k0, out + k0, cnt, cs.pick(), launches,
                                         scratch ? scratch + k0 : nullptr, N);
            if (e != cudaSuccess) break;
        }
        const cudaError_t j = cs.end();
        return e != cudaSuccess ? e : j;
    }
    return run_pairpass_chunk<T, n>(sms, d, A, lda, in, out, nb, st, launches, scratch, N);
}

template<typename T, int n>
static cudaError_t run_pairpass_chunk(int sms, int d, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                      cudaStream_t st, std::atomic<long long> &launches, T *const *scratch, long long N)
{
    constexpr int GM = pairpass_gmax<T>(n);
    T *const *work = scratch ? scratch : in;
    PassArgs pa{};
    pa.lda = lda; pa.nb = nb; pa.d = d;

    // pass 1: the GM fastest factors on contiguous tiles
    pa.j0 = d - GM; pa.L = 1; pa.lblocks = 1;
    pa.tiles_per_item = N / ipow(n, GM);
    pa.units          = pa.tiles_per_item * nb;
    cudaError_t e = launch_pairpass<T, n, GM, 0, false>(sms, A, in, work, out, pa, st, launches);
    if (e != cudaSuccess) return e;

    int done = GM;
    while (done < d)
    {
        const int rem = d - done;
        const int g   = rem < 2 ? rem : 2; // one or two factors per later pass: rows of n^(GM-g) contiguous elements
        const bool fin = (done + g == d);
        long long L = 1;
        for (int i = 0; i < done; ++i) L *= n;
        pa.j0 = d - done - g; pa.L = L;
        pa.lblocks        = L / ipow(n, GM - g);
        pa.tiles_per_item = N / ipow(n, GM);
        pa.units          = pa.tiles_per_item * nb;
        if (fin)
            e = (g == 1) ? launch_pairpass<T, n, 1, GM - 1, true>(sms, A, work, work, out, pa, st, launches)
                         : launch_pairpass<T, n, 2, GM - 2, true>(sms, A, work, work, out, pa, st, launches);
        else if constexpr (GM <= 3)
            e = (g == 1) ? launch_pairpass<T, n, 1, GM - 1, false>(sms, A, work, work, out, pa, st, launches)
                         : launch_pairpass<T, n, 2, GM - 2, false>(sms, A, work, work, out, pa, st, launches);
        else
            e = cudaErrorNotSupported;
        if (e != cudaSuccess) return e;
        done += g;
    }
    // the intermediate vectors are dead: drop their (dirty) lines from L2 instead of writing them back
    return launch_discard<T>(work, nb, N, st);
}

} // namespace kron
