// kernel_wspec.cuh -- warp-specialised two-phase kernel for n = 4, d = 5 and 6 ("wspec" path).
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78) for BASELINE configs 3 (d = 6) and 5 (d = 5).
//
// Why: the single-role register-tile kernel (kernel_regtile.cuh) moves every element through shared
// memory six times per item plus 48 broadcast factor loads per thread, and ncu shows the shared-memory
// pipe (65-80 % busy), not the FP64 pipe (53 %), limiting it.  Passes that apply three factors at once on
// 64-value register tiles halve that traffic, but a 64-value tile and a 64-value run accumulator do not
// fit one thread's registers together -- so the two halves of the work get different warps:
//
//   P1 warps: thread r owns column r = (the three fastest indices) of the item(s) of the step, which one
//       elected thread fetched with ONE 1-D TMA bulk copy per item (cp.async.bulk + mbarrier, issued a
//       full step ahead) into a linear 2-stage ring.  Column-wise reads of a linear buffer are bank-
//       conflict free.  It applies the slow factors (d = 6: three, on a 64-value register tile; d = 5: two,
//       on 16 values for each of the step's four items) and writes the result to the padded exchange
//       buffer E (pitch 64 values + 16 bytes).
//   P2 warps: thread p owns row p of E (64 contiguous values = the three fastest indices): it reads the row
//       in four slices of 16 (128-bit, conflict free thanks to the pitch), applies the two fastest
//       factors to the slice and folds the third into 64 register accumulators
//       (acc[i'][m] += M[i'][j] * t[m]).  Accumulators persist across consecutive items with the same
//       output pointer.  A flush writes the accumulators back into the thread's own row of E (which it
//       still owns), and after a named barrier the P2 threads read E column-wise and issue REDG whose
//       lanes are contiguous -- sector-complete (128-byte-strided REDs are 7x slower,
//       profiles/microbench_r01.jsonl).
//   Handoff: mbarriers e_full / e_empty on E; factor matrices are staged by the P1 threads with cp.async one
//       step ahead into a 4-deep ring (named barrier among the P1 warps, then published to P2 by e_full).
//   d = 5 runs four independent item streams side by side (slot q), so that run accumulation still sees
//       consecutive batch items.
//
// Shared-memory traffic per item: TMA write + P1 read + P1 write + P2 read = 4 x N x sizeof(T), and about half
// the broadcast factor loads.  Numerics: dot products run k ascending; slow factors first, then fast.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh"
#include <atomic>

namespace kron
{

template<typename T, int D>
struct Wspec4
{
    static constexpr int N       = ipow(4, D);
    static constexpr int RPI     = N / 64;                         // rows (of 64 values) per item
    static constexpr int IPS     = 64 / RPI;                       // item streams (slots) per CTA
    static constexpr int PITCH   = 64 + 16 / (int)sizeof(T);       // padded row of E, in elements
    static constexpr int NE      = (sizeof(T) == 4) ? 2 : 1;       // exchange buffers (fp64: shared memory allows one)
    static constexpr int MINB    = (sizeof(T) == 4) ? 3 : 2;       // resident CTAs per SM aimed at
    static constexpr int NMB     = 4;                              // factor-matrix ring depth
    static constexpr int MSTR    = D * 16 + 16 / (int)sizeof(T);   // per-item factor block
    static constexpr int MEL     = IPS * D * 16;                   // factor elements per step
    static constexpr int LD      = (MEL + 63) / 64;                // ... per P1 thread
    static constexpr int THREADS = 128;
    static constexpr int IN_EL   = 2 * IPS * N;                    // 2 stages, linear
    static constexpr int E_EL    = NE * 64 * PITCH;
    static constexpr int MS_EL   = NMB * IPS * MSTR;
    static constexpr int SMEM    = (IN_EL + E_EL + MS_EL) * (int)sizeof(T) + 8 * (2 + 2 * NE) + 8 * IPS + 16;
    static_assert(D == 5 || D == 6, "wspec covers d = 5, 6");
};

template<typename T, int STRIDE>
__device__ __forceinline__ void tile64_apply(T (&x)[64], const T *Ms)
{
    T m[16];
    lds16<T>(Ms, m);
#pragma unroll
    for (int f = 0; f < 16; ++f)
    {
        const int base = (STRIDE == 1) ? f * 4 : (STRIDE == 4 ? (f / 4) * 16 + (f % 4) : f);
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

// tile16_apply with the factor already in registers (row-major m[i*4+k])
template<typename T, int STRIDE>
__device__ __forceinline__ void tile16_apply_m(T (&x)[16], const T (&m)[16])
{
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

__device__ __forceinline__ void bar_sync_named(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// OR-reduction of a predicate over the 64 threads of a named barrier
__device__ __forceinline__ bool bar_or_named(int id, bool pred)
{
    int res;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\tbar.red.or.pred q, %2, 64, p;\n\tselp.b32 %0, 1, 0, q;\n\t}"
                 : "=r"(res) : "r"((int)pred), "r"(id) : "memory");
    return res != 0;
}

template<typename T, int D>
__global__ void __launch_bounds__(128, Wspec4<T, D>::MINB)
kron_wspec4_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                   const int lda, const int nb, const long long items_per_cta, const int sms)
{
    using C = Wspec4<T, D>;
    constexpr int N = C::N, RPI = C::RPI, IPS = C::IPS, PITCH = C::PITCH, NMB = C::NMB, NE = C::NE;
    constexpr int MSTR = C::MSTR, LD = C::LD;
    constexpr int VE = 16 / (int)sizeof(T);
    constexpr unsigned ITEM_BYTES = N * sizeof(T);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *IN          = reinterpret_cast<T *>(smem_raw);   // [2][IPS][N]   linear item buffers (TMA ring)
    T *E           = IN + C::IN_EL;                     // [64][PITCH]   exchange / flush buffer
    T *MS          = E + C::E_EL;                       // [NMB][IPS][MSTR]
    uint64_t *bars = reinterpret_cast<uint64_t *>(MS + C::MS_EL + (C::MS_EL & 1));
    uint64_t *full_in = bars, *e_full = bars + 2, *e_empty = bars + 2 + NE;

    // this CTA's items, split into IPS consecutive streams of `len` items (slot q: [K0 + q*len, ...))
    const long long K0 = (long long)blockIdx.x * items_per_cta;
    long long K1       = K0 + items_per_cta;
    if (K1 > nb) K1 = nb;
    if (K1 <= K0) return;
    const long long len = (K1 - K0 + IPS - 1) / IPS;
    const int nsteps    = (int)len;

    const int t = threadIdx.x;
    if (t == 0)
    {
        mbar_init(full_in + 0, 1);
        mbar_init(full_in + 1, 1);
        for (int i = 0; i < NE; ++i) { mbar_init(e_full + i, 64); mbar_init(e_empty + i, 64); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto item_of = [&](int s, int slot) -> long long {
        const long long k = K0 + (long long)slot * len + s;
        return (s < nsteps && k < K0 + (long long)(slot + 1) * len && k < K1) ? k : -1;
    };

    // Warp w runs on SM sub-partition w % 4.  Co-resident CTAs (b, b + #SMs, ...) swap the roles of their
    // warp pairs so that every sub-partition hosts P1 and P2 warps (their instruction mixes differ).
    const bool swap_roles = ((blockIdx.x / sms) & 1) != 0;
    if ((t < 64) != swap_roles)
    {
        // =================================================================== P1: slow factors, column-wise
        const int r = t & 63;
        int l_ok[LD], l_dst[LD], l_src[LD], l_q[LD], l_j[LD];
#pragma unroll
        for (int i = 0; i < LD; ++i)
        {
            const int e  = r + i * 64;
            const int eq = e / (D * 16), ej = (e / 16) % D, ec = (e % 16) / 4, er = e % 4;
            l_ok[i]  = e < C::MEL;
            l_q[i]   = eq; l_j[i] = ej;
            l_dst[i] = eq * MSTR + ej * 16 + er * 4 + ec; // row-major 4x4
            l_src[i] = er + ec * lda;
        }
        auto mat_ptrs = [&](int s, const T *(&ap)[LD]) {
#pragma unroll
            for (int i = 0; i < LD; ++i)
            {
                const long long k = l_ok[i] ? item_of(s, l_q[i]) : -1;
                ap[i] = (k >= 0) ? A[k * D + l_j[i]] : nullptr;
            }
        };
        auto stage_mats = [&](int s, const T *const (&ap)[LD]) {
#pragma unroll
            for (int i = 0; i < LD; ++i)
                if (ap[i]) cp_async_elem<T>(MS + (s % NMB) * (IPS * MSTR) + l_dst[i], ap[i] + l_src[i]);
        };
        auto in_ptrs = [&](int s, const T *(&ip)[IPS]) {
#pragma unroll
            for (int q = 0; q < IPS; ++q)
            {
                const long long k = item_of(s, q);
                ip[q] = (k >= 0) ? in[k] : nullptr;
            }
        };
        // operands of step s -> ring stage s&1: one TMA bulk copy per 16-byte-aligned vector (thread 0);
        // vectors that are only T-aligned are copied element-wise, each thread its own column
        auto stage_data = [&](int s, const T *const (&ip)[IPS]) {
            if (s >= nsteps) return;
            T *dst = IN + (s & 1) * (IPS * N);
            unsigned bytes = 0;
#pragma unroll
            for (int q = 0; q < IPS; ++q)
            {
                if (!ip[q]) continue;
                if (aligned16(ip[q])) bytes += ITEM_BYTES;
                else
                {
#pragma unroll 4
                    for (int h = 0; h < RPI; ++h) cp_async_elem<T>(dst + q * N + h * 64 + r, ip[q] + h * 64 + r);
                }
            }
            if (r == 0)
            {
                fence_proxy_async(); // generic-proxy reads of this stage (ordered by the P1 barrier) come first
                mbar_arrive_expect_tx(full_in + (s & 1), bytes);
#pragma unroll
                for (int q = 0; q < IPS; ++q)
                    if (ip[q] && aligned16(ip[q])) tma_load_1d(dst + q * N, ip[q], ITEM_BYTES, full_in + (s & 1));
            }
        };

        const T *ip_cur[IPS], *ip_nxt[IPS], *ap[LD];
        in_ptrs(0, ip_cur);
        stage_data(0, ip_cur);
        mat_ptrs(0, ap);
        stage_mats(0, ap);
        cp_async_commit();
        in_ptrs(1, ip_nxt);
        mat_ptrs(1, ap);
        unsigned par_in0 = 0, par_in1 = 0;

        for (int s = 0; s < nsteps; ++s)
        {
            const int st = s & 1;
            const T *ip_n2[IPS], *ap_n2[LD]; // pointer pipeline: fetched now, used next step
            in_ptrs(s + 2, ip_n2);
            mat_ptrs(s + 2, ap_n2);

            cp_async_wait_all();     // my element-wise copies for step s (factors, unaligned vectors)
            bar_sync_named(1);       // ... and those of the other P1 threads; everyone left step s-1,
                                     // so ring stage (s+1)&1 and factor slot (s+1)%4 are free
            stage_data(s + 1, ip_nxt);
            stage_mats(s + 1, ap);
            cp_async_commit();
            if (st == 0) { mbar_wait(full_in + 0, par_in0); par_in0 ^= 1; }
            else         { mbar_wait(full_in + 1, par_in1); par_in1 ^= 1; }

            const T *Xs = IN + st * (IPS * N) + r;
            const T *Ms = MS + (s % NMB) * (IPS * MSTR);
            const int eb = s % NE;          // exchange buffer of this step
            T *Eb        = E + eb * 64 * PITCH;
            // parity of the (s / NE)-th use of that buffer; the first use passes on a fresh barrier
            const unsigned par_empty = ((unsigned)(s / NE) & 1u) ^ 1u;
            if constexpr (D == 6)
            {
                T x[64]; // x[h], h = (i0, i1, i2)
#pragma unroll
                for (int h = 0; h < 64; ++h) x[h] = Xs[h * 64];
                tile64_apply<T, 1>(x, Ms + 2 * 16);
                tile64_apply<T, 4>(x, Ms + 1 * 16);
                tile64_apply<T, 16>(x, Ms + 0 * 16);
                mbar_wait(e_empty + eb, par_empty);
#pragma unroll
                for (int h = 0; h < 64; ++h) Eb[h * PITCH + r] = x[h];
            }
            else
            {
                // d = 5: a thread takes VE adjacent columns of one slot (128-bit accesses) and only that
                // slot's two slow factors; fp64 needs two passes to cover the four slots
                constexpr int TPS = 64 / VE, SPP = 64 / TPS, NPASS = IPS / SPP;
                const int qloc = r / TPS, cv = (r % TPS) * VE;
                T x[NPASS][VE][16]; // x[pass][column][h], h = (i0, i1)
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps)
                {
                    const int qs  = ps * SPP + qloc;
                    const T *src = IN + st * (IPS * N) + qs * N + cv;
#pragma unroll
                    for (int h = 0; h < 16; ++h)
                    {
                        if constexpr (sizeof(T) == 8)
                        {
                            const double2 w = *reinterpret_cast<const double2 *>(src + h * 64);
                            x[ps][0][h] = w.x; x[ps][1][h] = w.y;
                        }
                        else
                        {
                            const float4 w = *reinterpret_cast<const float4 *>(src + h * 64);
                            x[ps][0][h] = w.x; x[ps][1][h] = w.y; x[ps][2][h] = w.z; x[ps][3][h] = w.w;
                        }
                    }
                    T m1[16], m0[16];
                    lds16<T>(Ms + qs * MSTR + 1 * 16, m1);
                    lds16<T>(Ms + qs * MSTR + 0 * 16, m0);
#pragma unroll
                    for (int v = 0; v < VE; ++v)
                    {
                        tile16_apply_m<T, 1>(x[ps][v], m1);
                        tile16_apply_m<T, 4>(x[ps][v], m0);
                    }
                }
                mbar_wait(e_empty + eb, par_empty);
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps)
                {
                    const int qs = ps * SPP + qloc;
#pragma unroll
                    for (int h = 0; h < 16; ++h)
                    {
                        T *dst = Eb + (qs * 16 + h) * PITCH + cv;
                        if constexpr (sizeof(T) == 8) *reinterpret_cast<double2 *>(dst) = make_double2(x[ps][0][h], x[ps][1][h]);
                        else *reinterpret_cast<float4 *>(dst) = make_float4(x[ps][0][h], x[ps][1][h], x[ps][2][h], x[ps][3][h]);
                    }
                }
            }
            mbar_arrive(e_full + eb);

#pragma unroll
            for (int q = 0; q < IPS; ++q) { ip_nxt[q] = ip_n2[q]; }
#pragma unroll
            for (int i = 0; i < LD; ++i) ap[i] = ap_n2[i];
        }
    }
    else
    {
        // =================================================================== P2: fast factors + accumulation, row-wise
        const int p = t & 63, q = p / RPI, hrow = p % RPI;
        T acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = T(0);
        long long kq = item_of(0, q);
        T *o_cur     = (kq >= 0) ? out[kq] : nullptr;
        kq           = item_of(1, q);
        T *o_next    = (kq >= 0) ? out[kq] : nullptr; // output pointers are fetched two steps ahead

        for (int s = 0; s < nsteps; ++s)
        {
            const long long k  = item_of(s, q);
            const long long k2 = item_of(s + 2, q);
            T *o_next2 = (k2 >= 0) ? out[k2] : nullptr;
            const int eb = s % NE;
            T *Eb        = E + eb * 64 * PITCH;
            T *erow      = Eb + p * PITCH;
            mbar_wait(e_full + eb, (unsigned)(s / NE) & 1u);
            const T *Mq = MS + (s % NMB) * (IPS * MSTR) + q * MSTR;
            if (k >= 0)
            {
                // fp32 has the registers to keep the two fastest factors for the whole row
                [[maybe_unused]] T mf[16], mg[16];
                if constexpr (sizeof(T) == 4)
                {
                    lds16<T>(Mq + (D - 1) * 16, mf);
                    lds16<T>(Mq + (D - 2) * 16, mg);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    T x[16];
#pragma unroll
                    for (int c = 0; c < 16 / VE; ++c)
                    {
                        if constexpr (sizeof(T) == 8)
                        {
                            const double2 w = *reinterpret_cast<const double2 *>(erow + j * 16 + c * VE);
                            x[2 * c] = w.x; x[2 * c + 1] = w.y;
                        }
                        else
                        {
                            const float4 w = *reinterpret_cast<const float4 *>(erow + j * 16 + c * VE);
                            x[4 * c] = w.x; x[4 * c + 1] = w.y; x[4 * c + 2] = w.z; x[4 * c + 3] = w.w;
                        }
                    }
                    if constexpr (sizeof(T) == 4)
                    {
                        tile16_apply_m<T, 1>(x, mf);
                        tile16_apply_m<T, 4>(x, mg);
                    }
                    else
                    {
                        tile16_apply<T, 1>(x, Mq + (D - 1) * 16); // fastest index
                        tile16_apply<T, 4>(x, Mq + (D - 2) * 16);
                    }
                    const T *M3 = Mq + (D - 3) * 16;
                    const T c0 = M3[0 * 4 + j], c1 = M3[1 * 4 + j], c2 = M3[2 * 4 + j], c3 = M3[3 * 4 + j];
#pragma unroll
                    for (int m = 0; m < 16; ++m)
                    {
                        acc[m]      += c0 * x[m];
                        acc[16 + m] += c1 * x[m];
                        acc[32 + m] += c2 * x[m];
                        acc[48 + m] += c3 * x[m];
                    }
                }
            }
            // ---- flush when the run of equal output pointers ends here.  The accumulators go back into the
            // thread's own row of E (it is the only reader of that row), then the item's threads read E
            // column-wise so that the REDs of neighbouring lanes are contiguous.
            const bool my_flush = (k >= 0) && (o_next != o_cur); // uniform over the threads of one item
            if (my_flush)
            {
#pragma unroll
                for (int c = 0; c < 64 / VE; ++c)
                {
                    if constexpr (sizeof(T) == 8) *reinterpret_cast<double2 *>(erow + c * VE) = make_double2(acc[2 * c], acc[2 * c + 1]);
                    else *reinterpret_cast<float4 *>(erow + c * VE) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                }
#pragma unroll
                for (int i = 0; i < 64; ++i) acc[i] = T(0);
            }
            if constexpr (RPI == 64)
            {
                // d = 6: the item spans both P2 warps
                if (my_flush)
                {
                    bar_sync_named(2);
#pragma unroll 8
                    for (int h = 0; h < 64; ++h) red_add(o_cur + h * 64 + p, Eb[h * PITCH + p]);
                }
            }
            else
            {
                // d = 5: the item's 16 rows belong to one aligned half-warp; each thread takes 4 columns
                __syncwarp();
                if (my_flush)
                {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll 4
                        for (int h = 0; h < 16; ++h)
                            red_add(o_cur + h * 64 + cc * 16 + hrow, Eb[(q * 16 + h) * PITCH + cc * 16 + hrow]);
                }
            }
            mbar_arrive(e_empty + eb); // every read of E and of this step's factors is done
            o_cur  = o_next;
            o_next = o_next2;
        }
    }
}

template<typename T, int D>
static cudaError_t launch_wspec4(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                 cudaStream_t st, std::atomic<long long> &launches)
{
    using C  = Wspec4<T, D>;
    auto kfn = kron_wspec4_kernel<T, D>;
    int ctas_per_sm = 0;
    cudaError_t e = kernel_setup(kfn, C::THREADS, C::SMEM, ctas_per_sm); // per device (common.cuh)
    if (e != cudaSuccess) return e;
    long long grid = (long long)sms * ctas_per_sm;
    long long ipc  = ((long long)nb + grid - 1) / grid; // items per CTA
    // CTA and stream boundaries on multiples of 32 items when there is enough work, so that ASGarD-style
    // runs of equal output pointers do not straddle streams
    const long long align = 32LL * C::IPS;
    if (ipc > 2 * align) ipc = (ipc + align - 1) / align * align;
    grid = ((long long)nb + ipc - 1) / ipc;
    kfn<<<(int)grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb, ipc, sms);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// cudaErrorNotSupported when (n, d) is outside the family
template<typename T>
static cudaError_t run_wspec(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                             cudaStream_t st, std::atomic<long long> &launches, const char *&last_path)
{
    if (n != 4 || (d != 5 && d != 6)) return cudaErrorNotSupported;
    cudaError_t e = (d == 5) ? launch_wspec4<T, 5>(sms, A, lda, in, out, nb, st, launches)
                             : launch_wspec4<T, 6>(sms, A, lda, in, out, nb, st, launches);
    last_path = "wspec";
    return e;
}

} // namespace kron
