// kernel_wspec5.cuh -- warp-specialised two-phase kernel for n = 4, d = 5 ("wspec5" path, BASELINE config 5).
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78).
//
// Same split of the work as kernel_wspec.cuh (P1 warps: the two slow factors, column-wise; P2 warps: the three
// fast factors, row-wise, with run accumulators in registers), re-cut so that more warps are resident:
// ncu on the 4-items-per-step version showed two 255-register warps per scheduler issuing 37 % of the cycles,
// the consumer waiting on a single exchange buffer.  Here
//   * a step is TWO items (slot q = P1 warp q = P2 warp q), so the exchange buffer E (32 rows) can be double
//     buffered and three (fp64) or more (fp32) CTAs fit one SM;
//   * a row of E is shared by two P2 threads: lane l and lane l+16 both read the row's 16-value slices, but
//     each computes only two of the four fastest output indices (i4' in {2hf, 2hf+1}), so no arithmetic is
//     duplicated and a thread carries 32 accumulators instead of 64;
//   * an item's flush is the business of ONE warp (__syncwarp, no named barrier);
//   * a CTA is IPS independent pipelines (P1 warp q -> P2 warp q) that share nothing: every mbarrier has one
//     producer warp and one consumer warp, there is no CTA-wide or named barrier in the loop;
//   * whole items are pulled into L2 two steps ahead with one cp.async.bulk.prefetch.L2 each; the TMA copy into
//     the 2-stage ring follows one step ahead and brings the item's five factors along on the same mbarrier --
//     one 640-byte copy when they are contiguous (dense batches), one per factor or per column otherwise,
//     element-wise cp.async only when alignment rules TMA out.  P1 forwards the three fast factors to P2 inside
//     the exchange buffer, so P2 waits on one mbarrier per step and P1 on two;
// Per output element the products are accumulated k ascending from 0 like multiply_transpose
// (kronmult.cu:66-70); factors are applied slowest index first.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh"
#include "kernel_wspec.cuh"
#include <atomic>

namespace kron
{

template<typename T>
struct Wspec5
{
    static constexpr int D       = 5;
    static constexpr int N       = 1024;
    static constexpr int IPS     = 2;                              // item streams (warp pairs) per CTA
    static constexpr int PITCH   = 64 + 16 / (int)sizeof(T);       // padded row of E, in elements
    static constexpr int NE      = 2;                              // exchange buffers per stream
    static constexpr int NST     = 2;                              // TMA ring stages per stream
    static constexpr int MINB    = (sizeof(T) == 4) ? 4 : 3;       // resident CTAs per SM aimed at
    static constexpr int MSTR    = D * 16;                         // an item's factors: 5 column-major 4x4 blocks
    static constexpr int STG     = N + MSTR;                       // ring stage: the vector, then its factors
    static constexpr int EBUF    = 16 * PITCH + 3 * 16;            // exchange buffer: 16 rows, then factors 2..4 for P2
    static constexpr int THREADS = 64 * IPS;
    static constexpr int IN_EL   = IPS * NST * STG;
    static constexpr int E_EL    = IPS * NE * EBUF;
    static constexpr int NBQ     = NST + 2 * NE;                   // mbarriers per stream
    static constexpr int SMEM    = (IN_EL + E_EL) * (int)sizeof(T) + 8 * IPS * NBQ + 16;
};

__device__ __forceinline__ void l2_prefetch_bulk(const void *gsrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// mbarrier / TMA wrappers on 32-bit shared-window addresses that are computed once per thread (the generic-pointer
// wrappers of kernel_regtile.cuh made ptxas re-derive the window base with S2R SR_CgaCtaId at every use)
__device__ __forceinline__ void mbar_arrive_a(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_a(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_a(unsigned dst, const void *gsrc, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
template<typename T>
__device__ __forceinline__ void cp_async_elem_a(unsigned dst, const T *gsrc)
{
    if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gsrc) : "memory");
}

// Two-wide values: a thread works on pairs of independent outputs that share the factor element.  For float the
// pair operations are the packed FFMA2 / FMUL2 of sm_100 (one issue slot for two FMAs; the scalar operand is
// broadcast by the instruction itself), for double they are two DFMA / DMUL.
template<typename T> struct V2;
template<> struct V2<float>  { using type = float2; };
template<> struct V2<double> { using type = double2; };
__device__ __forceinline__ float2 pmul(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 pfma(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ double2 pmul(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double2 pfma(double2 a, double s, double2 c)
{
    return make_double2(fma(a.x, s, c.x), fma(a.y, s, c.y));
}

// x is a 4x4 tile of pairs indexed hi*4+lo; m is a COLUMN-major factor (m[k*4+i] = M(i,k)).
// STRIDE = 1: apply M along lo; STRIDE = 4: along hi.  Products are accumulated k ascending.
template<typename T, int STRIDE>
__device__ __forceinline__ void tile16_apply_cm2(typename V2<T>::type (&x)[16], const T (&m)[16])
{
    using P = typename V2<T>::type;
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        const int base = (STRIDE == 1) ? f * 4 : f;
        const P a0 = x[base], a1 = x[base + STRIDE], a2 = x[base + 2 * STRIDE], a3 = x[base + 3 * STRIDE];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            P dot = pmul(a0, m[i]);
            dot   = pfma(a1, m[4 + i], dot);
            dot   = pfma(a2, m[8 + i], dot);
            dot   = pfma(a3, m[12 + i], dot);
            x[base + i * STRIDE] = dot;
        }
    }
}

template<typename T, int dbg>
__global__ void __launch_bounds__(Wspec5<T>::THREADS, Wspec5<T>::MINB)
kron_wspec5_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                   const int lda, const int nb, const long long items_per_cta, const int sms)
{
    using C = Wspec5<T>;
    constexpr int D = 5, N = C::N, IPS = C::IPS, PITCH = C::PITCH, NE = C::NE, NST = C::NST;
    constexpr unsigned ITEM_BYTES = N * sizeof(T), FAC_BYTES = 16 * sizeof(T), COL_BYTES = 4 * sizeof(T);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int t = threadIdx.x;
    int lane, w;
    // opaque copies: ptxas otherwise re-reads SR_TID (20+ cycles each) wherever registers are tight
    asm volatile("mov.u32 %0, %1;" : "=r"(lane) : "r"(t & 31));
    asm volatile("mov.u32 %0, %1;" : "=r"(w) : "r"(t >> 5));
    // Warp w runs on SM sub-partition w % 4.  Co-resident CTAs swap the roles of their warp pairs so that every
    // sub-partition hosts P1 and P2 warps.
    const bool swap_roles = ((blockIdx.x / sms) & 1) != 0;
    const int q           = w % IPS;                      // item stream of this warp
    const bool is_p1      = ((w / IPS) == 0) != swap_roles;

    // shared memory of stream q: nothing is shared between streams
    constexpr int NBQ = C::NBQ, STG = C::STG, EBUF = C::EBUF;
    constexpr unsigned S = sizeof(T);
    T *IN  = reinterpret_cast<T *>(smem_raw) + q * (NST * STG);               // [NST][STG]  TMA ring
    T *E   = reinterpret_cast<T *>(smem_raw) + C::IN_EL + q * (NE * EBUF);    // [NE][EBUF]  exchange buffers
    // the same places as 32-bit shared addresses for the mbarrier / TMA instructions, derived once
    unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb)); // opaque: keeps ptxas from re-deriving the window base per use
    const unsigned a_in   = sb + q * (NST * STG) * S;
    const unsigned a_bar  = sb + (C::IN_EL + C::E_EL) * S + q * NBQ * 8;
    const unsigned b_full = a_bar, b_efull = a_bar + 8 * NST, b_eempty = a_bar + 8 * (NST + NE);

    // this CTA's items, split into IPS consecutive streams: stream q takes [kq0, kq0 + cnt)
    const long long K0 = (long long)blockIdx.x * items_per_cta;
    long long K1       = K0 + items_per_cta;
    if (K1 > nb) K1 = nb;
    if (K1 <= K0) return;
    const int tot = (int)(K1 - K0);
    const int len = (tot + IPS - 1) / IPS;
    int cnt       = tot - q * len;
    cnt           = cnt < 0 ? 0 : (cnt > len ? len : cnt);
    const long long kq0 = K0 + (long long)q * len;

    if (t < IPS)
    {
        uint64_t *b = reinterpret_cast<uint64_t *>(smem_raw + (C::IN_EL + C::E_EL) * sizeof(T)) + t * NBQ;
        for (int i = 0; i < NBQ; ++i) mbar_init(b + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cnt == 0) return;

    if (is_p1)
    {
        // =================================================================== P1: two slow factors, column-wise
        // Running pointers into the pointer arrays: lane j < 5 reads the pointer of factor j (and every lane that of
        // factor 0, to test contiguity without a shuffle), every lane the vector pointer.
        const T *const *pa  = A + (kq0 * D + (lane < D ? lane : 0));
        const T *const *pa0 = A + kq0 * D;
        T *const *pin       = in + kq0;
        const bool lda4  = (lda == 4);
        const bool lda16 = ((lda * (int)S) % 16 == 0);
        // Vector and factors of one step -> ring stage st, completion on ONE mbarrier.  By TMA where alignment allows
        // (vector: one bulk copy; factors: one copy per item when the five blocks are contiguous -- dense batches --
        // else one per factor or per column), element-wise cp.async otherwise.  Returns true if any element-wise
        // copy was issued (the consumer then also waits for its cp.async group).
        // No proxy fence: this warp's reads of the stage (step s-1) have returned -- their values were consumed
        // before the __syncwarp that ended that step -- so the bulk copies cannot overtake them.
        auto stage_step = [&](bool live, int st, const T *ip, const T *ap, const T *ap0) -> bool {
            if (!live) return false;
            const unsigned dst = a_in + st * (STG * S), fdst = dst + N * S;
            const unsigned bar = b_full + 8 * st;
            const bool vtma    = aligned16(ip);
            // fast path: five 16-byte aligned factor blocks back to back
            const bool contig  = lda4 && aligned16(ap0) && __all_sync(0xffffffffu, lane >= D || ap == ap0 + lane * 16);
            bool elem = !vtma;
            if (!vtma)
            {
#pragma unroll 8
                for (int h = 0; h < N / 32; ++h) cp_async_elem_a<T>(dst + (h * 32 + lane) * S, ip + h * 32 + lane);
            }
            if (contig)
            {
                if (lane == 0)
                {
                    if constexpr ((dbg & 16) != 0) fence_proxy_async();
                    mbar_expect_tx_a(bar, (vtma ? ITEM_BYTES : 0u) + D * FAC_BYTES);
                    if (vtma) tma_load_a(dst, ip, ITEM_BYTES, bar);
                    tma_load_a(fdst, ap0, D * FAC_BYTES, bar);
                }
            }
            else
            {
                const bool a16   = __all_sync(0xffffffffu, lane >= D || aligned16(ap));
                const bool ftma  = a16 && (lda4 || lda16);
                if (lane == 0)
                {
                    const unsigned bytes = (vtma ? ITEM_BYTES : 0u) + (ftma ? D * FAC_BYTES : 0u);
                    if (bytes) mbar_expect_tx_a(bar, bytes); else mbar_arrive_a(bar);
                    if (vtma) tma_load_a(dst, ip, ITEM_BYTES, bar);
                }
                __syncwarp();
                if (ftma && lda4) { if (lane < D) tma_load_a(fdst + lane * FAC_BYTES, ap, FAC_BYTES, bar); }
                else if (ftma)
                {
                    const T *apj = reinterpret_cast<const T *>(
                        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ap), (lane >> 2) % D));
                    if (lane < 4 * D) tma_load_a(fdst + lane * COL_BYTES, apj + (long long)(lane & 3) * lda, COL_BYTES, bar);
                }
                else
                {
                    elem = true;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                    {
                        const int e  = lane + 32 * i; // element e = factor e/16, column (e%16)/4, row e%4
                        const T *apj = reinterpret_cast<const T *>(
                            __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ap), (e >> 4) % D));
                        if (e < D * 16) cp_async_elem_a<T>(fdst + e * S, apj + (e & 3) + (long long)((e >> 2) & 3) * lda);
                    }
                }
            }
            return elem;
        };
        auto l2_pull = [&](const T *ip) {
            if (lane == 0 && ip && aligned16(ip)) l2_prefetch_bulk(ip, ITEM_BYTES);
        };
        auto ld_in  = [&](int s) -> const T * { return (s < cnt) ? pin[s] : nullptr; };
        auto ld_ap  = [&](int s) -> const T * { return (s < cnt) ? pa[(long long)s * D] : nullptr; };
        auto ld_ap0 = [&](int s) -> const T * { return (s < cnt) ? pa0[(long long)s * D] : nullptr; };

        // prologue: step 0 is staged, step 1 pulled into L2; pointers of steps 1 and 2 are on their way
        bool el_cur = stage_step(true, 0, ld_in(0), ld_ap(0), ld_ap0(0));
        cp_async_commit();
        const T *ip_b = ld_in(1), *ip_c = ld_in(2);          // inside the loop: vectors of steps s+1 and s+2
        const T *ap_b = ld_ap(1), *ap0_b = ld_ap0(1);        //                  factor pointers of step s+1
        l2_pull(ip_b);

        for (int s = 0; s < cnt; ++s)
        {
            const int st = s & 1;
            const T *ip_d  = ld_in(s + 3);       // pointer pipeline: fetched now, used in later steps
            const T *ap_c  = ld_ap(s + 2);
            const T *ap0_c = ld_ap0(s + 2);

            // stage st^1 was last read by this warp in step s-1
            const bool el_nxt = stage_step(s + 1 < cnt, st ^ 1, ip_b, ap_b, ap0_b);
            cp_async_commit();
            l2_pull(ip_c);                       // step s+2; its pointer was fetched a step ago
            if (el_cur)
            {
                // element-wise route: the copies of step s were committed one group ago
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
            }
            mbar_wait_a(b_full + 8 * st, (unsigned)(s >> 1) & 1u);

            using P = typename V2<T>::type;
            P x[16]; // x[h] = my two adjacent columns of row h = (i0, i1)
            const T *stage = IN + st * STG;
            {
                const T *src = stage + 2 * lane;
#pragma unroll
                for (int h = 0; h < 16; ++h) x[h] = *reinterpret_cast<const P *>(src + h * 64);
                T m1[16], m0[16];
                lds16<T>(stage + N + 1 * 16, m1);
                lds16<T>(stage + N + 0 * 16, m0);
                if constexpr ((dbg & 8) == 0)
                {
                    tile16_apply_cm2<T, 1>(x, m1);
                    tile16_apply_cm2<T, 4>(x, m0);
                }
                else x[0].x += m1[0] + m0[0];
            }
            // the three fast factors travel to P2 with the rows: 48 values, 16 bytes per lane
            constexpr int FV = 16 / (int)S, FL = 48 / FV; // values per lane, lanes taking part
            [[maybe_unused]] float4 fw;
            if (lane < FL) fw = *reinterpret_cast<const float4 *>(stage + N + 2 * 16 + lane * FV);
            const int eb = s & 1;
            T *Eb        = E + eb * EBUF;
            mbar_wait_a(b_eempty + 8 * eb, (((unsigned)(s >> 1)) & 1u) ^ 1u); // first use passes on a fresh barrier
#pragma unroll
            for (int h = 0; h < 16; ++h) *reinterpret_cast<P *>(Eb + h * PITCH + 2 * lane) = x[h];
            if (lane < FL) *reinterpret_cast<float4 *>(Eb + 16 * PITCH + lane * FV) = fw;
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_efull + 8 * eb);

            ip_b = ip_c; ip_c = ip_d; ap_b = ap_c; ap0_b = ap0_c;
            el_cur = el_nxt;
        }
    }
    else
    {
        // =================================================================== P2: three fast factors + run sums, row-wise
        using P = typename V2<T>::type;
        const int row = lane & 15, hf = lane >> 4;
        P acc[16]; // acc[i2' * 4 + i3'] = the pair i4' = 2 hf, 2 hf + 1
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i].x = acc[i].y = T(0);
        T *const *pout = out + kq0;
        T *o_cur  = pout[0];
        T *o_next = (cnt > 1) ? pout[1] : nullptr; // output pointers are fetched two steps ahead

        for (int s = 0; s < cnt; ++s)
        {
            T *o_next2   = (s + 2 < cnt) ? pout[s + 2] : nullptr;
            const int eb = s & 1;
            T *Eb        = E + eb * EBUF;
            T *erow      = Eb + row * PITCH;
            const T *Mq  = Eb + 16 * PITCH - 2 * 16; // factors 2, 3, 4 follow the rows (indexed Mq + j*16 below)
            mbar_wait_a(b_efull + 8 * eb, (unsigned)(s >> 1) & 1u);
            {
                // f4[k] = (F4(2hf, k), F4(2hf+1, k)): my two rows of the fastest factor; f3: the second fastest
                P f4[4];
                T f3[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) f4[k] = *reinterpret_cast<const P *>(Mq + 4 * 16 + k * 4 + 2 * hf);
                lds16<T>(Mq + 3 * 16, f3);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    T x[16], f2[4];
                    lds16<T>(erow + j * 16, x);
                    if constexpr ((dbg & 1) != 0)
                    {
                        // energy experiment: one more pass over the row slice (results discarded)
                        const unsigned a = (unsigned)__cvta_generic_to_shared(erow + j * 16);
#pragma unroll
                        for (int c = 0; c < 16 * (int)sizeof(T) / 16; ++c)
                        {
                            unsigned r0, r1, r2, r3;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a + 16 * c));
                        }
                    }
                    // fastest index: y[i3] = sum_k x[i3][k] F4(2hf + {0,1}, k)
                    P y[4];
#pragma unroll
                    for (int i3 = 0; i3 < 4; ++i3) y[i3] = pmul(f4[0], x[i3 * 4]);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
#pragma unroll
                        for (int i3 = 0; i3 < 4; ++i3) y[i3] = pfma(f4[k], x[i3 * 4 + k], y[i3]);
                    // second fastest: z[i3'] = sum_i3 F3(i3', i3) y[i3]
                    P z[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) z[i] = pmul(y[0], f3[i]);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) z[i] = pfma(y[k], f3[k * 4 + i], z[i]);
                    // third: acc[i2'][i3'] += F2(i2', j) z[i3']   (column j of F2 is contiguous)
                    if constexpr (sizeof(T) == 8)
                    {
                        const double2 v0 = reinterpret_cast<const double2 *>(Mq + 2 * 16 + j * 4)[0];
                        const double2 v1 = reinterpret_cast<const double2 *>(Mq + 2 * 16 + j * 4)[1];
                        f2[0] = v0.x; f2[1] = v0.y; f2[2] = v1.x; f2[3] = v1.y;
                    }
                    else
                    {
                        const float4 v = *reinterpret_cast<const float4 *>(Mq + 2 * 16 + j * 4);
                        f2[0] = v.x; f2[1] = v.y; f2[2] = v.z; f2[3] = v.w;
                    }
                    if constexpr ((dbg & 4) != 0)
                    {
                        // energy experiment: no arithmetic to speak of in P2 (wrong results)
#pragma unroll
                        for (int m = 0; m < 4; ++m) acc[m].x += x[m] + x[4 + m] + x[8 + m] + x[12 + m] + f2[m];
                    }
                    else if constexpr ((dbg & 2) != 0)
                    {
                        // energy experiment: a quarter of the third factor's FMAs only (wrong results)
#pragma unroll
                        for (int m = 0; m < 4; ++m) acc[m] = pfma(z[m], f2[0], acc[m]);
                    }
                    else
                    {
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
                            for (int m = 0; m < 4; ++m) acc[i2 * 4 + m] = pfma(z[m], f2[i2], acc[i2 * 4 + m]);
                    }
                }
            }
            // ---- flush when the run of equal output pointers ends here: the accumulators go back into the rows
            // of E this warp owns, then the warp reads E column-wise so that the REDs of its lanes are contiguous
            // (sector-complete).  One item = one warp, so __syncwarp orders everything.
            if (o_next != o_cur)
            {
                __syncwarp(); // both threads of a row are done reading it
#pragma unroll
                for (int i = 0; i < 16; ++i)
                {
                    *reinterpret_cast<P *>(erow + i * 4 + 2 * hf) = acc[i];
                    acc[i].x = acc[i].y = T(0);
                }
                __syncwarp();
#pragma unroll 4
                for (int h = 0; h < 16; ++h)
                {
                    red_add(o_cur + h * 64 + lane, Eb[h * PITCH + lane]);
                    red_add(o_cur + h * 64 + 32 + lane, Eb[h * PITCH + 32 + lane]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_eempty + 8 * eb); // every read of E and of this step's factors is done
            o_cur  = o_next;
            o_next = o_next2;
        }
    }
}

static std::atomic<int> g_wspec5_dbg{0}; // knob 2 of kronmult_b200_set_tuning: ablation variants (-DKRON_WSPEC5_EXPERIMENTS)

template<typename T>
static cudaError_t launch_wspec5(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                 cudaStream_t st, std::atomic<long long> &launches)
{
    using C  = Wspec5<T>;
    const int dbg = g_wspec5_dbg.load(std::memory_order_relaxed);
#ifdef KRON_WSPEC5_EXPERIMENTS
    // ablation variants behind knob 2 (profiles/ablation_wspec5_r01.md): 1 = extra shared-memory pass in P2,
    // 2 = 15 % fewer FMAs, 4 / 8 / 12 = no arithmetic in P2 / P1 / both, 16 = proxy fences before every TMA copy
    auto kfn = dbg == 1 ? kron_wspec5_kernel<T, 1> : dbg == 2 ? kron_wspec5_kernel<T, 2> : dbg == 4 ? kron_wspec5_kernel<T, 4>
             : dbg == 8 ? kron_wspec5_kernel<T, 8> : dbg == 12 ? kron_wspec5_kernel<T, 12> : dbg == 16 ? kron_wspec5_kernel<T, 16>
             : kron_wspec5_kernel<T, 0>;
#else
    auto kfn = kron_wspec5_kernel<T, 0>;
#endif
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
static cudaError_t run_wspec5(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                              cudaStream_t st, std::atomic<long long> &launches, const char *&last_path)
{
    if (n != 4 || d != 5) return cudaErrorNotSupported;
    cudaError_t e = launch_wspec5<T>(sms, A, lda, in, out, nb, st, launches);
    last_path = "wspec5";
    return e;
}

} // namespace kron
