// kernel_sym5.cuh -- symmetric single-role kernel for n = 4, d = 5 ("sym5" path, BASELINE config 5).
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78).
//
// Why (round-1 ncu of the warp-specialised wspec5 kernel): its P1 warps (two slow factors, 40 % of the FMAs) wait a
// third of the time on the exchange buffers of their P2 partner (three fast factors + flush), 14 % of the issue
// samples sit in mbarrier spins, and the exchange buffers take half of the shared memory.  Here every warp is a whole
// pipeline of its own:
//   * one warp = one stream of consecutive items; nothing is shared between warps, there is no CTA-wide or named
//     barrier and no warp-to-warp mbarrier -- the only mbarriers are the warp's own TMA "stage full" barriers;
//   * an item lands (ONE 1-D bulk TMA copy for the 1024-element vector + one for its five factors when they are
//     contiguous) in a stage of the warp's 2-deep ring and is transformed IN PLACE there:
//       phase A (column-wise, lane = two adjacent columns (i2,i3,i4), registers = the 16 values (i0,i1)): the two
//                slow factors on 16-value register tiles, written back to the same addresses;
//       phase B (row-wise, a row (i0,i1) of 64 values is shared by lanes l and l+16, each producing two of the four
//                fastest output indices): the three fast factors, the last one folded into 32 run accumulators.
//     No exchange buffer, no padding: the rows are 512 (256) bytes apart, so a straight row-wise read would be a
//     16-way bank conflict -- instead lane l visits the 16-byte chunks of its row in the order c ^ (row & 7).  All
//     eight lanes of a quarter-warp then touch eight different bank groups.  The permutation is absorbed by the
//     FACTORS: the lane reads the columns of the fast factors through the same XOR (its own, lane-private view
//     of F4 / F3 / F2), so the arithmetic stays lane-uniform and costs nothing;
//   * runs of equal output pointers are summed in registers; a flush goes back through the (finished) stage with the
//     same chunk permutation and leaves as sector-complete REDG (256 contiguous bytes per warp instruction).
// Shared memory per item: TMA write + A read + A write + B read (x2: two lanes per row) -- the same wavefronts as
// wspec5 -- but 17.7 KB per stream instead of 35 KB, and every resident warp always has work of its own.
// Numerics: products are summed k ascending within a lane's view (a lane with an odd row visits k in the order
// 2,3,0,1 for fp64); factors are applied slowest index first.  Covered by the 1e-12 / 1e-5 relative-L2 tolerance.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh"
#include "kernel_wspec5.cuh"
#include <atomic>

#ifndef KRON_SYM5_XO_REGS
#define KRON_SYM5_XO_REGS 0
#endif
#ifndef KRON_SYM5_G3_RELOAD
#define KRON_SYM5_G3_RELOAD (sizeof(T) == 8)
#endif

namespace kron
{

template<typename T, int WARPS_, int NST_ = 2>
struct Sym5
{
    static constexpr int D       = 5;
    static constexpr int N       = 1024;
    static constexpr int WARPS   = WARPS_;                            // item streams (warps) per CTA
    static constexpr int NST     = NST_;                              // TMA ring stages per stream (items staged NST-1 ahead)
    static constexpr int MSTR    = D * 16;                            // an item's factors: 5 column-major 4x4 blocks
    static constexpr int STG     = N + (sizeof(T) == 8 ? 80 : 96);    // ring stage (vector, then factors): k * 128 bytes
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int PRING   = 4;                                 // pointer ring: items s .. s+3, 64 bytes each
    static constexpr int SMEM    = WARPS * NST * STG * (int)sizeof(T) + WARPS * PRING * 64 + 8 * WARPS * NST + 16;
    static_assert((STG * sizeof(T)) % 128 == 0, "stages start on 128-byte lines (the chunk permutation assumes it)");
};

// column-major factor m[k*4+i] = M(i,k) on a 4x4 tile of scalars (tile16_apply_cm2 for single columns)
template<typename T, int STRIDE>
__device__ __forceinline__ void tile16_apply_cm(T (&x)[16], const T (&m)[16])
{
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        const int base = (STRIDE == 1) ? f * 4 : f;
        const T a0 = x[base], a1 = x[base + STRIDE], a2 = x[base + 2 * STRIDE], a3 = x[base + 3 * STRIDE];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            T dot = a0 * m[i];
            dot   = fma(a1, m[4 + i], dot);
            dot   = fma(a2, m[8 + i], dot);
            dot   = fma(a3, m[12 + i], dot);
            x[base + i * STRIDE] = dot;
        }
    }
}

// VAR bit 0: phase A on single columns (two half-passes of 16 values per lane) instead of column pairs
template<typename T, int WARPS, int MINB, int VAR>
__global__ void __launch_bounds__(Sym5<T, WARPS, ((VAR & 8) ? 3 : 2)>::THREADS, MINB)
kron_sym5_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                 const int lda, const int nb, const long long items_per_warp)
{
    using C = Sym5<T, WARPS, ((VAR & 8) ? 3 : 2)>;
    using P = typename V2<T>::type;
    constexpr int D = 5, N = C::N, NST = C::NST, STG = C::STG;
    constexpr unsigned S = sizeof(T);
    constexpr unsigned ITEM_BYTES = N * S, FAC_BYTES = 16 * S, COL_BYTES = 4 * S;
    constexpr int E = 16 / (int)S; // elements per 16-byte chunk
    constexpr bool G3_RELOAD = KRON_SYM5_G3_RELOAD;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    int lane, w;
    // opaque copies: ptxas otherwise re-reads SR_TID wherever registers are tight
    asm volatile("mov.u32 %0, %1;" : "=r"(lane) : "r"(threadIdx.x & 31));
    if constexpr (WARPS == 1) w = 0; // one warp per CTA: everything per-stream is CTA-uniform (uniform registers)
    else asm volatile("mov.u32 %0, %1;" : "=r"(w) : "r"(threadIdx.x >> 5));

    // this warp's stream of consecutive items
    const long long gw  = (long long)blockIdx.x * C::WARPS + w;
    const long long kq0 = gw * items_per_warp;
    if (kq0 >= nb) return;
    const int cnt = (int)((kq0 + items_per_warp <= nb) ? items_per_warp : (nb - kq0));

    T *IN = reinterpret_cast<T *>(smem_raw) + w * (NST * STG);
    unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb)); // opaque: keeps ptxas from re-deriving the window base per use
    const unsigned a_in   = sb + w * (NST * STG) * S;
    // pointer ring of this warp: slot (s & 3) = { A[k*5 + 0..4], in[k], out[k], - } of item k = kq0 + s, filled by
    // 8-byte cp.async three steps ahead (no registers are spent on the pointer pipeline)
    const unsigned a_pr   = sb + C::WARPS * NST * STG * S + w * (C::PRING * 64);
    const unsigned long long *PR =
        reinterpret_cast<const unsigned long long *>(smem_raw + C::WARPS * NST * STG * S) + w * (C::PRING * 8);
    const unsigned b_full = sb + C::WARPS * NST * STG * S + C::WARPS * C::PRING * 64 + w * NST * 8;

    if (lane == 0)
    {
        uint64_t *b = reinterpret_cast<uint64_t *>(smem_raw + C::WARPS * NST * STG * S + C::WARPS * C::PRING * 64) + w * NST;
        for (int i = 0; i < NST; ++i) mbar_init(b + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // Vector and factors of one item -> ring stage st, completion on ONE mbarrier.  By TMA where alignment allows
    // (vector: one bulk copy; factors: one copy per item when the five blocks are contiguous -- dense batches --
    // else one per factor or per column), element-wise cp.async otherwise.  Returns true if any element-wise copy
    // was issued (the consumer then also waits for its cp.async group).
    auto stage_item = [&](bool live, int st, const T *ip, const T *ap) -> bool {
        if (!live) return false;
        if constexpr ((VAR & 2) != 0)
        {
            // EXPERIMENT (dense, aligned batches only): no layout checks at all
            if (lane == 0)
            {
                const unsigned dst = a_in + st * (STG * S);
                const unsigned bar = b_full + 8 * st;
                mbar_expect_tx_a(bar, ITEM_BYTES + D * FAC_BYTES);
                tma_load_a(dst, ip, ITEM_BYTES, bar);
                tma_load_a(dst + N * S, ap, D * FAC_BYTES, bar);
            }
            return false;
        }
        const bool lda4  = (lda == 4);
        const bool lda16 = ((lda * (int)S) % 16 == 0);
        const unsigned dst = a_in + st * (STG * S), fdst = dst + N * S;
        const unsigned bar = b_full + 8 * st;
        const bool vtma    = aligned16(ip);
        const T *ap0       = reinterpret_cast<const T *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ap), 0));
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
                mbar_expect_tx_a(bar, (vtma ? ITEM_BYTES : 0u) + D * FAC_BYTES);
                if (vtma) tma_load_a(dst, ip, ITEM_BYTES, bar);
                tma_load_a(fdst, ap0, D * FAC_BYTES, bar);
            }
        }
        else
        {
            const bool a16  = __all_sync(0xffffffffu, lane >= D || aligned16(ap));
            const bool ftma = a16 && (lda4 || lda16);
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
    // lanes 0..4 fetch the item's factor pointers, lane 5 its input pointer, lane 6 its output pointer
    auto fetch_ptrs = [&](int s) {
        if (s < cnt && lane < 7)
        {
            const long long k = kq0 + s;
            const void *src = lane < D ? static_cast<const void *>(A + k * D + lane)
                            : lane == 5 ? static_cast<const void *>(in + k) : static_cast<const void *>(out + k);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a_pr + (s & 3) * 64 + lane * 8), "l"(src) : "memory");
        }
    };
    auto pr_in  = [&](int s) -> const T * { return reinterpret_cast<const T *>(PR[(s & 3) * 8 + 5]); };
    auto pr_ap  = [&](int s) -> const T * { return reinterpret_cast<const T *>(PR[(s & 3) * 8 + (lane < D ? lane : 0)]); };
    auto pr_out = [&](int s) -> T * { return reinterpret_cast<T *>(PR[(s & 3) * 8 + 6]); };

    // ---- lane-private view of the row-wise phase: row (i0,i1) = lane & 15, output pair i4' = 2 hf + {0,1}.
    // The lane visits chunk c of its row at position c ^ (row & 7); in element indices that is e ^ mask_e, which
    // permutes the fastest index pair-wise (fp64), the second fastest (both) and the third (fp32) as seen by it.
    const int row = lane & 15, hf = lane >> 4;
    const int mask_e = (row & 7) * E;
    const int km = mask_e & 3, i3m = (mask_e >> 2) & 3, i2m = (mask_e >> 4) & 3;
    // element offset of chunk slot t (the low three chunk-index bits) within the stage: row * 64 + ((t * E) ^ mask_e)
#if KRON_SYM5_XO_REGS
    int xo_[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) xo_[t] = row * 64 + ((t * E) ^ mask_e);
    auto xo = [&](int t) { return xo_[t]; };
#else
    const int rowm = row * 64 + mask_e; // row * 64 is a multiple of 64 > mask_e: '+' == '^' here
    auto xo = [&](int t) { return rowm ^ (t * E); };
#endif

    P acc[16]; // acc[i2' * 4 + i3'] = the pair i4' = 2 hf, 2 hf + 1 of my row
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i].x = acc[i].y = T(0);

    // prologue: pointers of items 0..2, item 0 staged, item 1 pulled into L2
    fetch_ptrs(0); fetch_ptrs(1); fetch_ptrs(2);
    cp_async_commit();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    stage_item(true, 0, pr_in(0), pr_ap(0));
    if constexpr (NST == 3) { if (cnt > 1) stage_item(true, 1, pr_in(1), pr_ap(1)); }
    cp_async_commit();
    if constexpr (NST == 2) { if (cnt > 1) l2_pull(pr_in(1)); }

    for (int s = 0; s < cnt; ++s)
    {
        const int st = (NST == 2) ? (s & 1) : (s % 3);
        // every cp.async group committed in earlier steps is complete: the pointers of items s+1 and s+2 and any
        // element-wise copies of item s (issued a whole step ago)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp(); // ... and visible to all lanes; every read of stage st^1 (item s-1) and ring slot (s+3)&3 is done
        // stage st^1 was last read by this warp in step s-1
        if constexpr (NST == 2)
        {
            if (s + 1 < cnt) stage_item(true, st ^ 1, pr_in(s + 1), pr_ap(s + 1));
            if constexpr ((VAR & 4) == 0) { if (s + 2 < cnt) l2_pull(pr_in(s + 2)); }
        }
        else
        {
            // three stages: item s+2 goes into the stage item s-1 has just left
            if (s + 2 < cnt) stage_item(true, (s + 2) % 3, pr_in(s + 2), pr_ap(s + 2));
        }
        fetch_ptrs(s + 3);
        cp_async_commit();
        mbar_wait_a(b_full + 8 * st, (unsigned)((NST == 2) ? (s >> 1) : (s / 3)) & 1u);

        T *stage    = IN + st * STG;
        const T *Mq = stage + N; // factors 0..4, column-major 4x4 blocks
        // ------------------------------------------------------------ phase A: factors 0 and 1, column-wise, in place
        if constexpr ((VAR & 1) == 0)
        {
            P x[16]; // x[h] = my two adjacent columns of row h = (i0, i1)
            T *col = stage + 2 * lane;
#pragma unroll
            for (int h = 0; h < 16; ++h) x[h] = *reinterpret_cast<const P *>(col + h * 64);
            {
                T m[16];
                lds16<T>(Mq + 1 * 16, m);
                tile16_apply_cm2<T, 1>(x, m);
            }
            {
                T m[16];
                lds16<T>(Mq + 0 * 16, m);
                tile16_apply_cm2<T, 4>(x, m);
            }
#pragma unroll
            for (int h = 0; h < 16; ++h) *reinterpret_cast<P *>(col + h * 64) = x[h];
        }
        else
        {
#pragma unroll 1
            for (int hp = 0; hp < 2; ++hp)
            {
                T x[16]; // x[h] = column lane + 32 hp of row h = (i0, i1)
                T *col = stage + lane + 32 * hp;
#pragma unroll
                for (int h = 0; h < 16; ++h) x[h] = col[h * 64];
                {
                    T m[16];
                    lds16<T>(Mq + 1 * 16, m);
                    tile16_apply_cm<T, 1>(x, m);
                }
                {
                    T m[16];
                    lds16<T>(Mq + 0 * 16, m);
                    tile16_apply_cm<T, 4>(x, m);
                }
#pragma unroll
                for (int h = 0; h < 16; ++h) col[h * 64] = x[h];
            }
        }
        __syncwarp();
        // ------------------------------------------------------------ phase B: factors 2, 3, 4 + run sums, row-wise
        {
            // g4[ks] = (F4(2hf, ks ^ km), F4(2hf+1, ks ^ km)): my two rows of the fastest factor, in my visiting order
            P g4[4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) g4[ks] = *reinterpret_cast<const P *>(Mq + 4 * 16 + ((ks ^ km) * 4 + 2 * hf));
            // g3[sl * 4 + i] = F3(i, sl ^ i3m): the columns of the second fastest factor in my visiting order
            // (fp64: re-read for every slice j, there are no registers to keep them)
            auto load_g3 = [&](T (&g3)[16]) {
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                {
                    const T *c3 = Mq + 3 * 16 + (sl ^ i3m) * 4;
                    if constexpr (sizeof(T) == 8)
                    {
                        const double2 v0 = reinterpret_cast<const double2 *>(c3)[0], v1 = reinterpret_cast<const double2 *>(c3)[1];
                        g3[sl * 4 + 0] = v0.x; g3[sl * 4 + 1] = v0.y; g3[sl * 4 + 2] = v1.x; g3[sl * 4 + 3] = v1.y;
                    }
                    else
                    {
                        const float4 v = *reinterpret_cast<const float4 *>(c3);
                        g3[sl * 4 + 0] = v.x; g3[sl * 4 + 1] = v.y; g3[sl * 4 + 2] = v.z; g3[sl * 4 + 3] = v.w;
                    }
                }
            };
            [[maybe_unused]] T g3k[16];
            if constexpr (!G3_RELOAD) load_g3(g3k);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                // fastest index: y = sum_ks x[j][sl][ks] * g4[ks] (a pair of outputs i4'), folded at once into the
                // second fastest: z[i] = sum_sl F3(i, sl ^ i3m) y[sl]
                P z[4];
                T g3r[16];
                if constexpr (G3_RELOAD) load_g3(g3r);
                const T (&g3)[16] = G3_RELOAD ? g3r : g3k;
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                {
                    T xv[4];
                    if constexpr (sizeof(T) == 8)
                    {
                        // chunks t = sl*2 + {0,1} of slice j; slice j itself is not permuted (i2m == 0)
                        const double2 v0 = *reinterpret_cast<const double2 *>(stage + xo(sl * 2 + 0) + j * 16);
                        const double2 v1 = *reinterpret_cast<const double2 *>(stage + xo(sl * 2 + 1) + j * 16);
                        xv[0] = v0.x; xv[1] = v0.y; xv[2] = v1.x; xv[3] = v1.y;
                    }
                    else
                    {
                        // chunk t = (j & 1) * 4 + sl, upper bit of j outside the permutation
                        const float4 v = *reinterpret_cast<const float4 *>(stage + xo((j & 1) * 4 + sl) + (j >> 1) * 32);
                        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
                    }
                    P yy = pmul(g4[0], xv[0]);
                    yy   = pfma(g4[1], xv[1], yy);
                    yy   = pfma(g4[2], xv[2], yy);
                    yy   = pfma(g4[3], xv[3], yy);
                    if (sl == 0)
                    {
#pragma unroll
                        for (int i = 0; i < 4; ++i) z[i] = pmul(yy, g3[i]);
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < 4; ++i) z[i] = pfma(yy, g3[sl * 4 + i], z[i]);
                    }
                }
                // third: acc[i2'][i3'] += F2(i2', j ^ i2m) z[i3']   (a column of F2 is contiguous)
                T f2[4];
                const T *c2 = Mq + 2 * 16 + (j ^ i2m) * 4;
                if constexpr (sizeof(T) == 8)
                {
                    const double2 v0 = reinterpret_cast<const double2 *>(c2)[0], v1 = reinterpret_cast<const double2 *>(c2)[1];
                    f2[0] = v0.x; f2[1] = v0.y; f2[2] = v1.x; f2[3] = v1.y;
                }
                else
                {
                    const float4 v = *reinterpret_cast<const float4 *>(c2);
                    f2[0] = v.x; f2[1] = v.y; f2[2] = v.z; f2[3] = v.w;
                }
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
                    for (int m = 0; m < 4; ++m) acc[i2 * 4 + m] = pfma(z[m], f2[i2], acc[i2 * 4 + m]);
            }
        }
        // ---- flush when the run of equal output pointers ends here: the accumulators go into the (finished) stage,
        // chunk-permuted like the reads, then the warp reads the stage column-wise so that its REDs are contiguous
        T *o_cur = pr_out(s);
        if (s + 1 >= cnt || pr_out(s + 1) != o_cur)
        {
            __syncwarp(); // both lanes of a row are done reading it
#pragma unroll
            for (int i = 0; i < 16; ++i)
            {
                *reinterpret_cast<P *>(stage + row * 64 + ((i * 4 + 2 * hf) ^ mask_e)) = acc[i];
                acc[i].x = acc[i].y = T(0);
            }
            __syncwarp();
#pragma unroll 4
            for (int h = 0; h < 16; ++h)
            {
                const int mh = (h & 7) * E;
                red_add(o_cur + h * 64 + lane, stage[h * 64 + (lane ^ mh)]);
                red_add(o_cur + h * 64 + 32 + lane, stage[h * 64 + ((32 + lane) ^ mh)]);
            }
            fence_proxy_async(); // the next bulk copy into this stage follows generic-proxy writes
        }
    }
}

static std::atomic<int> g_sym5_var{0}; // knob 5 of kronmult_b200_set_tuning: kernel variant (development)

template<typename T, int WARPS, int MINB, int VAR>
static cudaError_t launch_sym5(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                               cudaStream_t st, std::atomic<long long> &launches)
{
    using C  = Sym5<T, WARPS, ((VAR & 8) ? 3 : 2)>;
    auto kfn = kron_sym5_kernel<T, WARPS, MINB, VAR>;
    int ctas_per_sm = 0;
    cudaError_t e = kernel_setup(kfn, C::THREADS, C::SMEM, ctas_per_sm);
    if (e != cudaSuccess) return e;
    const long long warps = (long long)sms * ctas_per_sm * C::WARPS;
    long long ipw = ((long long)nb + warps - 1) / warps; // items per warp
    // stream boundaries on multiples of 32 items when there is enough work, so that ASGarD-style runs of equal
    // output pointers do not straddle streams
    if (ipw > 64) ipw = (ipw + 31) / 32 * 32;
    const long long nw   = ((long long)nb + ipw - 1) / ipw;
    const long long grid = (nw + C::WARPS - 1) / C::WARPS;
    kfn<<<(int)grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb, ipw);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// cudaErrorNotSupported when (n, d) is outside the family
template<typename T>
static cudaError_t run_sym5(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                            cudaStream_t st, std::atomic<long long> &launches, const char *&last_path)
{
    if (n != 4 || d != 5) return cudaErrorNotSupported;
    const int var = g_sym5_var.load(std::memory_order_relaxed);
    cudaError_t e;
#define KRON_SYM5(W, B, V) launch_sym5<T, W, B, V>(sms, A, lda, in, out, nb, st, launches)
    if constexpr (sizeof(T) == 8)
    {
        switch (var)
        {
        default: e = KRON_SYM5(1, 8, 0); break; // single-warp CTAs, 8 per SM, no register cap to speak of
#ifdef KRON_SYM5_VARIANTS
        case 15: e = KRON_SYM5(4, 3, 0); break; // 12 warps per SM, 168 registers (spills)
        case 1: e = KRON_SYM5(4, 3, 1); break;  // ... phase A on single columns
        case 2: e = KRON_SYM5(5, 2, 0); break;  // 10 warps per SM, 200 registers
        case 3: e = KRON_SYM5(5, 2, 1); break;
        case 4: e = KRON_SYM5(4, 2, 0); break;  // 8 warps per SM
        case 5: e = KRON_SYM5(11, 1, 0); break; // 11 warps per SM, 184 registers
        case 6: e = KRON_SYM5(1, 12, 0); break; // single-warp CTAs, 12 per SM
        case 7: e = KRON_SYM5(1, 12, 1); break;
        case 8: e = KRON_SYM5(1, 10, 0); break; // ... 10 per SM, 200 registers
        case 9: e = KRON_SYM5(1, 8, 0); break;
        case 10: e = KRON_SYM5(1, 8, 2); break;  // experiments: unchecked staging (dense aligned batches only)
        case 11: e = KRON_SYM5(1, 8, 6); break;  // ... and no L2 pull
        case 12: e = KRON_SYM5(1, 12, 2); break;
        case 13: e = KRON_SYM5(1, 12, 6); break;
        case 14: e = KRON_SYM5(4, 2, 2); break;
        case 16: e = KRON_SYM5(1, 8, 8); break;  // three-stage ring
        case 17: e = KRON_SYM5(1, 8, 4); break;  // no L2 pull
#endif
        }
    }
    else
    {
        switch (var)
        {
        default: e = KRON_SYM5(1, 16, 0); break; // single-warp CTAs, 16 per SM, 119 registers
#ifdef KRON_SYM5_VARIANTS
        case 15: e = KRON_SYM5(4, 4, 0); break; // 16 warps per SM, 128 registers
        case 1: e = KRON_SYM5(4, 5, 0); break;  // 20 warps per SM, 96 registers
        case 2: e = KRON_SYM5(4, 3, 0); break;  // 12 warps per SM
        case 3: e = KRON_SYM5(6, 3, 0); break;  // 18 warps per SM, 112 registers
        case 6: e = KRON_SYM5(1, 16, 0); break; // single-warp CTAs
        case 7: e = KRON_SYM5(1, 20, 0); break;
        case 8: e = KRON_SYM5(1, 12, 0); break;
#endif
        }
    }
#undef KRON_SYM5
    last_path = "sym5";
    return e;
}

} // namespace kron
