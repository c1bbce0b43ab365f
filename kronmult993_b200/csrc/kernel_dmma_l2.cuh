// kernel_dmma_l2.cuh -- n = 8, d = 5 and 6, double precision: all passes of the DMMA route in ONE persistent kernel whose
// intermediate never leaves L2 ("dmma-l2" route; included by kernel_dmma.cuh, which provides dmma884 / dmma_sigma).
//
// Replaces cuda_kronmult (kronmult_gpu/kronmult.cu:95-130) for the reference's `large` / `realistic` cases
// (tests/kronmult_bench_gpu.cpp:71-72: n = 8, d = 6, 2 MiB per vector) and for n = 8, d = 5 of its sweep
// (tests/kronmult_fullbench_gpu.cpp:70-74).  The reference moves 2 x n^d x 8 bytes through global memory per FACTOR; the
// multi-kernel routes move 3 x n^d x 8 bytes per item through HBM (pass A in place in `input`, pass B reads it back);
// profiles/multipass_chunks_r02.md showed that cutting them into short per-chunk kernels loses more to launch gaps and
// ramp-up than L2 residency gains.
//
// Here the batch is cut into chunks of CH items (2 MiB of vectors) and ALL work is a single queue of units that
// persistent CTAs pull with one atomicAdd each:
//     A(c, item, tile): the four fastest factors on one contiguous 4096-element tile (TMA in, two in-place DMMA
//                       phases exactly as kron_dmma8_tile4_kernel), sent as ONE bulk copy (shared -> global) to a RING
//                       of R x CH vectors that belongs to the library (48 MiB by default: it stays in the 126 MB L2 and
//                       is overwritten before it is ever evicted);
//     B(c, column block): the remaining factor(s) on every item of chunk c, read back from the ring (L2 hits), summed
//                       over runs of equal output pointers in the accumulator fragments, REDG flush.
//                       d = 6: factors 0 and 1 (chained DMMAs) on 64 rows x 32 columns; d = 5: factor 0 (one DMMA
//                       product) on 8 rows x 256 columns.
// Queue order: block b = { A(b, *), B(b - LAG, *) }.  B(c) waits for a counter that the A(c) units bump (release /
// acquire through global memory), A(c) waits for B(c - R) before it overwrites that ring slot.  Every wait is on
// units with SMALLER queue positions, which running CTAs already hold, so the scheme cannot deadlock whatever the
// number of resident CTAs (no co-residency assumption, no cooperative launch); waits are bounded and trap.
// HBM traffic per item: the vector once + the output adds -- the algorithmic bytes.  `input` is not written at all, so
// the read-only-input entry points need no scratch vectors for these shapes.
// Measured (profiles/ncu_dmma_l2_r02.md, 2 GB of vectors, fraction of the FP64 roofline): d = 6 0.54 -> 0.63, d = 5 0.34 -> 0.61.
#pragma once

namespace kron
{

template<int D>
struct DmmaL2
{
    static_assert(D == 5 || D == 6, "n = 8, d = 5 or 6");
    static constexpr int CH    = D == 6 ? 4 : 8;  // items per chunk (2 MiB of vectors either way)
    // ring slots of CH vectors allocated; R <= RMAX of them are used (knob 13 / 15), B(c) is queued in block c + LAG,
    // 1 <= LAG < R (knob 14 / 16).  The blocks of d = 5 are five times shorter, so their lag counts more of them.
    static constexpr int RMAX  = D == 6 ? 8 : 32;
    static constexpr int TPI   = D == 6 ? 64 : 8;        // 4096-element tiles per item
    static constexpr long long NV = D == 6 ? 262144 : 32768; // 8^d
    static constexpr int AU    = CH * TPI;        // A positions per block
    // B units per chunk.  d = 6: 64 rows x 32 columns each (half a column tile); d = 5: 8 rows x 256 columns each
    static constexpr int BT    = D == 6 ? 2 * TPI : 16;
    static constexpr int BLOCK = AU + BT;         // queue positions per block
    static constexpr int NST   = 2;               // shared-memory tile slots
    static constexpr int SMEM  = NST * 4096 * 8 + 64;
    static constexpr size_t RING_BYTES = (size_t)RMAX * CH * NV * 8; // 64 MiB either way
};
using Dmma86F = DmmaL2<6>;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 1-D TMA with an L2 eviction-priority hint (the streamed input must not push the ring out of L2)
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy) : "memory");
}
// thread 0 only: wait until *p >= target (bounded: a lost signal is a bug, not a hang)
__device__ __forceinline__ void wait_counter(const unsigned *p, unsigned target)
{
    if (ld_acquire_u32(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_u32(p) < target)
    {
        __nanosleep(64);
        if (clock64() - t0 > 6000000000LL) __trap();
    }
}

template<int D, int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
kron_dmma8_l2_kernel(const double *const *__restrict__ A, double *const *__restrict__ in, double *const *__restrict__ out,
                      const int lda, const int nb, double *__restrict__ ring, unsigned *__restrict__ ctr, const int nchunks,
                      const int hints, const int R, const int LAG)
{
    using F = DmmaL2<D>;
    constexpr int N = 4096, THREADS = WARPS * 32, T1 = 64 / WARPS, P2 = 32 / WARPS, NST = F::NST;
    constexpr int G1 = T1 < 8 ? T1 : 8, G2 = P2 < 4 ? P2 : 4; // slices / slice pairs a warp works on at once
    static_assert(NST == 2, "slot arithmetic below is written for two tile slots");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *Rg    = reinterpret_cast<double *>(smem_raw);       // [2][4096]
    uint64_t *bar = reinterpret_cast<uint64_t *>(Rg + NST * N); // [2] "tile landed"
    __shared__ long long s_next[2];

    unsigned *queue = ctr, *doneA = ctr + 4, *doneB = ctr + 4 + nchunks;
    const long long total = (long long)(nchunks + LAG) * F::BLOCK;

    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int g = lane >> 2, q = lane & 3;
    const int lane_off0 = g + (2 * q) * lda;

    if (t == 0)
    {
        for (int i = 0; i < NST; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint64_t pol_first = 0;
    if (hints) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));

    // queue position -> unit.  type 1: A (c, item k, tile), 2: B (c, tile), 0: padding of the block structure
    struct Unit { int type, c, tile; long long k; };
    auto decode = [&](long long p) {
        Unit u{0, 0, 0, 0};
        if (p >= total) return u;
        const int b = (int)(p / F::BLOCK), r = (int)(p - (long long)b * F::BLOCK);
        if (r < F::AU)
        {
            u.c = b; u.k = (long long)b * F::CH + r / F::TPI; u.tile = r % F::TPI;
            u.type = (b < nchunks && u.k < nb) ? 1 : 0;
        }
        else
        {
            u.c = b - LAG; u.tile = r - F::AU;
            u.type = (u.c >= 0) ? 2 : 0;
        }
        return u;
    };
    // tile of an A unit -> shared-memory slot (thread 0 only); unaligned vectors are read from global memory in phase 1
    auto issue = [&](const Unit &x, int slot) {
        const double *src = in[x.k] + (long long)x.tile * N;
        if (aligned16(src))
        {
            fence_proxy_async();
            mbar_arrive_expect_tx(bar + slot, N * 8);
            if (hints) tma_load_1d_hint(Rg + slot * N, src, N * 8, bar + slot, pol_first);
            else tma_load_1d(Rg + slot * N, src, N * 8, bar + slot);
        }
        else mbar_arrive(bar + slot);
    };
    auto load_frags = [&](long long k, double (&a)[8]) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const double *p0 = A[k * D + (D - 4) + j];
            a[2 * j]     = __ldg(p0 + lane_off0);
            a[2 * j + 1] = __ldg(p0 + lane_off0 + lda);
        }
    };
    // thread 0: the bulk store of this CTA's previous A unit is complete -> publish it (release) to the B units
    int pend_c = -1;
    auto flush_pending = [&]() {
        if (t == 0 && pend_c >= 0)
        {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __threadfence();
            atomicAdd(doneA + pend_c, 1u);
            pend_c = -1;
        }
    };

    long long p_cur = 0;
    if (t == 0)
    {
        p_cur     = (long long)atomicAdd(queue, 1u);
        s_next[1] = p_cur;
    }
    __syncthreads();
    p_cur = s_next[1];
    Unit u = decode(p_cur);
    int s  = 0;           // smem slot of the current unit's tile
    unsigned parity = 0;  // bit i = phase parity of mbarrier i
    double a_nxt[8];
    const double *base_nxt = nullptr; // the next A unit's tile in global memory (its alignment decides TMA vs. plain loads)
    if (u.type == 1)
    {
        if (t == 0) issue(u, 0);
        load_frags(u.k, a_nxt);
        base_nxt = in[u.k] + (long long)u.tile * N;
    }

    for (int it = 0; p_cur < total; ++it)
    {
        unsigned pn_reg = 0;
        if (t == 0) pn_reg = atomicAdd(queue, 1u); // consumed at the unit's first barrier
        Unit un{0, 0, 0, 0};
        long long p_nxt = total;
        // publish the next position (thread 0, before a barrier), read it (everybody, after that barrier), start the
        // next A unit's tile and factor fragments on their way
        auto publish = [&]() { if (t == 0) s_next[it & 1] = (long long)pn_reg; };
        auto pickup  = [&](int pf_slot) {
            p_nxt = s_next[it & 1];
            un    = decode(p_nxt);
            if (un.type == 1)
            {
                if (t == 0) issue(un, pf_slot);
                load_frags(un.k, a_nxt);
                base_nxt = in[un.k] + (long long)un.tile * N;
            }
        };

        if (u.type == 1)
        {
            // ---------------------------------------------------------------- A unit (cf. kron_dmma8_tile4_kernel)
            double *Ec         = Rg + s * N;
            // the ring slot of chunk c was last read by the B units of chunk c - R: the counter is sampled now (relaxed: the
            // store below is control-dependent on it) and looked at after phase 2
            unsigned seenB = (unsigned)F::BT;
            if (t == 0 && u.c >= R) seenB = ld_relaxed_u32(doneB + (u.c - R));
            const double *base = base_nxt; // fetched while the previous unit was computing
            const bool vec     = aligned16(base);
            double a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a_nxt[i];
            mbar_wait(bar + s, (parity >> s) & 1u);
            parity ^= 1u << s;
            dmma_phase1_inplace<T1, G1>(Ec, vec ? nullptr : base, w, g, q, a[6], a[7], a[4], a[5]);
            publish();
            __syncthreads();
            // the other slot is the source of the previous A unit's bulk store: once that is complete, signal it and
            // refill the slot with the next tile
            flush_pending();
            pickup(s ^ 1);
            dmma_phase2_inplace<P2, G2>(Ec, w, g, q, a[2], a[3], a[0], a[1]);
            if (seenB < (unsigned)F::BT) wait_counter(doneB + (u.c - R), (unsigned)F::BT);
            fence_proxy_async(); // every thread: its phase-2 writes to the slot become visible to the bulk copy below
            __syncthreads();
            // The tile leaves as ONE bulk copy, chunk-swizzled as it stands in shared memory (slice h of the tile has its
            // 16-byte chunks at c ^ sigma(h)); the B units undo the permutation in their gather addresses.
            if (t == 0)
            {
                double *wb = ring + ((size_t)(u.c % R) * F::CH + (size_t)(u.k - (long long)u.c * F::CH)) * F::NV + (size_t)u.tile * N;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(wb), "r"((unsigned)__cvta_generic_to_shared(Ec)), "r"(N * 8) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                pend_c = u.c;
            }
            s ^= 1;
        }
        else if (u.type == 2)
        {
            // ---------------------------------------------------------------- B unit (cf. kron_dmma8_rows2_kernel)
            flush_pending(); // (this CTA's own last tile may be one of those the wait below is for)
            const long long k0 = (long long)u.c * F::CH;
            const int cnt      = (int)((k0 + F::CH <= nb) ? F::CH : (nb - k0));
            if constexpr (D == 6)
            {
                const int tile = u.tile >> 1, half = u.tile & 1; // 64 rows x 32 columns: 16-byte chunks [16 half, 16 half + 16)
                const double *rb   = ring + (size_t)(u.c % R) * F::CH * F::NV + (size_t)tile * 64;
                const int sgT      = dmma_sigma(tile); // the A units stored slice h = tile of every row chunk-swizzled by sigma(h)
                // three compact half-tile buffers (64 rows x 32 columns, 16 KiB each) in the two 32 KiB slots: items i+1 and i+2 are
                // on their way from L2 while item i is computed
                constexpr int HN = 2048, PB = P2 / 2; // elements per buffer; slice pairs per warp
                auto buf = [&](int i) { return Rg + (i % 3) * HN; };
                auto fetch = [&](int i) {
                    double *Eb        = buf(i);
                    const double *src = rb + (size_t)i * F::NV;
#pragma unroll 4
                    for (int r = 0; r < HN / 2 / THREADS; ++r)
                    {
                        const int c2 = t + r * THREADS, h = c2 >> 4, ci = c2 & 15;
                        const unsigned sa = (unsigned)__cvta_generic_to_shared(Eb + h * 32 + ((ci ^ dmma_sigma(h)) << 1));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (size_t)h * 4096 + 2 * ((ci + 16 * half) ^ sgT)) : "memory");
                    }
                };
                // factor fragments and output pointers of the whole chunk up front (two dependent global loads each)
                double fa[F::CH][4];
                double *op[F::CH + 1];
#pragma unroll
                for (int i = 0; i < F::CH; ++i)
                {
                    op[i] = nullptr;
                    fa[i][0] = fa[i][1] = fa[i][2] = fa[i][3] = 0.0;
                    if (i < cnt)
                    {
                        const double *p0 = A[(k0 + i) * D + 0], *p1 = A[(k0 + i) * D + 1];
                        fa[i][0] = __ldg(p0 + lane_off0); fa[i][1] = __ldg(p0 + lane_off0 + lda);
                        fa[i][2] = __ldg(p1 + lane_off0); fa[i][3] = __ldg(p1 + lane_off0 + lda);
                        op[i] = out[k0 + i];
                    }
                }
                op[F::CH] = nullptr;
                if (t == 0) wait_counter(doneA + u.c, (unsigned)(cnt * F::TPI));
                publish();
                __syncthreads();
                fetch(0);
                cp_async_commit();
                if (1 < cnt) fetch(1);
                cp_async_commit();
                double acc[PB][4];
#pragma unroll
                for (int j = 0; j < PB; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
#pragma unroll
                for (int i = 0; i < F::CH; ++i)
                {
                    if (i < cnt)
                    {
                        double *Ec = buf(i);
                        asm volatile("cp.async.wait_group 1;" ::: "memory"); // item i has landed (item i+1 may be pending)
                        __syncthreads(); // ... for everybody; everyone left item i-1, whose buffer item i+2 takes
                        if (i + 2 < cnt) fetch(i + 2);
                        cp_async_commit(); // (possibly empty: keeps the group count in step)
                        dmma_phase2_acc<PB, PB, 32>(Ec, w, g, q, fa[i][2], fa[i][3], fa[i][0], fa[i][1], acc, 0);
                        double *o_next = (i + 1 < cnt) ? op[i + 1] : nullptr;
                        if (o_next != op[i]) // uniform over the CTA: end of a run of equal output pointers (or of the unit)
                        {
#pragma unroll
                            for (int jj = 0; jj < PB; ++jj)
                            {
                                const int j  = w * PB + jj;
                                const int h0 = g * 8 + 2 * q;
                                const int sg = ((g & 1) << 2) | q;
                                *reinterpret_cast<double2 *>(Ec + h0 * 32 + ((j ^ sg) << 1))       = make_double2(acc[jj][0], acc[jj][2]);
                                *reinterpret_cast<double2 *>(Ec + (h0 + 1) * 32 + ((j ^ sg) << 1)) = make_double2(acc[jj][1], acc[jj][3]);
                                acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.0;
                            }
                            __syncthreads();
                            double *obase = op[i] + (long long)tile * 64 + 32 * half;
#pragma unroll 4
                            for (int r = 0; r < HN / 2 / THREADS; ++r)
                            {
                                const int c2 = t + r * THREADS, h = c2 >> 4, ci = c2 & 15;
                                const double2 v = *reinterpret_cast<const double2 *>(Ec + h * 32 + ((ci ^ dmma_sigma(h)) << 1));
                                red_add(obase + (long long)h * 4096 + 2 * ci, v.x);
                                red_add(obase + (long long)h * 4096 + 2 * ci + 1, v.y);
                            }
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            else
            {
                // d = 5: only factor 0 is left, 8 rows (i0) x 256 columns per unit and item.  One product per 8 x 8 block:
                //   D[i0' = g][col 2q, 2q+1] += sum_k F0[g][k] Z[k][col],   A fragment F0[g][q + 4s], B fragment Z[q + 4s][c + g].
                // Buffer layout: row r has its 16-byte chunks at ci ^ ((r & 3) << 1) (rows are 2 KiB apart: without it the four
                // rows a half-warp reads would share their banks).
                constexpr int HN = 2048, CG = 32 / WARPS; // elements per buffer; 8-column groups per warp
                const int cb = u.tile;                    // columns [256 cb, 256 cb + 256) = slices h = 4 cb .. 4 cb + 3 of every tile
                const double *rb = ring + (size_t)(u.c % R) * F::CH * F::NV + (size_t)cb * 256;
                auto buf = [&](int i) { return Rg + (i % 3) * HN; };
                auto fetch = [&](int i) {
                    double *Eb        = buf(i);
                    const double *src = rb + (size_t)i * F::NV;
#pragma unroll 4
                    for (int r = 0; r < HN / 2 / THREADS; ++r)
                    {
                        const int c2 = t + r * THREADS, row = c2 >> 7, ci = c2 & 127; // chunk ci of row `row`: columns 2 ci, 2 ci + 1
                        const int h  = 4 * cb + (ci >> 5);                             // the slice those columns belong to
                        const unsigned sa = (unsigned)__cvta_generic_to_shared(Eb + row * 256 + ((ci ^ ((row & 3) << 1)) << 1));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa),
                                     "l"(src + (size_t)row * 4096 + 2 * ((ci & ~31) | ((ci & 31) ^ dmma_sigma(h)))) : "memory");
                    }
                };
                double fa[F::CH][2];
                double *op[F::CH + 1];
#pragma unroll
                for (int i = 0; i < F::CH; ++i)
                {
                    op[i] = nullptr;
                    fa[i][0] = fa[i][1] = 0.0;
                    if (i < cnt)
                    {
                        const double *p0 = A[(k0 + i) * D + 0];
                        fa[i][0] = __ldg(p0 + g + q * lda); fa[i][1] = __ldg(p0 + g + (q + 4) * lda);
                        op[i] = out[k0 + i];
                    }
                }
                op[F::CH] = nullptr;
                if (t == 0) wait_counter(doneA + u.c, (unsigned)(cnt * F::TPI));
                publish();
                __syncthreads();
                fetch(0);
                cp_async_commit();
                if (1 < cnt) fetch(1);
                cp_async_commit();
                double acc[CG][2];
#pragma unroll
                for (int j = 0; j < CG; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll
                for (int i = 0; i < F::CH; ++i)
                {
                    if (i < cnt)
                    {
                        double *Ec = buf(i);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                        __syncthreads();
                        if (i + 2 < cnt) fetch(i + 2);
                        cp_async_commit();
                        double b0[CG], b1[CG];
#pragma unroll
                        for (int j = 0; j < CG; ++j)
                        {
                            const int col = (w * CG + j) * 8 + g; // my column of this group
                            b0[j] = Ec[q * 256 + (col ^ (q << 2))];       // row q:     (q & 3) << 2 in 8-byte units
                            b1[j] = Ec[(q + 4) * 256 + (col ^ (q << 2))]; // row q + 4: same low bits
                        }
#pragma unroll
                        for (int j = 0; j < CG; ++j) dmma884v(acc[j][0], acc[j][1], fa[i][0], b0[j], acc[j][0], acc[j][1]);
#pragma unroll
                        for (int j = 0; j < CG; ++j) dmma884v(acc[j][0], acc[j][1], fa[i][1], b1[j], acc[j][0], acc[j][1]);
                        double *o_next = (i + 1 < cnt) ? op[i + 1] : nullptr;
                        if (o_next != op[i])
                        {
                            // transpose through the (now free) buffer: row g, columns c + 2q, c + 2q + 1 -> linear REDG.  A warp
                            // rewrites only the 64-column stripe it alone read, but lane by lane other elements of it
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < CG; ++j)
                            {
                                const int ci = (w * CG + j) * 4 + q;
                                *reinterpret_cast<double2 *>(Ec + g * 256 + ((ci ^ ((g & 3) << 1)) << 1)) = make_double2(acc[j][0], acc[j][1]);
                                acc[j][0] = acc[j][1] = 0.0;
                            }
                            __syncthreads();
                            double *obase = op[i] + (long long)cb * 256;
#pragma unroll 4
                            for (int r = 0; r < HN / 2 / THREADS; ++r)
                            {
                                const int c2 = t + r * THREADS, row = c2 >> 7, ci = c2 & 127;
                                const double2 v = *reinterpret_cast<const double2 *>(Ec + row * 256 + ((ci ^ ((row & 3) << 1)) << 1));
                                red_add(obase + (long long)row * 4096 + 2 * ci, v.x);
                                red_add(obase + (long long)row * 4096 + 2 * ci + 1, v.y);
                            }
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads(); // all reads of the ring slot (and of both tile buffers) are done
            if (t == 0) atomicAdd(doneB + u.c, 1u);
            pickup(s); // both slots were in use until here: the next A tile starts its way only now
        }
        else
        {
            flush_pending();
            publish();
            __syncthreads();
            pickup(s);
        }
        p_cur = p_nxt;
        u     = un;
    }
    flush_pending();
}

// The ring, the counters and the event that orders successive launches (they share the ring) belong to the device.
struct DmmaL2State
{
    double *ring = nullptr;
    unsigned *ctr = nullptr;
    size_t ctr_cap = 0; // in unsigned
    cudaEvent_t ev = nullptr;
};
// knob 12: 0 = the multi-kernel routes through `input`, 1 = this kernel, 2 = this kernel with an L2 evict-first hint on
// the input loads
inline std::atomic<int> &dmma86_l2_mode() { static std::atomic<int> v{2}; return v; }
inline std::atomic<int> &dmma86_l2_ring() { static std::atomic<int> v{6}; return v; }  // knob 13 (d = 6)
inline std::atomic<int> &dmma86_l2_lag() { static std::atomic<int> v{3}; return v; }   // knob 14 (d = 6)
inline std::atomic<int> &dmma85_l2_ring() { static std::atomic<int> v{24}; return v; } // knob 15 (d = 5)
inline std::atomic<int> &dmma85_l2_lag() { static std::atomic<int> v{12}; return v; }  // knob 16 (d = 5)

template<int D>
static cudaError_t launch_dmma8_l2(int sms, const double *const *A, int lda, double *const *in, double *const *out, int nb,
                                   cudaStream_t st, std::atomic<long long> &launches)
{
    using F = DmmaL2<D>;
    static_assert(DmmaL2<5>::RING_BYTES == DmmaL2<6>::RING_BYTES, "one ring serves both");
    static std::mutex mtx;
    static DmmaL2State states[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
    // 3 CTAs of 4 warps per SM; measured against 2 and 3 CTAs of 8 warps (1.15 / 1.43 / 1.31 ms for 976 items of d = 6): the
    // more independent CTAs, the less the tensor pipe idles at their barriers
    e = kernel_setup(kron_dmma8_l2_kernel<D, 4, 3>, F::SMEM);
    if (e != cudaSuccess) return e;
    const int nchunks  = (nb + F::CH - 1) / F::CH;
    const size_t need  = 4 + 2 * (size_t)nchunks;
    std::lock_guard<std::mutex> lk(mtx);
    DmmaL2State &S = states[dev];
    if (!S.ring)
    {
        e = cudaMalloc(&S.ring, F::RING_BYTES);
        if (e != cudaSuccess) { S.ring = nullptr; return e; }
    }
    if (!S.ev)
    {
        // first launch on this device (or the event could not be created last time): nothing to wait for
        e = cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming);
        if (e != cudaSuccess) { S.ev = nullptr; return e; }
    }
    else
    {
        // the previous launch on this device (possibly on another stream) owns the ring until it is done
        e = cudaStreamWaitEvent(st, S.ev, 0);
        if (e != cudaSuccess) return e;
    }
    if (S.ctr_cap < need)
    {
        if (S.ctr) { e = cudaFree(S.ctr); S.ctr = nullptr; S.ctr_cap = 0; if (e != cudaSuccess) return e; }
        const size_t cap = need * 2 + 1024;
        e = cudaMalloc(&S.ctr, cap * sizeof(unsigned));
        if (e != cudaSuccess) { S.ctr = nullptr; return e; }
        S.ctr_cap = cap;
    }
    e = cudaMemsetAsync(S.ctr, 0, need * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    int R   = (D == 6 ? dmma86_l2_ring() : dmma85_l2_ring()).load(std::memory_order_relaxed);
    int LAG = (D == 6 ? dmma86_l2_lag() : dmma85_l2_lag()).load(std::memory_order_relaxed);
    if (R > F::RMAX) R = F::RMAX;
    if (R < 2) R = 2;
    if (LAG >= R) LAG = R - 1;
    if (LAG < 1) LAG = 1;
    const long long total    = (long long)(nchunks + LAG) * F::BLOCK;
    const long long max_grid = (long long)sms * 3;
    const int grid           = (int)(total < max_grid ? total : max_grid);
    const int hints          = dmma86_l2_mode().load(std::memory_order_relaxed) >= 2 ? 1 : 0;
    kron_dmma8_l2_kernel<D, 4, 3><<<grid, 128, F::SMEM, st>>>(A, in, out, lda, nb, S.ring, S.ctr, nchunks, hints, R, LAG);
    launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaEventRecord(S.ev, st);
}

} // namespace kron
