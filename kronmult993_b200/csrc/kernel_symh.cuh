// kernel_symh.cuh -- symmetric single-role kernels with a HALF-WARP per item ("sym4" path): n = 4, d = 4 (both types) and
// n = 4, d = 5 in single precision.
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :54-78) for 256-element vectors.
//
// The d = 5 kernel of kernel_sym5.cuh, one size down.  A 256-element item is 16 rows (i0,i1) of 16 values (i2,i3):
// with 16 lanes per item a lane owns ONE column in phase A and ONE row in phase B, so nothing is read twice and a
// lane's run accumulator is just its 16 outputs.  A warp therefore carries TWO independent item streams (lanes 0-15
// and 16-31), each with its own 2-deep TMA ring, pointer ring and run state; the instruction stream is shared.
//   phase A (lane = column (i2,i3)): the two slow factors on the 16 values (i0,i1) in registers, written back in place;
//   phase B (lane = row (i0,i1)):    the two fast factors on the row's 16 values, the last folded into 16 accumulators.
// The rows are 128 (fp64) / 64 (fp32) bytes apart, so lane r visits the 16-byte chunks of its row at c ^ key(r)
// (key = r & 7, resp. (r >> 1) & 3): the eight lanes of a quarter-warp hit eight different bank groups, and the
// permutation is absorbed by lane-private views of the two fast factors, exactly as in kernel_sym5.cuh.
// Per item: 64 shared-memory wavefronts and 64 FP64-pipe cycles against 117 cycles of HBM time -- the first n = 4
// shape that is bound by HBM alone.  Measured (2 GB of inputs, 32 items per output): fp64 0.613 -> 0.523 ms (0.67 -> 0.79 of
// the roofline), fp32 0.837 -> 0.698 ms (0.51 -> 0.61) against the register-tile kernel of kernel_regtile.cuh.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh"
#include "kernel_wspec5.cuh"
#include "kernel_sym5.cuh"
#include <atomic>

namespace kron
{

// D = 5 (single precision only; fp64 keeps kernel_sym5.cuh -- 64 run accumulators per lane do not fit its registers): the
// same structure with rows of 64 values; a lane owns four columns in phase A (two packed FFMA2 pairs) and all four
// fastest outputs of its row in phase B, so the row is read once instead of twice.  Built, parity-tested and measured
// SLOWER than kernel_sym5.cuh (see knob 10 below), hence off by default.
template<typename T, int D_>
struct SymH
{
    static constexpr int D     = D_;
    static constexpr int N     = ipow(4, D_);
    static constexpr int RL    = N / 16;                              // values per row = columns
    static constexpr int NST   = 2;                                   // TMA ring stages per item stream
    static constexpr int STG   = (N + D * 16 + 31) / 32 * 32;         // vector + factors, k * 128 bytes for both types
    static constexpr int HALF  = NST * STG * (int)sizeof(T);          // bytes of one stream's ring
    static constexpr int SMEM  = 2 * HALF + 2 * 4 * 64 + 2 * NST * 8 + 16; // two rings, two pointer rings, barriers
    static_assert((STG * sizeof(T)) % 128 == 0, "stages start on 128-byte lines");
    static_assert(D_ == 4 || (D_ == 5 && sizeof(T) == 4), "half-warp kernels: d = 4, and d = 5 in single precision");
};

template<typename T, int D_, int MINB>
__global__ void __launch_bounds__(32, MINB)
kron_symh_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out, const int lda,
                 const int nb, const long long items_per_warp)
{
    using C = SymH<T, D_>;
    constexpr int D = D_, N = C::N, NST = C::NST, STG = C::STG, RL = C::RL;
    constexpr unsigned S = sizeof(T);
    constexpr unsigned ITEM_BYTES = N * S, FAC_BYTES = 16 * S, COL_BYTES = 4 * S;
    constexpr int E = 16 / (int)S; // elements per 16-byte chunk

    extern __shared__ __align__(128) unsigned char smem_raw[];
    int lane;
    asm volatile("mov.u32 %0, %1;" : "=r"(lane) : "r"(threadIdx.x & 31));
    const int hw = lane >> 4, hl = lane & 15;        // which half-warp (= item stream), lane within it
    const unsigned hmask = hw ? 0xffff0000u : 0x0000ffffu;

    // this warp's items, cut into two consecutive streams
    const long long K0 = (long long)blockIdx.x * items_per_warp;
    if (K0 >= nb) return;
    const int tot = (int)((K0 + items_per_warp <= nb) ? items_per_warp : (nb - K0));
    int len0 = (tot + 1) / 2;
    if (tot > 128) len0 = (len0 + 31) / 32 * 32; // stream boundaries on multiples of 32 items (ASGarD-style runs)
    if (len0 > tot) len0 = tot;
    const long long kq0 = K0 + (hw ? len0 : 0);
    const int cnt       = hw ? tot - len0 : len0;    // items of MY stream
    const int steps     = len0;                      // len0 >= tot - len0

    unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb));
    T *IN = reinterpret_cast<T *>(smem_raw + hw * C::HALF);
    const unsigned a_in   = sb + hw * C::HALF;
    const unsigned a_pr   = sb + 2 * C::HALF + hw * (4 * 64);
    const unsigned long long *PR = reinterpret_cast<const unsigned long long *>(smem_raw + 2 * C::HALF) + hw * (4 * 8);
    const unsigned b_full = sb + 2 * C::HALF + 2 * 4 * 64 + hw * NST * 8;

    if (hl == 0)
    {
        uint64_t *b = reinterpret_cast<uint64_t *>(smem_raw + 2 * C::HALF + 2 * 4 * 64) + hw * NST;
        for (int i = 0; i < NST; ++i) mbar_init(b + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // Vector and factors of one item of MY stream -> ring stage st, completion on the stream's mbarrier.  Executed by
    // the 16 lanes of a stream together (votes and shuffles carry the half-warp mask).
    auto stage_item = [&](int st, const T *ip, const T *ap) {
        const bool lda4  = (lda == 4);
        const bool lda16 = ((lda * (int)S) % 16 == 0);
        const unsigned dst = a_in + st * (STG * S), fdst = dst + N * S;
        const unsigned bar = b_full + 8 * st;
        const bool vtma    = aligned16(ip);
        const T *ap0 = reinterpret_cast<const T *>(__shfl_sync(hmask, reinterpret_cast<unsigned long long>(ap), hw * 16));
        const bool contig = lda4 && aligned16(ap0) && __all_sync(hmask, hl >= D || ap == ap0 + hl * 16);
        if (!vtma)
        {
#pragma unroll 8
            for (int h = 0; h < N / 16; ++h) cp_async_elem_a<T>(dst + (h * 16 + hl) * S, ip + h * 16 + hl);
        }
        if (contig)
        {
            if (hl == 0)
            {
                mbar_expect_tx_a(bar, (vtma ? ITEM_BYTES : 0u) + D * FAC_BYTES);
                if (vtma) tma_load_a(dst, ip, ITEM_BYTES, bar);
                tma_load_a(fdst, ap0, D * FAC_BYTES, bar);
            }
        }
        else
        {
            const bool a16  = __all_sync(hmask, hl >= D || aligned16(ap));
            const bool ftma = a16 && (lda4 || lda16);
            if (hl == 0)
            {
                const unsigned bytes = (vtma ? ITEM_BYTES : 0u) + (ftma ? D * FAC_BYTES : 0u);
                if (bytes) mbar_expect_tx_a(bar, bytes); else mbar_arrive_a(bar);
                if (vtma) tma_load_a(dst, ip, ITEM_BYTES, bar);
            }
            __syncwarp(hmask);
            if (ftma && lda4) { if (hl < D) tma_load_a(fdst + hl * FAC_BYTES, ap, FAC_BYTES, bar); }
            else if (ftma)
            {
                // one copy per column: column c & 3 of factor c >> 2
#pragma unroll
                for (int i = 0; i < (4 * D + 15) / 16; ++i)
                {
                    const int c  = hl + 16 * i;
                    const T *apj = reinterpret_cast<const T *>(
                        __shfl_sync(hmask, reinterpret_cast<unsigned long long>(ap), hw * 16 + (c >> 2) % D));
                    if (c < 4 * D) tma_load_a(fdst + c * COL_BYTES, apj + (long long)(c & 3) * lda, COL_BYTES, bar);
                }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < D; ++i)
                {
                    const int e  = hl + 16 * i; // element e = factor e/16, column (e%16)/4, row e%4
                    const T *apj = reinterpret_cast<const T *>(
                        __shfl_sync(hmask, reinterpret_cast<unsigned long long>(ap), hw * 16 + (e >> 4)));
                    cp_async_elem_a<T>(fdst + e * S, apj + (e & 3) + (long long)((e >> 2) & 3) * lda);
                }
            }
        }
    };
    // lanes 0..D-1 of a stream fetch the item's factor pointers, lane 5 its input pointer, lane 6 its output pointer
    auto fetch_ptrs = [&](int s) {
        if (s < cnt && hl < 7 && (hl < D || hl > 4))
        {
            const long long k = kq0 + s;
            const void *src = hl < D ? static_cast<const void *>(A + k * D + hl)
                            : hl == 5 ? static_cast<const void *>(in + k) : static_cast<const void *>(out + k);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a_pr + (s & 3) * 64 + hl * 8), "l"(src) : "memory");
        }
    };
    auto pr_in  = [&](int s) -> const T * { return reinterpret_cast<const T *>(PR[(s & 3) * 8 + 5]); };
    auto pr_ap  = [&](int s) -> const T * { return reinterpret_cast<const T *>(PR[(s & 3) * 8 + (hl < D ? hl : 0)]); };
    auto pr_out = [&](int s) -> T * { return reinterpret_cast<T *>(PR[(s & 3) * 8 + 6]); };

    // lane-private view of the row-wise phase: my row r = hl, chunk c of it sits at c ^ key
    // (rows of 128 bytes and more: key = r & 7; 64-byte rows: two per line, key = (r >> 1) & 3)
    const int key    = (RL * S >= 128) ? (hl & 7) : ((hl >> 1) & 3);
    const int mask_e = key * E;                                    // the same permutation on element indices
    // ... as seen by the row's indices: d = 4 (i2, i3); d = 5 (i2, i3, i4), whose fastest index i4 fills a chunk
    [[maybe_unused]] const int km = mask_e & 3, i2m = (mask_e >> 2) & 3;
    [[maybe_unused]] const int i3m5 = (mask_e >> 2) & 3, i2m5 = (mask_e >> 4) & 3;

    T acc[RL]; // d = 4: acc[i2' * 4 + i3']; d = 5: acc[(i2' * 4 + i3') * 4 + i4'] of my row
#pragma unroll
    for (int i = 0; i < RL; ++i) acc[i] = T(0);

    fetch_ptrs(0); fetch_ptrs(1); fetch_ptrs(2);
    cp_async_commit();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (cnt > 0) stage_item(0, pr_in(0), pr_ap(0));
    cp_async_commit();

    for (int s = 0; s < steps; ++s)
    {
        const int st    = s & 1;
        const bool live = s < cnt;
        // every cp.async group committed in earlier steps is complete (pointers of items s+1, s+2; element-wise copies
        // of item s) and visible to the whole warp; every read of stage st^1 and ring slot (s+3)&3 is done
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (s + 1 < cnt) stage_item(st ^ 1, pr_in(s + 1), pr_ap(s + 1));
        fetch_ptrs(s + 3);
        cp_async_commit();
        T *stage    = IN + st * STG;
        const T *Mq = stage + N; // factors 0..3, column-major 4x4 blocks
        if (live)
        {
            mbar_wait_a(b_full + 8 * st, (unsigned)(s >> 1) & 1u);
            // -------------------------------------------------------- phase A: factors 0 and 1 on my column(s), in place
            if constexpr (D == 4)
            {
                T x[16];
                T *col = stage + hl;
#pragma unroll
                for (int h = 0; h < 16; ++h) x[h] = col[h * 16];
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
                for (int h = 0; h < 16; ++h) col[h * 16] = x[h];
            }
            else
            {
                // four adjacent columns = two packed pairs (FFMA2)
                using P = typename V2<T>::type;
                P xa[16], xb[16];
                T *col = stage + 4 * hl;
#pragma unroll
                for (int h = 0; h < 16; ++h)
                {
                    const float4 v = *reinterpret_cast<const float4 *>(col + h * RL);
                    xa[h] = make_float2(v.x, v.y); xb[h] = make_float2(v.z, v.w);
                }
                {
                    T m[16];
                    lds16<T>(Mq + 1 * 16, m);
                    tile16_apply_cm2<T, 1>(xa, m);
                    tile16_apply_cm2<T, 1>(xb, m);
                }
                {
                    T m[16];
                    lds16<T>(Mq + 0 * 16, m);
                    tile16_apply_cm2<T, 4>(xa, m);
                    tile16_apply_cm2<T, 4>(xb, m);
                }
#pragma unroll
                for (int h = 0; h < 16; ++h) *reinterpret_cast<float4 *>(col + h * RL) = make_float4(xa[h].x, xa[h].y, xb[h].x, xb[h].y);
            }
        }
        __syncwarp();
        if (live)
        {
            // -------------------------------------------------------- phase B: the fast factors on my row + run sums
            if constexpr (D == 4)
            {
                T x[16], g[16];
                const T *rowp = stage + hl * 16;
                if constexpr (S == 8)
                {
    #pragma unroll
                    for (int c = 0; c < 8; ++c)
                    {
                        const double2 v = *reinterpret_cast<const double2 *>(rowp + ((c * 2) ^ mask_e));
                        x[2 * c] = v.x; x[2 * c + 1] = v.y;
                    }
                }
                else
                {
    #pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        const float4 v = *reinterpret_cast<const float4 *>(rowp + ((c * 4) ^ mask_e));
                        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                    }
                }
                // g[ks*4 + i] = F3(i, ks ^ km): the columns of the fastest factor in my visiting order
    #pragma unroll
                for (int ks = 0; ks < 4; ++ks)
    #pragma unroll
                    for (int i = 0; i < 4; ++i) g[ks * 4 + i] = Mq[3 * 16 + (ks ^ km) * 4 + i];
                tile16_apply_cm<T, 1>(x, g);
                // acc[i*4 + m] += sum_ks F2(i, ks ^ i2m) x[ks*4 + m]
    #pragma unroll
                for (int ks = 0; ks < 4; ++ks)
    #pragma unroll
                    for (int i = 0; i < 4; ++i) g[ks * 4 + i] = Mq[2 * 16 + (ks ^ i2m) * 4 + i];
    #pragma unroll
                for (int ks = 0; ks < 4; ++ks)
    #pragma unroll
                    for (int i = 0; i < 4; ++i)
    #pragma unroll
                        for (int m = 0; m < 4; ++m) acc[i * 4 + m] = fma(x[ks * 4 + m], g[ks * 4 + i], acc[i * 4 + m]);
            }
            else
            {
                // d = 5: slices j = (i2 slot) of 16 values; y = F4 x (all four i4'), z += F3 column * y, acc += F2 column * z
                using P = typename V2<T>::type;
                const T *rowp = stage + hl * RL;
                P g4a[4], g4b[4]; // (F4(0,k), F4(1,k)), (F4(2,k), F4(3,k)): i4 is not permuted (a whole chunk)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const float4 v = *reinterpret_cast<const float4 *>(Mq + 4 * 16 + k * 4);
                    g4a[k] = make_float2(v.x, v.y); g4b[k] = make_float2(v.z, v.w);
                }
                T g3[16]; // g3[sl*4 + i] = F3(i, sl ^ i3m5)
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                {
                    const float4 v = *reinterpret_cast<const float4 *>(Mq + 3 * 16 + (sl ^ i3m5) * 4);
                    g3[sl * 4 + 0] = v.x; g3[sl * 4 + 1] = v.y; g3[sl * 4 + 2] = v.z; g3[sl * 4 + 3] = v.w;
                }
                P *acca = reinterpret_cast<P *>(acc); // acc[(i2'*4+i3')*4 + i4'] viewed as pairs (i4' = 0,1) and (2,3)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    P za[4], zb[4];
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl)
                    {
                        const float4 v = *reinterpret_cast<const float4 *>(rowp + (((j * 4 + sl) * 4) ^ mask_e));
                        P ya = pmul(g4a[0], v.x), yb = pmul(g4b[0], v.x);
                        ya = pfma(g4a[1], v.y, ya); yb = pfma(g4b[1], v.y, yb);
                        ya = pfma(g4a[2], v.z, ya); yb = pfma(g4b[2], v.z, yb);
                        ya = pfma(g4a[3], v.w, ya); yb = pfma(g4b[3], v.w, yb);
                        if (sl == 0)
                        {
#pragma unroll
                            for (int i = 0; i < 4; ++i) { za[i] = pmul(ya, g3[i]); zb[i] = pmul(yb, g3[i]); }
                        }
                        else
                        {
#pragma unroll
                            for (int i = 0; i < 4; ++i) { za[i] = pfma(ya, g3[sl * 4 + i], za[i]); zb[i] = pfma(yb, g3[sl * 4 + i], zb[i]); }
                        }
                    }
                    const float4 f2 = *reinterpret_cast<const float4 *>(Mq + 2 * 16 + (j ^ i2m5) * 4); // F2(:, j ^ i2m5)
                    const T f2v[4] = {f2.x, f2.y, f2.z, f2.w};
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                        {
                            acca[(i2 * 4 + m) * 2]     = pfma(za[m], f2v[i2], acca[(i2 * 4 + m) * 2]);
                            acca[(i2 * 4 + m) * 2 + 1] = pfma(zb[m], f2v[i2], acca[(i2 * 4 + m) * 2 + 1]);
                        }
                }
            }
        }
        // ---- flush when my stream's run of equal output pointers ends here (the two streams decide independently)
        T *o_cur         = live ? pr_out(s) : nullptr;
        const bool flush = live && (s + 1 >= cnt || pr_out(s + 1) != o_cur);
        if (__any_sync(0xffffffffu, flush))
        {
            __syncwarp(); // every lane is done reading its row
            if (flush)
            {
                T *rowp = stage + hl * RL;
                if constexpr (S == 8)
                {
#pragma unroll
                    for (int c = 0; c < RL / 2; ++c) *reinterpret_cast<double2 *>(rowp + ((c * 2) ^ mask_e)) = make_double2(acc[2 * c], acc[2 * c + 1]);
                }
                else
                {
#pragma unroll
                    for (int c = 0; c < RL / 4; ++c)
                        *reinterpret_cast<float4 *>(rowp + ((c * 4) ^ mask_e)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                }
#pragma unroll
                for (int i = 0; i < RL; ++i) acc[i] = T(0);
            }
            __syncwarp();
            if (flush)
            {
#pragma unroll 4
                for (int h = 0; h < 16; ++h)
                {
                    const int mh = ((RL * S >= 128) ? (h & 7) : ((h >> 1) & 3)) * E;
#pragma unroll
                    for (int q = 0; q < RL / 16; ++q) red_add(o_cur + h * RL + q * 16 + hl, stage[h * RL + ((q * 16 + hl) ^ mh)]);
                }
                fence_proxy_async(); // the next bulk copy into this stage follows generic-proxy writes
            }
        }
    }
}

template<typename T, int D, int MINB>
static cudaError_t launch_symh(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb, cudaStream_t st,
                               std::atomic<long long> &launches)
{
    using C  = SymH<T, D>;
    auto kfn = kron_symh_kernel<T, D, MINB>;
    int ctas_per_sm = 0;
    cudaError_t e = kernel_setup(kfn, 32, C::SMEM, ctas_per_sm);
    if (e != cudaSuccess) return e;
    const long long warps = (long long)sms * ctas_per_sm;
    long long ipw = ((long long)nb + warps - 1) / warps; // items per warp (two streams)
    if (ipw > 128) ipw = (ipw + 63) / 64 * 64;
    const long long grid = ((long long)nb + ipw - 1) / ipw;
    kfn<<<(int)grid, 32, C::SMEM, st>>>(A, in, out, lda, nb, ipw);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// knob 10: single-precision n = 4, d = 5 on the half-warp kernel (1: 8 CTAs per SM, 2: 12).  Default 0: measured on
// B200 (C5-f32, 8 Mi items) 8.88-9.15 ms against 8.25-8.77 ms for kernel_sym5.cuh -- the fp32 kernel is bound by issue
// slots and latency, not by the shared-memory wavefronts this layout saves.
static std::atomic<int> g_symh_f32_d5{0};

// cudaErrorNotSupported when (T, n, d) is outside the family
template<typename T>
static cudaError_t run_sym4(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                            cudaStream_t st, std::atomic<long long> &launches, const char *&last_path, bool forced = false)
{
    if (n != 4) return cudaErrorNotSupported;
    if (d == 4)
    {
        cudaError_t e = (sizeof(T) == 8) ? launch_symh<T, 4, 16>(sms, A, lda, in, out, nb, st, launches)
                                         : launch_symh<T, 4, 20>(sms, A, lda, in, out, nb, st, launches);
        last_path = "sym4";
        return e;
    }
    if constexpr (sizeof(T) == 4)
    {
        if (d == 5 && (forced || g_symh_f32_d5.load(std::memory_order_relaxed)))
        {
            // 8 single-warp CTAs per SM (no register cap, no spills) or 12 (168 registers: measured below)
            cudaError_t e = (g_symh_f32_d5.load(std::memory_order_relaxed) == 2)
                                ? launch_symh<T, 5, 12>(sms, A, lda, in, out, nb, st, launches)
                                : launch_symh<T, 5, 8>(sms, A, lda, in, out, nb, st, launches);
            last_path = "sym4";
            return e;
        }
    }
    return cudaErrorNotSupported;
}

} // namespace kron
