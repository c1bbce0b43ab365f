// kernel_rows2.cuh -- d = 2, n = 5 .. 10: one LANE per fibre, floor(32 / n) (n odd: floor(32 / (n + 1))) items side by side in a warp ("rows2").
//
// Replaces cuda_kronmult_batchelement + cuda_kronmult (kronmult_gpu/kronmult.cu:139-167, :95-130) for the items of
// 25 .. 100 elements with two factors.  These shapes are HBM-bound with room to spare (n = 10, fp64: 2000 FMAs against
// 2.4 KB of compulsory traffic per item, a third of the FP64 pipe's time), but the pair-tile kernel they ran on spends
// 503 warp instructions per 800-byte item, a quarter of them DFMA (profiles/ncu_pairtile_small_r02.md: run-time tile
// addressing), and a thread-per-item kernel needs 2 n^2 registers.  Here an item is Out = M1 . In . M0^T with n x n
// column-major matrices (factor d-1 = M1 acts on the fast index, kronmult.cu:107-118), and a warp works on
// IPW items at once (a slot takes n lanes, n + 1 for odd n), lane (s, j) = fibre j of item slot s:
//   stage    In, M1, M0 of every slot arrive in shared memory with cp.async (16-byte chunks when the data is compact and
//            aligned, element copies otherwise: lda > n, windows into big matrices, unaligned vectors), one round ahead;
//   phase 1  lane j reads column j of In (n contiguous values, vector loads), multiplies by M1 -- whose columns are read
//            as slot-uniform vector loads (two wavefronts for the whole warp) -- and writes the result back in place;
//   phase 0  lane j reads row j of the intermediate (stride n: consecutive lanes, consecutive words), multiplies by M0
//            and adds onto its n run accumulators.
// A slot walks CONSECUTIVE batch items, so runs of equal output pointers are summed in registers; at the end of a run
// lane j adds its n values with REDG -- for every r the lanes of a slot cover n consecutive elements.
// Everything is compile-time: per round of IPW items the warp issues 2 n^2 DFMA, ~n^2 + 2n shared loads, n / 2 stores and
// ~3 n / 2 copies (n = 10: ~130 warp instructions per item).
#pragma once
#include "common.cuh"
#include <type_traits>
#include <utility>

namespace kron
{

inline std::atomic<int> &rows2_enabled() { static std::atomic<int> v{1}; return v; } // knob 17
inline std::atomic<int> &rows2_variant() { static std::atomic<int> v{-1}; return v; } // knob 18: 0 / 1 = 2 / 3 stages, -1 = per shape

// which (T, n) the kernel takes for d = 2 (knob 17: 0 = off, 1 = the shapes where it measured faster, 2 = all it is built for)
template<typename T>
static bool rows2_takes(int n, int d)
{
    const int mode = rows2_enabled().load(std::memory_order_relaxed);
    if (mode == 0 || d != 2 || n < 5 || n > 10) return false;
    if (mode == 2) return true;
    // measured against the kernels it replaces (tools/rows2_session.sh, profiles/rows2_r02.md; fraction of the roofline,
    // rows2 / before):  fp64 n = 5 0.62 / 0.64 (tiny), 6 0.80 / 0.69 (tiny), 7 0.72 / 0.79 (dmma), 8 0.74 / 0.92 (dmma),
    //                   9 0.76 / 0.32 (pairtile), 10 0.85 / 0.36 (pairtile)
    //                   fp32 n = 5 0.42 / 0.55, 6 0.69 / 0.76, 7 0.49 / 0.42, 8 0.85 / 0.66, 9 0.45 / 0.31, 10 0.75 / 0.51 (all tiny)
    if (sizeof(T) == 8) return n == 6 || n == 9 || n == 10;
    return n >= 7;
}

// defined in rows2.cu (own translation unit: compiled in parallel); cudaErrorNotSupported outside n = 5 .. 10
template<typename T>
cudaError_t run_rows2(int sms, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb, cudaStream_t st,
                      std::atomic<long long> &launches);

#ifdef KRON_ROWS2_DEFINE

template<int I, int E, typename F>
__device__ __forceinline__ void rows2_static_for(F &&f)
{
    if constexpr (I < E)
    {
        f(std::integral_constant<int, I>{});
        rows2_static_for<I + 1, E>(f);
    }
}

// COUNT consecutive values starting OFF elements behind a 16-byte aligned shared-memory address: scalar loads up to
// the next 16-byte boundary, 128-bit loads, scalar tail (OFF is a compile-time constant, so the split is too)
template<typename T, int COUNT, int OFF>
__device__ __forceinline__ void rows2_lds_run(const T *base16, T (&v)[COUNT])
{
    constexpr int VEC   = 16 / (int)sizeof(T);
    constexpr int LEAD0 = (VEC - OFF % VEC) % VEC;
    constexpr int LEAD  = LEAD0 < COUNT ? LEAD0 : COUNT;
    constexpr int NV    = (COUNT - LEAD) / VEC;
#pragma unroll
    for (int e = 0; e < LEAD; ++e) v[e] = base16[OFF + e];
#pragma unroll
    for (int i = 0; i < NV; ++i)
    {
        const int4 w = *reinterpret_cast<const int4 *>(base16 + OFF + LEAD + i * VEC);
        if constexpr (sizeof(T) == 8)
        {
            v[LEAD + 2 * i]     = __hiloint2double(w.y, w.x);
            v[LEAD + 2 * i + 1] = __hiloint2double(w.w, w.z);
        }
        else
        {
            v[LEAD + 4 * i]     = __int_as_float(w.x);
            v[LEAD + 4 * i + 1] = __int_as_float(w.y);
            v[LEAD + 4 * i + 2] = __int_as_float(w.z);
            v[LEAD + 4 * i + 3] = __int_as_float(w.w);
        }
    }
#pragma unroll
    for (int e = LEAD + NV * VEC; e < COUNT; ++e) v[e] = base16[OFF + e];
}

// COUNT consecutive values at a lane-dependent address p = base16 + j * COUNT: the widest loads the pitch guarantees
template<typename T, int COUNT>
__device__ __forceinline__ void rows2_lds_row(const T *p, T (&v)[COUNT])
{
    constexpr int RB = COUNT * (int)sizeof(T);
    if constexpr (RB % 16 == 0) rows2_lds_run<T, COUNT, 0>(p, v);
    else if constexpr (sizeof(T) == 4 && RB % 8 == 0)
    {
#pragma unroll
        for (int i = 0; i < COUNT / 2; ++i)
        {
            const float2 w = *reinterpret_cast<const float2 *>(p + 2 * i);
            v[2 * i] = w.x; v[2 * i + 1] = w.y;
        }
    }
    else
    {
#pragma unroll
        for (int e = 0; e < COUNT; ++e) v[e] = p[e];
    }
}
template<typename T, int COUNT>
__device__ __forceinline__ void rows2_sts_row(T *p, const T (&v)[COUNT])
{
    constexpr int RB = COUNT * (int)sizeof(T);
    if constexpr (RB % 16 == 0)
    {
        constexpr int VEC = 16 / (int)sizeof(T);
#pragma unroll
        for (int i = 0; i < COUNT / VEC; ++i)
        {
            int4 w;
            if constexpr (sizeof(T) == 8)
            {
                w.x = __double2loint(v[2 * i]); w.y = __double2hiint(v[2 * i]);
                w.z = __double2loint(v[2 * i + 1]); w.w = __double2hiint(v[2 * i + 1]);
            }
            else
            {
                w.x = __float_as_int(v[4 * i]); w.y = __float_as_int(v[4 * i + 1]);
                w.z = __float_as_int(v[4 * i + 2]); w.w = __float_as_int(v[4 * i + 3]);
            }
            *reinterpret_cast<int4 *>(p + i * VEC) = w;
        }
    }
    else if constexpr (sizeof(T) == 4 && RB % 8 == 0)
    {
#pragma unroll
        for (int i = 0; i < COUNT / 2; ++i) *reinterpret_cast<float2 *>(p + 2 * i) = make_float2(v[2 * i], v[2 * i + 1]);
    }
    else
    {
#pragma unroll
        for (int e = 0; e < COUNT; ++e) p[e] = v[e];
    }
}

template<typename T, int NN, int STAGES_, int WARPS_>
struct Rows2Cfg
{
    // lanes per slot: n rounded up to even.  A 128-bit shared load whose even / odd lane PAIRS disagree on the address takes
    // the slow path (tools/probes/lds_probe.cu on B200: groups of 9 or 5 lanes 2.1 cycles per warp load, groups of 10,
    // 8, 4 or 2 lanes 1.3 -- like a warp-wide broadcast), so a slot never starts on an odd lane.
    static constexpr int LPS    = NN + (NN & 1);
    static constexpr int IPW    = 32 / LPS;                      // item slots per warp
    static constexpr int NSQ    = NN * NN;
    static constexpr int VEC    = 16 / (int)sizeof(T);
    static constexpr int MAT    = (NSQ + VEC - 1) / VEC * VEC;   // elements per staged matrix (a multiple of 16 bytes)
    // fp64 with odd n: n^2 * 8 is not a multiple of 16, so in a compact batch every other vector / factor starts 8 bytes
    // off a 16-byte boundary.  Such a matrix is staged SHIFTED by one element (its image starts at element 1 of its
    // region, which has the spare element) so that source and destination agree modulo 16 and 16-byte copies still
    // work; the three shift bits of a slot's round travel in a meta word behind the matrices.
    static constexpr bool SHIFT = sizeof(T) == 8 && (NN % 2) == 1;
    static constexpr int META   = SHIFT ? VEC : 0;
    static constexpr int SLOT0  = 3 * MAT + META;                // In, M1, M0 (, meta)
    // an odd number of 16-byte units per slot: the slot-uniform 128-bit loads of the IPW slots then fall into IPW
    // different bank groups (ncu: 2.1 wavefronts per load for three slots of ten lanes, profiles/rows2_r02.md)
    static constexpr int SLOT   = ((SLOT0 / VEC) % 2 == 0) ? SLOT0 + VEC : SLOT0;
    static constexpr int STAGE  = IPW * SLOT;
    static constexpr int STAGES = STAGES_;                       // 2 or 3: rounds in flight + the one being computed
    static constexpr int WARPS  = WARPS_;
    static constexpr int SMEM   = WARPS * STAGES * STAGE * (int)sizeof(T);
    static constexpr int MINB0  = (220 * 1024) / (SMEM + 1024);
    static constexpr int MINBC  = 512 / (32 * WARPS);            // 128 registers per thread are enough (ptxas: 108 .. 126)
    static constexpr int MINB   = MINB0 > MINBC ? MINBC : MINB0;
    static constexpr int NCH    = NSQ / VEC;                     // whole 16-byte chunks of a compact matrix
    static constexpr int TAIL   = NSQ - NCH * VEC;               // elements behind them (< VEC <= 4 < NN)
};

template<typename T>
__device__ __forceinline__ void rows2_cp_elem(unsigned dst, const T *src)
{
    if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// lane j of a slot copies its share of one n x n column-major matrix (leading dimension ld) into the compact staged copy.
// Returns the shift of the image in elements (0, or 1 with SHIFT_OK for a compact source 8 bytes off a 16-byte boundary).
template<typename T, int NN, bool SHIFT_OK>
__device__ __forceinline__ int rows2_copy_matrix(T *dst, const T *__restrict__ src, const int ld, const int j)
{
    constexpr int VEC = 16 / (int)sizeof(T), NSQ = NN * NN, NCH = NSQ / VEC, TAIL = NSQ - NCH * VEC;
    constexpr int RB = NN * (int)sizeof(T); // bytes per column
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    if (ld == NN && (SHIFT_OK || aligned16(src)))
    {
        // compact: 16-byte chunks.  SHIFT_OK (fp64, n odd, NCH = (n^2 - 1) / 2): shift 0 = chunks + last element,
        // shift 1 = first element + chunks, the image one element further into the region.
        // (Measured instead, both dropped -- profiles/rows2_r02.md: one bulk copy (1-D TMA) per matrix by the slot's lane 0
        // -- ptxas serialises the per-lane UBLKCP with an ELECT loop, nine per round: n = 6 +8 %, n = 10 and fp32 n = 8
        // -6..15 %; and the whole warp copying matrix after matrix with the pointers shuffled from the slot's first lane
        // -- fewer shared-memory wavefronts, but 9 serial copy blocks per round: -13..35 % everywhere.  cp.async.ca
        // instead of .cg for the chunks: -10..35 %.)
        int sh = 0;
        if constexpr (SHIFT_OK) sh = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
#pragma unroll
        for (int q0 = 0; q0 < NCH; q0 += NN)
        {
            const int q = q0 + j;
            if (q0 + NN <= NCH || q < NCH)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + (q + sh) * 16), "l"(src + q * VEC + sh) : "memory");
        }
        if constexpr (SHIFT_OK)
        {
            const int e = sh ? 0 : NSQ - 1;
            if (j == 0) rows2_cp_elem<T>(sa + (unsigned)(e + sh) * 8u, src + e);
        }
        else if constexpr (TAIL > 0)
        {
            if (j < TAIL) rows2_cp_elem<T>(sa + (NCH * VEC + j) * (unsigned)sizeof(T), src + NCH * VEC + j);
        }
        return sh;
    }
    if constexpr (RB % 16 == 0)
    {
        if (aligned16(src) && (((long long)ld * (long long)sizeof(T)) & 15) == 0)
        {
            // every column starts on a 16-byte boundary (ASGarD: windows into big coefficient matrices, lda = 64 n)
            constexpr int CPC = RB / 16; // chunks per column
#pragma unroll
            for (int q0 = 0; q0 < NCH; q0 += NN)
            {
                const int q = q0 + j, c = q / CPC, r = q - c * CPC;
                if (q0 + NN <= NCH || q < NCH)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + q * 16), "l"(src + (long long)c * ld + r * VEC) : "memory");
            }
            return 0;
        }
    }
    // row j of every column: for each column the lanes of a slot read n consecutive elements
#pragma unroll
    for (int c = 0; c < NN; ++c) rows2_cp_elem<T>(sa + (c * NN + j) * (unsigned)sizeof(T), src + j + (long long)c * ld);
    return 0;
}

// y += M x for the staged factor whose region starts at `reg` (16-byte aligned) and whose image is shifted by sh
// elements.  Column k of M is read with slot-uniform vector loads whose split into 64- and 128-bit pieces depends on
// the parity of its first element, k n + sh.  For a shifted image (n odd) the columns are taken in the order
// 1, 0, 3, 2, ... instead: column k ^ 1 of a shifted image has the alignment of column k of an unshifted one, so one
// instruction sequence serves both kinds of slot in a warp, with the operand x[k ^ 1] selected once per phase.  The
// unpaired last column is read with 64-bit loads.
template<typename T, int NN, bool SHIFT>
__device__ __forceinline__ void rows2_mul(const T *reg, const int sh, const T (&x)[NN], T (&y)[NN])
{
    if constexpr (!SHIFT)
    {
        rows2_static_for<0, NN>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            T m[NN];
            rows2_lds_run<T, NN, k * NN>(reg, m);
#pragma unroll
            for (int c = 0; c < NN; ++c) y[c] = m[c] * x[k] + y[c];
        });
    }
    else
    {
        T xs[NN];
#pragma unroll
        for (int k = 0; k < NN - 1; ++k) xs[k] = sh ? x[k ^ 1] : x[k];
        const T *even = reg + (sh ? 1 + NN : 0); // steps with even k: column k (+1 when shifted)
        const T *odd  = reg + (sh ? 1 - NN : 0); // steps with odd k: column k (-1 when shifted)
        rows2_static_for<0, NN - 1>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            T m[NN];
            rows2_lds_run<T, NN, k * NN>((k % 2 == 0) ? even : odd, m);
#pragma unroll
            for (int c = 0; c < NN; ++c) y[c] = m[c] * xs[k] + y[c];
        });
        const T *last = reg + sh + (NN - 1) * NN;
#pragma unroll
        for (int c = 0; c < NN; ++c) y[c] = last[c] * x[NN - 1] + y[c];
    }
}

template<typename T, int NN, int STAGES, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, Rows2Cfg<T, NN, STAGES, WARPS>::MINB)
kron_rows2_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out, const int lda,
                  const int nb, const int L)
{
    using C = Rows2Cfg<T, NN, STAGES, WARPS>;
    static_assert(STAGES == 2 || STAGES == 3, "the pointer pipeline below is written for 2 or 3 stages");
    extern __shared__ __align__(16) unsigned char rows2_smem[];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int s = lane / C::LPS, j = lane - s * C::LPS;
    const bool lane_on = s < C::IPW && j < NN; // the other lanes of a warp have no fibre
    T *const wbase = reinterpret_cast<T *>(rows2_smem) + (size_t)w * (C::STAGES * C::STAGE) + (lane_on ? s : 0) * C::SLOT;

    // slot g of the grid owns the L consecutive items from g * L on
    const long long g     = ((long long)blockIdx.x * C::WARPS + w) * C::IPW + s;
    const long long first = g * L;
    int cnt = 0;
    if (lane_on && first < nb) cnt = (int)(nb - first < L ? nb - first : L);
    const int rounds = __shfl_sync(0xffffffffu, cnt, 0); // slot 0 of a warp has the most items
    if (rounds == 0) return;

    struct Ptrs { const T *ip; T *op; const T *a0; const T *a1; };
    auto load_ptrs = [&](int i, Ptrs &p) {
        if (i < cnt)
        {
            const long long k = first + i;
            p.ip = in[k]; p.op = out[k];
            p.a0 = A[2 * k]; p.a1 = A[2 * k + 1];
        }
        else { p.ip = nullptr; p.op = nullptr; p.a0 = nullptr; p.a1 = nullptr; }
    };
    auto issue = [&](const Ptrs &p, int stage) {
        if (p.ip)
        {
            T *X = wbase + stage * C::STAGE;
            const int sx = rows2_copy_matrix<T, NN, C::SHIFT>(X, p.ip, NN, j);
            const int s1 = rows2_copy_matrix<T, NN, C::SHIFT>(X + C::MAT, p.a1, lda, j);
            const int s0 = rows2_copy_matrix<T, NN, C::SHIFT>(X + 2 * C::MAT, p.a0, lda, j);
            if constexpr (C::SHIFT)
            {
                // ordered before the round's reads by the __syncwarp that follows its wait_group; the previous use of
                // this stage ended with a __syncwarp as well
                if (j == 0) *reinterpret_cast<int *>(X + 3 * C::MAT) = sx | (s1 << 1) | (s0 << 2);
            }
        }
        // always a group (possibly empty): the consumer's wait_group counts groups
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // round i is computed while rounds i+1 .. i+STAGES-1 are in flight; the pointers of round i+STAGES are loaded meanwhile
    Ptrs p_issue, p_load;
    T *op_a, *op_b; // output pointers of rounds i and i+1
    {
        Ptrs p0;
        load_ptrs(0, p0);
        issue(p0, 0);
        op_a = p0.op;
        load_ptrs(1, p_issue);
        op_b = p_issue.op;
        if constexpr (STAGES == 3)
        {
            issue(p_issue, 1);
            load_ptrs(2, p_issue);
        }
    }

    T acc[NN];
#pragma unroll
    for (int r = 0; r < NN; ++r) acc[r] = T(0);

    int st_cur = 0, st_iss = STAGES - 1;
    for (int i = 0; i < rounds; ++i)
    {
        load_ptrs(i + STAGES, p_load);
        issue(p_issue, st_iss);
        if constexpr (STAGES == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();

        T *X = wbase + st_cur * C::STAGE;
        int sx = 0, s1 = 0, s0 = 0;
        if constexpr (C::SHIFT)
        {
            const int meta = *reinterpret_cast<const int *>(X + 3 * C::MAT);
            sx = meta & 1; s1 = (meta >> 1) & 1; s0 = (meta >> 2) & 1;
        }
        T *Xp = X + sx; // image of In (and of the intermediate)
        if (lane_on) // (the spare lanes would read slot 0's fibres while their owners rewrite them)
        {
            // phase 1: column j of In times M1, in place
            T x[NN], y[NN];
            rows2_lds_row<T, NN>(Xp + j * NN, x);
#pragma unroll
            for (int c = 0; c < NN; ++c) y[c] = T(0);
            rows2_mul<T, NN, C::SHIFT>(X + C::MAT, s1, x, y);
            rows2_sts_row<T, NN>(Xp + j * NN, y); // own fibre: no other lane reads or writes it in this phase
        }
        __syncwarp();
        if (lane_on)
        {
            // phase 0: row j of the intermediate times M0, onto the run accumulators
            T z[NN];
#pragma unroll
            for (int k = 0; k < NN; ++k) z[k] = Xp[k * NN + j];
            rows2_mul<T, NN, C::SHIFT>(X + 2 * C::MAT, s0, z, acc);
        }
        if (op_b != op_a) // end of a run of equal output pointers (op_b is null behind the slot's last item)
        {
            if (i < cnt)
            {
#pragma unroll
                for (int r = 0; r < NN; ++r) red_add(op_a + j + NN * r, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < NN; ++r) acc[r] = T(0);
        }
        __syncwarp(); // every lane is done with this stage before a later round is copied into it
        op_a = op_b;
        op_b = (STAGES == 2) ? p_load.op : p_issue.op; // round i+2
        p_issue = p_load;
        st_cur = (st_cur + 1 == STAGES) ? 0 : st_cur + 1;
        st_iss = (st_iss + 1 == STAGES) ? 0 : st_iss + 1;
    }
}

template<typename T, int NN, int STAGES, int WARPS>
static cudaError_t launch_rows2v(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb, cudaStream_t st,
                                 std::atomic<long long> &launches)
{
    using C = Rows2Cfg<T, NN, STAGES, WARPS>;
    int occ = 1;
    cudaError_t e = kernel_setup(kron_rows2_kernel<T, NN, STAGES, WARPS>, 32 * WARPS, C::SMEM, occ);
    if (e != cudaSuccess) return e;
    const long long slots = (long long)sms * occ * C::WARPS * C::IPW;
    long long L = ((long long)nb + slots - 1) / slots;
    if (L > 64) L = (L + 31) / 32 * 32; // ASGarD-style runs of 32 items per output are not cut more often than needed
    const long long per_cta = L * C::IPW * C::WARPS;
    const long long grid    = ((long long)nb + per_cta - 1) / per_cta;
    kron_rows2_kernel<T, NN, STAGES, WARPS><<<(int)grid, 32 * WARPS, C::SMEM, st>>>(A, in, out, lda, nb, (int)L);
    launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template<typename T, int NN>
static cudaError_t launch_rows2(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb, cudaStream_t st,
                                std::atomic<long long> &launches)
{
    int v = rows2_variant().load(std::memory_order_relaxed);
    // measured (profiles/rows2_r02.md): a third stage pays where a round moves the most bytes
    if (v < 0) v = (NN == 10 || (sizeof(T) == 4 && NN == 8)) ? 1 : 0;
    switch (v)
    {
    case 1: return launch_rows2v<T, NN, 3, 4>(sms, A, lda, in, out, nb, st, launches);
    // (two-warp CTAs with 2 / 3 stages were measured too: never ahead, profiles/rows2_ab_v1_r02.jsonl)
    default: return launch_rows2v<T, NN, 2, 4>(sms, A, lda, in, out, nb, st, launches);
    }
}

#define KRON_ROWS2_DEFINE_RUN(TYPE)                                                                                    \
    template<>                                                                                                        \
    cudaError_t run_rows2<TYPE>(int sms, int n, const TYPE *const *A, int lda, TYPE *const *in, TYPE *const *out,     \
                                int nb, cudaStream_t st, std::atomic<long long> &launches)                            \
    {                                                                                                                 \
        switch (n)                                                                                                    \
        {                                                                                                             \
        case 5: return launch_rows2<TYPE, 5>(sms, A, lda, in, out, nb, st, launches);                                 \
        case 6: return launch_rows2<TYPE, 6>(sms, A, lda, in, out, nb, st, launches);                                 \
        case 7: return launch_rows2<TYPE, 7>(sms, A, lda, in, out, nb, st, launches);                                 \
        case 8: return launch_rows2<TYPE, 8>(sms, A, lda, in, out, nb, st, launches);                                 \
        case 9: return launch_rows2<TYPE, 9>(sms, A, lda, in, out, nb, st, launches);                                 \
        case 10: return launch_rows2<TYPE, 10>(sms, A, lda, in, out, nb, st, launches);                               \
        }                                                                                                             \
        return cudaErrorNotSupported;                                                                                 \
    }

#endif // KRON_ROWS2_DEFINE

} // namespace kron
