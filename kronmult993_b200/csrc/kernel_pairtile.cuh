// kernel_pairtile.cuh -- compile-time (n, d) kernels built from n x n register tiles ("pairtile" path).
//
// Replaces cuda_kronmult_batchelement / cuda_kronmult / transpose / multiply_transpose
// (kronmult_gpu/kronmult.cu:139-167, :95-130, :34-43, :54-78) for the shapes of the reference's sweep
// envelope (tests/kronmult_fullbench_gpu.cpp:70-74, n in [2,10], d in [2,6]) whose vector fits in shared
// memory and that no more specialised family (tiny, dmma, wspec, wspec5, regtile) claims.
//
// Design (not a port):
//  * The item's vector lives in shared memory; the d mode products are applied in place, TWO factors per
//    round trip: a thread owns an n x n tile (the two indices of a factor pair), pulls it into registers,
//    applies the faster factor along the rows and the slower one along the columns (2 n^3 FMAs per
//    2 n^2 shared-memory accesses) and writes it back.  Odd d ends with a single-factor pass on the
//    same tile shape.  Every dot product sums k ascending from 0 like multiply_transpose
//    (kronmult.cu:66-70).  The reference moves 2 n^d elements through GLOBAL memory per factor.
//  * Layout: logical index i sits at i + PADE * (i / n^2): for even n the n^2-element blocks are padded
//    by 16 bytes so that their pitch is an odd number of 16-byte units -- the first pass, where every
//    lane owns one contiguous block, then runs on conflict-free 128-bit accesses; odd n needs no pad
//    (odd pitch, scalar accesses).  Later passes walk consecutive lanes along the fastest free index.
//  * B items ("streams") side by side per CTA when n^(d-2) tiles do not fill it; every stream walks a
//    range of consecutive items, so runs of equal output pointers are summed on chip (accumulator in
//    shared memory, touched by the same thread every time -> no synchronisation) and flushed once.
//    Every flush is REDG, like the reference's atomicAdd (kronmult.cu:126-129).
//  * Two stages: vector and factors of the next step arrive by cp.async (16-byte chunks when the
//    item is 16-byte aligned, element-wise otherwise; factors stay column-major with a 16-byte column
//    pitch, so columns are fetched as broadcast 128-bit loads), pointers one step further ahead.
//    Vectors too large for two stages plus the accumulator use one stage.
//  * CTAs per SM / streams per CTA / threads are picked at compile time per (T, n, d) from the register
//    and shared-memory footprints (PairCfg::pick): small CTAs in different phases overlap each other.
#pragma once
#include "common.cuh"
#include "kernel_regtile.cuh" // cp_async_elem / cp_async_commit / cp_async_wait_all
#include <atomic>

namespace kron
{

struct PairPick
{
    int minb, stages, acc, B, threads, fits;
};

// Geometry shared by the resident kernel (a whole item per tile) and the pass kernel of the multi-pass route (a tile
// of a longer vector): a tensor of DT indices of extent n, of which the GF slowest carry a factor and the Q = DT - GF
// fastest are only along for the ride (contiguous rows of n^Q elements).
template<typename T_, int n_, int DT_, int GF_>
struct PairBase
{
    using T = T_;
    static constexpr int n     = n_;
    static constexpr int DT    = DT_;
    static constexpr int GF    = GF_;
    static constexpr int Q     = DT_ - GF_;
    static constexpr int S     = (int)sizeof(T);
    static constexpr int N     = ipow(n, DT);
    static constexpr int NSQ   = n * n;
    static constexpr int TP    = N / NSQ; // tiles (= columns of every pass) per item
    static constexpr int VEC   = 16 / S;
    static constexpr bool VECTILE = (NSQ % VEC) == 0;
    static constexpr int PADE  = (VECTILE && ((NSQ / VEC) % 2) == 0) ? VEC : 0;
    static constexpr int BLK   = NSQ + PADE;
    static constexpr int ITEM  = TP * BLK;
    static constexpr int NPAIR = GF / 2;
    static constexpr int ODD   = GF % 2;
    static constexpr int NPASS = NPAIR + ODD;

    // Item pitch.  When an item has fewer than 32 tiles a warp spans several streams, and the pitch between them
    // decides the bank conflicts of every pass: the first pass (one contiguous block per lane) likes the plain
    // TP * BLK, the later ones (consecutive lanes on consecutive elements of one item, next item a pitch away) like
    // a pitch that lets the items of a half-warp tile the banks.  A model of the shared-memory wavefronts (16-byte
    // accesses by quarter-warps, 8-byte by half-warps, 4-byte by whole warps; equal addresses broadcast) picks the
    // pad, at compile time.
    static constexpr int wavefronts(const int (&addr)[32], int abytes) // addr in bytes, one access of abytes per lane
    {
        const int group = abytes == 16 ? 8 : (abytes == 8 ? 16 : 32);
        int total = 0;
        for (int g0 = 0; g0 < 32; g0 += group)
        {
            int cnt[32] = {};
            int worst   = 0;
            for (int l = g0; l < g0 + group; ++l)
            {
                if (addr[l] < 0) continue;
                bool dup = false;
                for (int m = g0; m < l; ++m) dup = dup || addr[m] == addr[l];
                if (dup) continue;
                const int bank = (addr[l] / abytes) % (128 / abytes);
                if (++cnt[bank] > worst) worst = cnt[bank];
            }
            total += worst;
        }
        return total;
    }
    static constexpr long long pitch_cost(int pitch)
    {
        long long cost = 0;
        int addr[32]   = {};
        // first pass (Q = 0 only): lane tl = (b, c) owns block c of stream b
        if (Q == 0)
        {
            for (int l = 0; l < 32; ++l) addr[l] = l < TP * 32 ? ((l / TP) * pitch + (l % TP) * BLK) * S : -1;
            cost += (long long)wavefronts(addr, VECTILE ? 16 : S) * (VECTILE ? 2 * (NSQ / VEC) : 2 * NSQ);
        }
        return cost + later_cost(pitch, 32, 0);
    }
    // later passes: lane tl = (b, c), c = lo + LO * hi; lanemap 0: c fastest (b = tl / TP), 1: b fastest (b = tl % B)
    static constexpr long long later_cost(int pitch, int B, int lanemap)
    {
        long long cost = 0;
        int addr[32]   = {};
        for (int p = (Q == 0 ? 1 : 0); p < NPASS; ++p)
        {
            const bool single = (p == NPAIR);
            const int sB      = single ? ipow(n, DT - 2) : ipow(n, Q + 2 * p);
            const int sH      = sB * NSQ;
            for (int l = 0; l < 32; ++l)
            {
                const int b = lanemap ? l % B : l / TP, c = lanemap ? l / B : l % TP;
                const int lo = c % sB, hi = c / sB;
                addr[l] = (b < B && c < TP) ? (b * pitch + lo + (lo / NSQ) * PADE + hi * (sH / NSQ) * BLK) * S : -1;
            }
            cost += (long long)wavefronts(addr, S) * (2 * NSQ);
        }
        return cost;
    }
    static constexpr int pick_pitch()
    {
        const int base = (ITEM + VEC - 1) / VEC * VEC;
        if (TP >= 32 || SPLIT > 1) return base; // (split tiles follow another lane map: not modelled)
        int best = base;
        long long best_cost = pitch_cost(base);
        for (int pad = VEC; pad <= 16 * (8 / S) * 2; pad += VEC)
        {
            const long long c = pitch_cost(base + pad);
            if (c < best_cost) { best_cost = c; best = base + pad; }
        }
        return best;
    }
    static constexpr int ITEMP = pick_pitch(); // item pitch: a 16-byte multiple
    static constexpr int RP    = (n + VEC - 1) / VEC * VEC;    // pitch of a factor column (column-major, as in global)
    static constexpr int MAT   = GF * n * RP;
    // factor pitch per stream: an odd number of 16-byte units, so that the broadcast column loads of lanes that
    // belong to different streams fall into different banks
    static constexpr int MATP  = ((MAT / VEC) % 2 == 0) ? MAT + VEC : MAT;
    static constexpr int NCH   = N / VEC; // whole 16-byte chunks of a vector
    static constexpr int TAIL  = N % VEC;
    static constexpr int LBQ   = ipow(n, Q); // contiguous row length of a tile of a longer vector

    // register tiles: rows are processed in HB blocks so that n^2 + n^2/HB values are live at a time
    static constexpr int HB   = (S == 8) ? (n == 9 ? 3 : (n >= 7 ? 2 : 1)) : 1;
    static constexpr int RB   = (n + HB - 1) / HB;
    // tiles of n >= 9 exceed the register file in fp64 (2 n^2 live values) and leave too few warps in fp32: two threads
    // share a tile, each producing CB of its n output columns (both read the whole tile: adjacent lanes, so the loads
    // are broadcasts)
    // (also fp32 n = 8, d = 4: 64 tiles per item -> 128 threads per CTA instead of 64; measured 1.42x, the other n = 7, 8
    // shapes lose 1-22 % with split tiles; and fp64 n = 8: d = 2 1.13x, d = 3 1.50x, multi-pass d = 5 1.13x;
    // fp64 n = 7 is a wash: d = 3, 5 +6 %, d = 4, 6 -15 %)
    static constexpr int SPLIT = (n >= 9 || (S == 4 && n == 8 && DT_ == 4 && GF_ == 4) || (S == 8 && n == 8)) ? 2 : 1;
    static constexpr int CB    = (n + SPLIT - 1) / SPLIT;
    static constexpr int TPS   = TP * SPLIT; // thread-tiles per item
    static constexpr int KUNROLL = (n == 9 && DT_ == GF_) ? 1 : n; // split tiles: trips of the first product to unroll
    static constexpr int REGS  = (SPLIT == 1 ? (NSQ + RB * n + n) * (S / 4) + 48 : (S == 8 ? 255 : 168)); // estimate
};

template<typename T_, int n_, int d>
struct PairCfg : PairBase<T_, n_, d, d>
{
    using Base = PairBase<T_, n_, d, d>;
    using Base::S; using Base::ITEMP; using Base::MATP; using Base::TPS; using Base::REGS;
    static constexpr bool TO_OUT   = true;   // the last pass accumulates into `output`
    static constexpr bool REGFLUSH = d >= 3; // ... straight from registers (d = 2: one thread per item, via shared memory)
    static constexpr int PSTRIDE   = d + 2;  // pointer slot layout per stream: d factors, input, output
    static constexpr int POUT      = d + 1;

    static constexpr int PTRB = 3 * ((d + 2) * 8 + 4); // three slots of pointers + flag per stream
    static constexpr int bytes_item(int stages, int acc) { return (stages + acc) * ITEMP * S + stages * MATP * S + PTRB; }

    // CTAs per SM, streams per CTA and threads: as many resident tiles per SM as registers and shared memory
    // allow, in as many independent CTAs as possible (CTAs in different phases overlap shared-memory traffic,
    // FMAs and barriers of each other)
    static constexpr PairPick pick()
    {
        int tt = 65536 / REGS;
        if (tt > 512) tt = 512;
        tt = tt / 32 * 32;
        PairPick best{1, 1, 0, 1, 32, 0};
        int best_total = -1;
        for (int minb = 16; minb >= 1; --minb)
        {
            if (minb == 15 || minb == 14 || minb == 13 || minb == 11 || minb == 9 || minb == 7 || minb == 5) continue;
            const int budget = 227 * 1024 / minb - 1024 - 128; // the driver reserves 1 KiB per CTA
            int cap = (tt / minb) / 32 * 32;
            if (cap > 256) cap = 256;
            if (cap < 32) continue;
            // one stage suffices when >= 3 CTAs share the SM (the others' work hides the load); a single fat CTA falls
            // back to one stage (and no accumulator) only when nothing else fits
            for (int stages = 2; stages >= 1; --stages)
            {
                int acc = 1;
                if (stages == 1 && minb < 3 && bytes_item(2, 1) <= budget) continue;
                if (bytes_item(stages, 1) > budget)
                {
                    if (minb > 1 || stages == 2) continue;
                    acc = 0;
                    if (bytes_item(1, 0) > budget) continue;
                }
                const int bmax = budget / bytes_item(stages, acc);
                int B          = TPS >= cap ? 1 : cap / TPS;
                if (B > bmax) B = bmax;
                const int tiles   = B * TPS;
                const int iters   = (tiles + cap - 1) / cap;
                const int threads = ((tiles + iters - 1) / iters + 31) / 32 * 32;
                int total         = minb * (tiles / iters);
                if (total > tt) total = tt;
                if (threads <= 32) total += total / 4; // single-warp CTAs never wait for another warp at a barrier
                if (total * 10 > best_total * 11) // fewer, fatter CTAs only for >10% more resident tiles
                {
                    best_total = total;
                    best       = PairPick{minb, stages, acc, B, threads, 1};
                }
            }
        }
        return best;
    }
    static constexpr PairPick P = pick();
    static constexpr int MINB    = P.minb;
    static constexpr int STAGES  = P.stages;
    static constexpr int ACC     = P.acc;
    static constexpr bool FITS   = P.fits != 0;
    static constexpr int B       = P.B;
    static constexpr int THREADS = P.threads;
    static constexpr int TILES   = B * TPS; // thread-tiles per CTA step
    // lane -> (stream, column) map of every pass but the first: streams fastest when the wavefront model of
    // PairBase says so for this pitch and stream count (e.g. n = 6, d = 3 fp64: 2 wavefronts per access instead of 4)
    static constexpr int LANEMAP = (B > 1 && Base::SPLIT == 1 && Base::TP < 32
                                    && Base::later_cost(ITEMP, B, 1) < Base::later_cost(ITEMP, B, 0)) ? 1 : 0;

    // byte offsets into dynamic shared memory
    static constexpr int OFF_ACC  = STAGES * B * ITEMP * S;
    static constexpr int OFF_MAT  = OFF_ACC + ACC * B * ITEMP * S;
    static constexpr int OFF_PTR  = OFF_MAT + STAGES * B * MATP * S;
    static constexpr int OFF_FLAG = OFF_PTR + 3 * B * (d + 2) * 8;
    static constexpr int SMEM     = OFF_FLAG + ((3 * B * 4 + 15) / 16) * 16;
};

template<typename C_, int PASS>
struct PairGeom
{
    using C = C_;
    static constexpr int n = C::n;
    static constexpr bool SINGLE = (PASS == C::NPAIR); // the trailing single-factor pass of an odd factor count
    static constexpr bool FINAL  = (PASS == C::NPASS - 1);
    static constexpr int sB_log  = SINGLE ? ipow(n, C::DT - 2) : ipow(n, C::Q + 2 * PASS);
    static constexpr int sA_log  = sB_log * n;
    static constexpr int LO      = sB_log; // columns below the tile's two indices
    static constexpr int sH_log  = sA_log * n;
    static constexpr bool CONTIG = !SINGLE && sB_log == 1; // the thread's tile is one contiguous block
    static constexpr int phys_stride(int x) { return x >= C::NSQ ? x / C::NSQ * C::BLK : x; }
    static constexpr int sB_ph = phys_stride(sB_log);
    static constexpr int sA_ph = phys_stride(sA_log);
    static constexpr int sH_ph = phys_stride(sH_log);
    static constexpr int jb    = C::GF - 1 - 2 * PASS; // faster factor of the pair (unused when SINGLE)
    static constexpr int ja    = SINGLE ? 0 : C::GF - 2 - 2 * PASS;
    // results of the last pass leave through REDG straight from registers
    static constexpr bool REGFLUSH = FINAL && C::TO_OUT && C::REGFLUSH;
};

// Offset in `output` of logical tile index i: rows of LBQ contiguous elements lie Lg apart in the long vector
// (resident items: Q = 0 and the whole item is one tile -> the identity).
template<typename C>
__device__ __forceinline__ long long tile_goff(int i, long long Lg)
{
    if constexpr (C::DT == C::GF) return i;
    else return (long long)(i / C::LBQ) * Lg + (i % C::LBQ);
}

// one column of a factor (column-major, pitch RP) as broadcast 128-bit loads
template<typename T, int n, int RP>
__device__ __forceinline__ void load_col(const T *__restrict__ col, T (&m)[n])
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int4 *q = reinterpret_cast<const int4 *>(col);
#pragma unroll
    for (int i = 0; i < RP / VEC; ++i)
    {
        const int4 w = q[i];
        if constexpr (sizeof(T) == 8)
        {
            if (2 * i < n) m[2 * i] = __hiloint2double(w.y, w.x);
            if (2 * i + 1 < n) m[2 * i + 1] = __hiloint2double(w.w, w.z);
        }
        else
        {
            if (4 * i < n) m[4 * i] = __int_as_float(w.x);
            if (4 * i + 1 < n) m[4 * i + 1] = __int_as_float(w.y);
            if (4 * i + 2 < n) m[4 * i + 2] = __int_as_float(w.z);
            if (4 * i + 3 < n) m[4 * i + 3] = __int_as_float(w.w);
        }
    }
}

// One tile of one pass.  `vec` = the item's vector in shared memory, `acc` = the item's run accumulator,
// `mats` = the item's d factors (column-major, pitch RP), `c` = column index in [0, TP), flag bits: 2 = first
// item of a run of equal output pointers, 4 = last item of the run.
// Both products run with the summation index k outermost: n * RB independent FMA chains advance together.
template<typename C, int PASS>
__device__ __forceinline__ void pair_tile(typename C::T *__restrict__ vec, typename C::T *__restrict__ acc,
                                          const typename C::T *__restrict__ mats, int c, int flag,
                                          typename C::T *__restrict__ outp, long long Lg)
{
    using T = typename C::T;
    using G = PairGeom<C, PASS>;
    constexpr int n         = C::n;
    constexpr bool REGFLUSH = G::REGFLUSH;
    constexpr bool CONTIG   = G::CONTIG;
    constexpr int RB        = C::RB;

    const int lo = (G::LO > 1) ? c % G::LO : 0;
    const int hi = (G::LO > 1) ? c / G::LO : c;
    int base_ph  = hi * G::sH_ph + lo;
    if constexpr (C::PADE > 0 && G::LO > C::NSQ) base_ph += (lo / C::NSQ) * C::PADE;
    const int base_log = hi * G::sH_log + lo;
    T *__restrict__ x  = vec + base_ph;
    // where the tile sits in `output` (REGFLUSH only): the tile's two indices lie above the contiguous rows
    long long gbase = 0, gA = G::sA_log, gB = G::sB_log;
    if constexpr (REGFLUSH)
    {
        gbase = tile_goff<C>(base_log, Lg);
        if constexpr (C::DT != C::GF)
        {
            // an index at or above the contiguous rows moves by whole rows; the second index of a single-factor
            // pass may lie inside the row
            if constexpr (G::sA_log >= C::LBQ) gA = (long long)(G::sA_log / C::LBQ) * Lg;
            if constexpr (G::sB_log >= C::LBQ) gB = (long long)(G::sB_log / C::LBQ) * Lg;
        }
    }

    // faster factor along the rows: Z[a][b'] = sum_k Mb(b', k) X[a][k]
    T Z[n][n];
#pragma unroll
    for (int h = 0; h < C::HB; ++h)
    {
        const int a0 = h * RB;
        T X[RB][n];
        if constexpr (CONTIG && C::VECTILE && (RB * n) % C::VEC == 0)
        {
            const int4 *q = reinterpret_cast<const int4 *>(x + a0 * n);
#pragma unroll
            for (int i = 0; i < RB * n / C::VEC; ++i)
            {
                if (a0 * n + i * C::VEC >= C::NSQ) continue;
                const int4 w = q[i];
                if constexpr (sizeof(T) == 8)
                {
                    X[(2 * i) / n][(2 * i) % n]         = __hiloint2double(w.y, w.x);
                    X[(2 * i + 1) / n][(2 * i + 1) % n] = __hiloint2double(w.w, w.z);
                }
                else
                {
                    X[(4 * i) / n][(4 * i) % n]         = __int_as_float(w.x);
                    X[(4 * i + 1) / n][(4 * i + 1) % n] = __int_as_float(w.y);
                    X[(4 * i + 2) / n][(4 * i + 2) % n] = __int_as_float(w.z);
                    X[(4 * i + 3) / n][(4 * i + 3) % n] = __int_as_float(w.w);
                }
            }
        }
        else
        {
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int b = 0; b < n; ++b)
                    if (a0 + a < n) X[a][b] = x[(a0 + a) * G::sA_ph + b * G::sB_ph];
        }
        if constexpr (!G::SINGLE)
        {
            const T *__restrict__ Mb = mats + G::jb * n * C::RP;
#pragma unroll
            for (int k = 0; k < n; ++k)
            {
                T m[n];
                load_col<T, n, C::RP>(Mb + k * C::RP, m);
#pragma unroll
                for (int a = 0; a < RB; ++a)
#pragma unroll
                    for (int bp = 0; bp < n; ++bp)
                        if (a0 + a < n) Z[a0 + a][bp] = (k == 0) ? X[a][0] * m[bp] : fma(X[a][k], m[bp], Z[a0 + a][bp]);
            }
        }
        else
        {
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int b = 0; b < n; ++b)
                    if (a0 + a < n) Z[a0 + a][b] = X[a][b];
        }
    }

    // slower factor down the columns: Y[a'][b'] = sum_k Ma(a', k) Z[k][b']
    const T *__restrict__ Ma = mats + G::ja * n * C::RP;
#pragma unroll
    for (int h = 0; h < C::HB; ++h)
    {
        const int a0 = h * RB;
        T Y[RB][n];
#pragma unroll
        for (int k = 0; k < n; ++k)
        {
            T m[n];
            load_col<T, n, C::RP>(Ma + k * C::RP, m);
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int bp = 0; bp < n; ++bp)
                    if (a0 + a < n)
                    {
                        if (k == 0)
                        {
                            // inside a run of equal outputs the running sum is the addend of the first FMA: its
                            // shared-memory load is issued ahead of the whole product instead of after it
                            if (REGFLUSH && C::ACC && !(flag & 2))
                                Y[a][bp] = fma(Z[0][bp], m[a0 + a], acc[base_ph + (a0 + a) * G::sA_ph + bp * G::sB_ph]);
                            else
                                Y[a][bp] = Z[0][bp] * m[a0 + a];
                        }
                        else Y[a][bp] = fma(Z[k][bp], m[a0 + a], Y[a][bp]);
                    }
        }
#pragma unroll
        for (int a = 0; a < RB; ++a)
        {
            const int ap = a0 + a;
            if (ap >= n) continue;
            if constexpr (REGFLUSH)
            {
#pragma unroll
                for (int bp = 0; bp < n; ++bp)
                {
                    const int ph       = base_ph + ap * G::sA_ph + bp * G::sB_ph;
                    const long long lg = gbase + ap * gA + bp * gB;
                    T v                = Y[a][bp];
                    if constexpr (C::ACC)
                    {
                        if (flag & 4) red_add(outp + lg, v);
                        else acc[ph] = v;
                    }
                    else { red_add(outp + lg, v); }
                }
            }
            else if constexpr (CONTIG && (n % 2) == 0)
            {
                // contiguous row of the tile: 2-element stores (n even keeps them aligned)
#pragma unroll
                for (int bp = 0; bp < n; bp += 2)
                {
                    if constexpr (sizeof(T) == 8)
                        *reinterpret_cast<double2 *>(x + ap * n + bp) = make_double2(Y[a][bp], Y[a][bp + 1]);
                    else
                        *reinterpret_cast<float2 *>(x + ap * n + bp) = make_float2(Y[a][bp], Y[a][bp + 1]);
                }
            }
            else
            {
#pragma unroll
                for (int bp = 0; bp < n; ++bp) x[ap * G::sA_ph + bp * G::sB_ph] = Y[a][bp];
            }
        }
    }
}

// The same for SPLIT = 2: this thread produces CB of the n output columns of the tile.  `half` is a
// run-time value, so everything that depends on it is an address offset, never a register index.
template<typename C, int PASS>
__device__ __forceinline__ void pair_tile_split(typename C::T *__restrict__ vec, typename C::T *__restrict__ acc,
                                                const typename C::T *__restrict__ mats, int c, int half, int flag,
                                                typename C::T *__restrict__ outp, long long Lg)
{
    using T = typename C::T;
    using G = PairGeom<C, PASS>;
    constexpr int n         = C::n;
    constexpr bool REGFLUSH = G::REGFLUSH;
    constexpr int RB = C::RB, CB = C::CB;

    const int lo = (G::LO > 1) ? c % G::LO : 0;
    const int hi = (G::LO > 1) ? c / G::LO : c;
    int base_ph  = hi * G::sH_ph + lo;
    if constexpr (C::PADE > 0 && G::LO > C::NSQ) base_ph += (lo / C::NSQ) * C::PADE;
    const int base_log = hi * G::sH_log + lo;
    // odd n: the halves overlap in one column, which the second thread computes too but does not store
    const int b0       = half * (n - CB);
    const int skip     = half * (2 * CB - n);
    long long gbase = 0, gA = G::sA_log, gB = G::sB_log;
    if constexpr (REGFLUSH)
    {
        gbase = tile_goff<C>(base_log, Lg);
        if constexpr (C::DT != C::GF)
        {
            // an index at or above the contiguous rows moves by whole rows; the second index of a single-factor
            // pass may lie inside the row
            if constexpr (G::sA_log >= C::LBQ) gA = (long long)(G::sA_log / C::LBQ) * Lg;
            if constexpr (G::sB_log >= C::LBQ) gB = (long long)(G::sB_log / C::LBQ) * Lg;
        }
    }
    T *__restrict__ x  = vec + base_ph;

    T Z[n][CB];
    if constexpr (!G::SINGLE)
    {
        const T *__restrict__ Mb = mats + G::jb * n * C::RP + b0;
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int j = 0; j < CB; ++j) Z[a][j] = T(0);
        // k only moves addresses here, so the loop may stay rolled: for n = 9 ptxas otherwise hoists the loads of
        // all trips (n^2 more live values) and spills; n = 10 fits and runs 7 % faster unrolled
#pragma unroll(C::KUNROLL)
        for (int k = 0; k < n; ++k)
        {
            T xc[n], m[CB];
#pragma unroll
            for (int a = 0; a < n; ++a) xc[a] = x[a * G::sA_ph + k * G::sB_ph];
#pragma unroll
            for (int j = 0; j < CB; ++j) m[j] = Mb[k * C::RP + j];
#pragma unroll
            for (int a = 0; a < n; ++a)
#pragma unroll
                for (int j = 0; j < CB; ++j) Z[a][j] = fma(xc[a], m[j], Z[a][j]);
        }
    }
    else
    {
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int j = 0; j < CB; ++j) Z[a][j] = x[a * G::sA_ph + (b0 + j) * G::sB_ph];
    }

    // The two lanes that share this tile both read ALL of it above and store their own columns in place below:
    // order the partner's loads before my stores.  Partners (adjacent lanes, same tile, same flag, same trip count)
    // follow identical control flow, so they are in the same convergence group.
    if constexpr (!REGFLUSH) __syncwarp(__activemask());
    const T *__restrict__ Ma = mats + G::ja * n * C::RP;
#pragma unroll
    for (int h = 0; h < C::HB; ++h)
    {
        const int a0 = h * RB;
        T Y[RB][CB];
#pragma unroll
        for (int k = 0; k < n; ++k)
        {
            T m[n];
            load_col<T, n, C::RP>(Ma + k * C::RP, m);
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int j = 0; j < CB; ++j)
                    if (a0 + a < n)
                    {
                        if (k == 0)
                        {
                            if (REGFLUSH && C::ACC && !(flag & 2))
                                Y[a][j] = fma(Z[0][j], m[a0 + a], acc[base_ph + (a0 + a) * G::sA_ph + (b0 + j) * G::sB_ph]);
                            else
                                Y[a][j] = Z[0][j] * m[a0 + a];
                        }
                        else Y[a][j] = fma(Z[k][j], m[a0 + a], Y[a][j]);
                    }
        }
#pragma unroll
        for (int a = 0; a < RB; ++a)
        {
            const int ap = a0 + a;
            if (ap >= n) continue;
#pragma unroll
            for (int j = 0; j < CB; ++j)
            {
                if (j < skip) continue;
                const int ph = base_ph + ap * G::sA_ph + (b0 + j) * G::sB_ph;
                if constexpr (REGFLUSH)
                {
                    const long long lg = gbase + ap * gA + (b0 + j) * gB;
                    T v                = Y[a][j];
                    if constexpr (C::ACC)
                    {
                        if (flag & 4) red_add(outp + lg, v);
                        else acc[ph] = v;
                    }
                    else { red_add(outp + lg, v); }
                }
                else { vec[ph] = Y[a][j]; }
            }
        }
    }
}

// Work distribution of the copy loops: groups of GL lanes (a power of two, or the whole CTA when there is a single
// stream) walk the streams, f(b, gl, GL) then loops over the stream's PER_ITEM pieces from gl in steps of GL -- the
// per-stream values (flag, pointers, alignment) are fetched once per stream and no division by PER_ITEM is needed.
template<int PER_ITEM, int B, int THREADS, typename F>
__device__ __forceinline__ void for_items(int tid, F f)
{
    if constexpr (B == 1) { f(0, tid, THREADS); }
    else
    {
        constexpr int GL = PER_ITEM >= 32 ? 32 : (PER_ITEM > 16 ? 32 : (PER_ITEM > 8 ? 16 : (PER_ITEM > 4 ? 8 : (PER_ITEM > 2 ? 4 : (PER_ITEM > 1 ? 2 : 1)))));
        constexpr int GROUPS = THREADS / GL;
        const int g = tid / GL, gl = tid % GL;
        for (int b = g; b < B; b += GROUPS) f(b, gl, GL);
    }
}

template<typename C, int PASS>
__device__ __forceinline__ void pair_passes(unsigned char *smem, int stage, const int *__restrict__ s_flag,
                                            typename C::T *const *__restrict__ s_ptr, long long Lg)
{
    using T = typename C::T;
    T *vecs       = reinterpret_cast<T *>(smem) + (size_t)stage * C::B * C::ITEMP;
    T *accs       = reinterpret_cast<T *>(smem + C::OFF_ACC);
    const T *mats = reinterpret_cast<const T *>(smem + C::OFF_MAT) + (size_t)stage * C::B * C::MATP;
#pragma unroll 1
    for (int tl = threadIdx.x; tl < C::TILES; tl += C::THREADS)
    {
        const int tile = (C::SPLIT > 1) ? tl / C::SPLIT : tl;
        constexpr bool BFAST = C::LANEMAP == 1 && !PairGeom<C, PASS>::CONTIG;
        const int b    = (C::B > 1) ? (BFAST ? tile % C::B : tile / C::TP) : 0;
        const int c    = (C::B > 1) ? (BFAST ? tile / C::B : tile - b * C::TP) : tile;
        const int flag = s_flag[b];
        if (!flag) continue;
        if constexpr (C::SPLIT > 1)
            pair_tile_split<C, PASS>(vecs + b * C::ITEMP, accs + b * C::ITEMP, mats + b * C::MATP, c,
                                     tl - tile * C::SPLIT, flag, s_ptr[b * C::PSTRIDE + C::POUT], Lg);
        else
            pair_tile<C, PASS>(vecs + b * C::ITEMP, accs + b * C::ITEMP, mats + b * C::MATP, c, flag,
                               s_ptr[b * C::PSTRIDE + C::POUT], Lg);
    }
    if constexpr (PASS + 1 < C::NPASS)
    {
        __syncthreads();
        pair_passes<C, PASS + 1>(smem, stage, s_flag, s_ptr, Lg);
    }
}

template<typename T, int n, int d>
__global__ void __launch_bounds__(PairCfg<T, n, d>::THREADS, PairCfg<T, n, d>::MINB)
    kron_pairtile_kernel(const T *const *__restrict__ A, T *const *__restrict__ in, T *const *__restrict__ out,
                         int lda, int nb)
{
    using C = PairCfg<T, n, d>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;

    // this CTA's contiguous item range, cut into B streams of L consecutive items
    const long long per = ((long long)nb + gridDim.x - 1) / gridDim.x;
    const long long r0  = per * blockIdx.x;
    long long r1        = r0 + per;
    if (r1 > nb) r1 = nb;
    if (r0 >= r1) return;
    const int L = (int)((r1 - r0 + C::B - 1) / C::B);

    auto slot_ptrs = [&](int slot) { return reinterpret_cast<T **>(smem + C::OFF_PTR) + (size_t)slot * C::B * (d + 2); };
    auto slot_flag = [&](int slot) { return reinterpret_cast<int *>(smem + C::OFF_FLAG) + slot * C::B; };

    // pointers of step t: per stream d factor pointers, input, output; flag = valid | first<<1 | last<<2
    auto fetch_ptrs = [&](int t) {
        T **sp  = slot_ptrs(t % 3);
        int *sf = slot_flag(t % 3);
        for (int e = tid; e < C::B * (d + 2); e += C::THREADS)
        {
            const int b        = e / (d + 2);
            const int w        = e - b * (d + 2);
            const long long s0 = r0 + (long long)b * L;
            long long s1       = s0 + L;
            if (s1 > r1) s1 = r1;
            const long long k = s0 + t;
            const bool valid  = k < s1;
            T *p              = nullptr;
            if (valid)
            {
                if (w < d) p = const_cast<T *>(A[k * d + w]);
                else if (w == d) p = in[k];
                else
                {
                    p                = out[k];
                    const bool first = (t == 0) || (out[k - 1] != p);
                    const bool last  = (k + 1 >= s1) || (out[k + 1] != p);
                    sf[b]            = 1 | (first ? 2 : 0) | (last ? 4 : 0);
                }
            }
            else if (w == d + 1) sf[b] = 0;
            sp[e] = p;
        }
    };

    // vector and factors of step t -> stage
    auto issue_copies = [&](int t, int stage) {
        T *const *sp  = slot_ptrs(t % 3);
        const int *sf = slot_flag(t % 3);
        T *vecs       = reinterpret_cast<T *>(smem) + (size_t)stage * C::B * C::ITEMP;
        T *mats       = reinterpret_cast<T *>(smem + C::OFF_MAT) + (size_t)stage * C::B * C::MATP;
        // the vector: 16-byte chunks (padded destination) when the item is 16-byte aligned, element-wise otherwise
        for_items<(C::NCH > 0 ? C::NCH : 1), C::B, C::THREADS>(tid, [&](int b, int gl, int GL) {
            if (!sf[b]) return;
            const T *src = sp[b * (d + 2) + d];
            T *dst       = vecs + b * C::ITEMP;
            if (aligned16(src))
            {
                for (int q = gl; q < C::NCH; q += GL)
                {
                    int el = q * C::VEC;
                    if constexpr (C::PADE > 0) el += (q / (C::NSQ / C::VEC)) * C::PADE;
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + el);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + q * C::VEC) : "memory");
                }
                for (int i = C::NCH * C::VEC + gl; i < C::N; i += GL) cp_async_elem<T>(dst + i, src + i); // no pad here
            }
            else
            {
                for (int i = gl; i < C::N; i += GL)
                {
                    int el = i;
                    if constexpr (C::PADE > 0) el += (i / C::NSQ) * C::PADE;
                    cp_async_elem<T>(dst + el, src + i);
                }
            }
        });
        // factors stay column-major (pitch RP): whole 16-byte chunks of a column when n, lda and the base address
        // allow it, element-wise otherwise
        constexpr int CPC = (n % C::VEC == 0) ? n / C::VEC : 1; // chunks per column
        const bool lda_ok = (n % C::VEC == 0) && (lda % C::VEC == 0);
        for_items<d * C::NSQ, C::B, C::THREADS>(tid, [&](int b, int gl, int GL) {
            if (!sf[b]) return;
            if (lda_ok)
            {
                for (int r = gl; r < d * n * CPC; r += GL)
                {
                    const int j  = r / (n * CPC);
                    const int rc = r - j * (n * CPC);
                    const int cc = rc / CPC;
                    const int q  = rc - cc * CPC;
                    const T *src = sp[b * (d + 2) + j] + (long long)cc * lda + q * C::VEC;
                    T *dst       = mats + b * C::MATP + j * n * C::RP + cc * C::RP + q * C::VEC;
                    if (aligned16(src))
                    {
                        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < C::VEC; ++i) cp_async_elem<T>(dst + i, src + i);
                    }
                }
            }
            else
            {
                for (int r = gl; r < d * C::NSQ; r += GL)
                {
                    const int j  = r / C::NSQ;
                    const int rc = r - j * C::NSQ;
                    const int cc = rc / n;
                    const int rr = rc - cc * n;
                    cp_async_elem<T>(mats + b * C::MATP + j * n * C::RP + cc * C::RP + rr,
                                     sp[b * (d + 2) + j] + rr + (long long)cc * lda);
                }
            }
        });
        cp_async_commit();
    };

    fetch_ptrs(0);
    __syncthreads();
    issue_copies(0, 0);
    if (L > 1) fetch_ptrs(1);

    for (int t = 0; t < L; ++t)
    {
        const int stage = (C::STAGES == 2) ? (t & 1) : 0;
        cp_async_wait_all();
        __syncthreads(); // step t's data and step t+1's pointers are visible; the other stage is free
        if constexpr (C::STAGES == 2)
        {
            if (t + 1 < L) issue_copies(t + 1, stage ^ 1);
        }
        if (t + 2 < L) fetch_ptrs(t + 2);

        const int *sf = slot_flag(t % 3);
        pair_passes<C, 0>(smem, stage, sf, slot_ptrs(t % 3), 0);

        if constexpr (d == 2)
        {
            // one thread per item in the pass above: leave through shared memory so that the REDG are coalesced
            __syncthreads();
            T *vecs = reinterpret_cast<T *>(smem) + (size_t)stage * C::B * C::ITEMP;
            T *accs = reinterpret_cast<T *>(smem + C::OFF_ACC);
            T *const *sp = slot_ptrs(t % 3);
            for (int e = tid; e < C::B * C::N; e += C::THREADS)
            {
                const int b    = e / C::N;
                const int i    = e - b * C::N;
                const int flag = sf[b];
                if (!flag) continue;
                const int ph = b * C::ITEMP + i + (C::PADE > 0 ? (i / C::NSQ) * C::PADE : 0);
                T v          = vecs[ph];
                T *o         = sp[b * (d + 2) + d + 1] + i;
                if constexpr (C::ACC)
                {
                    if (!(flag & 2)) v += accs[ph];
                    if (flag & 4) red_add(o, v);
                    else accs[ph] = v;
                }
                else { red_add(o, v); }
            }
        }
        if constexpr (C::STAGES == 1)
        {
            if (t + 1 < L)
            {
                __syncthreads(); // the single stage is free again
                issue_copies(t + 1, 0);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
template<typename T, int n, int d>
constexpr bool pairtile_has() { return PairCfg<T, n, d>::FITS; }
template<typename T>
static bool pairtile_fits(int d, int n)
{
    switch (n * 16 + d)
    {
#define KRON_PTF(NN, DD) case NN * 16 + DD: return pairtile_has<T, NN, DD>();
#define KRON_PTF_N(NN) KRON_PTF(NN, 2) KRON_PTF(NN, 3) KRON_PTF(NN, 4) KRON_PTF(NN, 5) KRON_PTF(NN, 6)
        KRON_PTF_N(2) KRON_PTF_N(3) KRON_PTF_N(4) KRON_PTF_N(5) KRON_PTF_N(6) KRON_PTF_N(7) KRON_PTF_N(8) KRON_PTF_N(9)
        KRON_PTF_N(10)
#undef KRON_PTF_N
#undef KRON_PTF
    default: return false;
    }
}

template<typename T, int n, int d>
static cudaError_t launch_pairtile(int sms, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                   cudaStream_t st, std::atomic<long long> &launches)
{
    using C = PairCfg<T, n, d>;
    if constexpr (!pairtile_has<T, n, d>()) { return cudaErrorNotSupported; }
    else
    {
        auto kfn = kron_pairtile_kernel<T, n, d>;
        {
            cudaError_t e = kernel_setup(kfn, C::SMEM); // per device, cached (common.cuh)
            if (e != cudaSuccess) return e;
        }
        const long long want = (long long)sms * C::MINB;
        const int grid       = (int)(nb < want ? nb : want);
        kfn<<<grid, C::THREADS, C::SMEM, st>>>(A, in, out, lda, nb);
        launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
}

// cudaErrorNotSupported when (n, d) is outside the family; defined in pairtile_f64.cu / pairtile_f32.cu
template<typename T>
cudaError_t run_pairtile(int sms, int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                         cudaStream_t st, std::atomic<long long> &launches);

#define KRON_PAIRTILE_DEFINE(TYPE)                                                                                  \
    template<>                                                                                                      \
    cudaError_t run_pairtile<TYPE>(int sms, int d, int n, const TYPE *const *A, int lda, TYPE *const *in,           \
                                   TYPE *const *out, int nb, cudaStream_t st, std::atomic<long long> &launches)    \
    {                                                                                                               \
        switch (n * 16 + d)                                                                                         \
        {                                                                                                           \
            KRON_PT_N(2) KRON_PT_N(3) KRON_PT_N(4) KRON_PT_N(5) KRON_PT_N(6) KRON_PT_N(7) KRON_PT_N(8) KRON_PT_N(9)   \
            KRON_PT_N(10)                                                                                           \
        default: return cudaErrorNotSupported;                                                                      \
        }                                                                                                           \
    }
#define KRON_PT_CASE(NN, DD) \
    case NN * 16 + DD: return launch_pairtile<KRON_PT_TYPE, NN, DD>(sms, A, lda, in, out, nb, st, launches);
#define KRON_PT_N(NN) KRON_PT_CASE(NN, 2) KRON_PT_CASE(NN, 3) KRON_PT_CASE(NN, 4) KRON_PT_CASE(NN, 5) KRON_PT_CASE(NN, 6)

} // namespace kron
