/*
 * oracle/kronmult_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's `kronmult_batched` algorithm
 * (project-asgard/kronmult993, kronmult_omp flavour, no-BLAS path).  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path; nothing in the product (kronmult993_b200/, include/) links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every routine here
 *   (1) bit-for-bit against oracle/_ref/libkronmult_ref.so, which is the reference's own
 *       header kronmult_omp/kronmult.hpp compiled where it lies (oracle/Makefile), when that
 *       library is present (it is built in the dev container and travels to the GPU box);
 *   (2) against the committed golden vectors tests/golden/*.npz that were produced by that
 *       same reference build (tests/golden/make_golden.py);
 *   (3) against the explicit-Kronecker naive product restated below (the reference's own
 *       second oracle, tests/utils/kronmult_naive.h).
 *
 * Each function cites the reference file:line it follows.  The reference has no
 * third-party arithmetic on this path (KRONMULT_USE_BLAS is off: linear_algebra.hpp:4).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* reference: kronmult_omp/kronmult.hpp:11-15 -- integer power by repeated multiplication,
 * int arithmetic, no overflow check. */
int oracle_pow_int(int base, int exponent)
{
    int acc = 1;
    for (int e = 0; e < exponent; ++e) acc *= base;
    return acc;
}

/* col-major addressing used everywhere in the reference: linear_algebra.hpp:66-69 */
#define CM(r, c, ld) ((size_t)(r) + (size_t)(c) * (size_t)(ld))

#define ORACLE_DEFINE(T, SFX)                                                                        \
    /* reference: linear_algebra.hpp:78-88 -- Mt (compact, ld = n) <- transpose of M (ld = lda) */   \
    static void transpose_##SFX(const T *M, T *Mt, int n, int lda)                                   \
    {                                                                                                \
        for (int r = 0; r < n; ++r)                                                                  \
            for (int c = 0; c < n; ++c) Mt[CM(r, c, n)] = M[CM(c, r, lda)];                          \
    }                                                                                                \
                                                                                                     \
    /* reference: linear_algebra.hpp:102-121 -- Y = X^T * M^T.                                        \
     * X is n x cols (ld n), Y is cols x n (ld cols).  The dot product starts from 0 and             \
     * accumulates k ascending, X operand first (linear_algebra.hpp:113-117). */                      \
    static void multiply_transpose_##SFX(const T *X, int cols, const T *M, int n, int lda, T *Y,     \
                                         T *Mt)                                                      \
    {                                                                                                \
        transpose_##SFX(M, Mt, n, lda);                                                              \
        for (int c = 0; c < cols; ++c)                                                               \
            for (int r = 0; r < n; ++r)                                                              \
            {                                                                                        \
                T dot = (T)0;                                                                        \
                for (int k = 0; k < n; ++k) dot += X[CM(k, c, n)] * Mt[CM(k, r, n)];                 \
                Y[CM(c, r, cols)] = dot;                                                             \
            }                                                                                        \
    }                                                                                                \
                                                                                                     \
    /* reference: kronmult_omp/kronmult.hpp:33-60 -- one batch item.  d passes, matrices             \
     * consumed last -> first (:41), ping-pong between `in` and `ws` (:51), then an element-wise     \
     * atomic add into `out` (:55-59).  `in` and `ws` are clobbered. */                               \
    void oracle_kronmult_##SFX(int d, int n, const T *const *mats, int lda, T *in, int N, T *out,    \
                               T *ws, T *Mt)                                                         \
    {                                                                                                \
        const int cols = N / n;                                                                      \
        T *src = in, *dst = ws;                                                                      \
        for (int j = d - 1; j >= 0; --j)                                                             \
        {                                                                                            \
            multiply_transpose_##SFX(src, cols, mats[j], n, lda, dst, Mt);                           \
            T *t = src;                                                                              \
            src  = dst;                                                                              \
            dst  = t;                                                                                \
        }                                                                                            \
        for (int i = 0; i < N; ++i)                                                                  \
        {                                                                                            \
            _Pragma("omp atomic") out[i] += src[i];                                                  \
        }                                                                                            \
    }                                                                                                \
                                                                                                     \
    /* reference: kronmult_omp/kronmult.hpp:77-104 -- the batch loop: OpenMP over items (:86-94),    \
     * one n*n transpose scratch per thread (:90), item k uses mats[k*d .. k*d+d) (:96). */           \
    void oracle_kronmult_batched_##SFX(int d, int n, const T *const *mats, int lda, T **in,          \
                                       T **out, T **ws, int nb)                                      \
    {                                                                                                \
        const int N = oracle_pow_int(n, d);                                                          \
        _Pragma("omp parallel")                                                                      \
        {                                                                                            \
            T *Mt = (T *)malloc(sizeof(T) * (size_t)n * (size_t)n);                                  \
            _Pragma("omp for") for (int k = 0; k < nb; ++k)                                          \
                oracle_kronmult_##SFX(d, n, mats + (size_t)k * d, lda, in[k], N, out[k], ws[k], Mt); \
            free(Mt);                                                                                \
        }                                                                                            \
    }                                                                                                \
                                                                                                     \
    /* Convenience for the Python harness: same call on slab-allocated data described by             \
     * element offsets instead of pointers (no arithmetic of its own). */                            \
    void oracle_kronmult_batched_slab_##SFX(int d, int n, const T *mat_slab,                         \
                                            const long long *mat_off, int lda, T *in_slab,           \
                                            const long long *in_off, T *out_slab,                    \
                                            const long long *out_off, T *ws_slab,                    \
                                            const long long *ws_off, int nb)                         \
    {                                                                                                \
        const T **mp = (const T **)malloc(sizeof(T *) * (size_t)nb * (size_t)d);                     \
        T **ip = (T **)malloc(sizeof(T *) * (size_t)nb);                                             \
        T **op = (T **)malloc(sizeof(T *) * (size_t)nb);                                             \
        T **wp = (T **)malloc(sizeof(T *) * (size_t)nb);                                             \
        for (size_t i = 0; i < (size_t)nb * (size_t)d; ++i) mp[i] = mat_slab + mat_off[i];           \
        for (int k = 0; k < nb; ++k)                                                                 \
        {                                                                                            \
            ip[k] = in_slab + in_off[k];                                                             \
            op[k] = out_slab + out_off[k];                                                           \
            wp[k] = ws_slab + ws_off[k];                                                             \
        }                                                                                            \
        oracle_kronmult_batched_##SFX(d, n, mp, lda, ip, op, wp, nb);                                \
        free(mp); free(ip); free(op); free(wp);                                                      \
    }                                                                                                \
                                                                                                     \
    /* reference: tests/utils/kronmult_naive.h:48-68 (explicit Kronecker product, row_out =           \
     * row1*size2 + row2), :31-41 (mat-vec, += into output, col ascending), :73-102 (left-to-right   \
     * product of the d factors then one mat-vec).  O(N^2) memory: tiny cases only. */                \
    void oracle_kronmult_naive_##SFX(int d, int n, const T *const *mats, int lda, const T *in,       \
                                     T *out)                                                         \
    {                                                                                                \
        int sz = n, ld = lda;                                                                        \
        const T *K = mats[0];                                                                        \
        T *owned = NULL;                                                                             \
        for (int m = 1; m < d; ++m)                                                                  \
        {                                                                                            \
            const int nsz = sz * n;                                                                  \
            T *Kn = (T *)malloc(sizeof(T) * (size_t)nsz * (size_t)nsz);                              \
            const T *B = mats[m];                                                                    \
            for (int r1 = 0; r1 < sz; ++r1)                                                          \
                for (int r2 = 0; r2 < n; ++r2)                                                       \
                    for (int c1 = 0; c1 < sz; ++c1)                                                  \
                        for (int c2 = 0; c2 < n; ++c2)                                               \
                            Kn[CM(r1 * n + r2, c1 * n + c2, nsz)] = K[CM(r1, c1, ld)] * B[CM(r2, c2, lda)]; \
            free(owned);                                                                             \
            owned = Kn;                                                                              \
            K = Kn; sz = nsz; ld = nsz;                                                              \
        }                                                                                            \
        for (int r = 0; r < sz; ++r)                                                                 \
            for (int c = 0; c < sz; ++c) out[r] += K[CM(r, c, ld)] * in[c];                          \
        free(owned);                                                                                 \
    }                                                                                                \
                                                                                                     \
    /* reference: tests/utils/kronmult_naive.h:108-121 -- sequential batch loop of the above */       \
    void oracle_kronmult_batched_naive_slab_##SFX(int d, int n, const T *mat_slab,                   \
                                                  const long long *mat_off, int lda,                 \
                                                  const T *in_slab, const long long *in_off,         \
                                                  T *out_slab, const long long *out_off, int nb)     \
    {                                                                                                \
        const T **mp = (const T **)malloc(sizeof(T *) * (size_t)d);                                  \
        for (int k = 0; k < nb; ++k)                                                                 \
        {                                                                                            \
            for (int j = 0; j < d; ++j) mp[j] = mat_slab + mat_off[(size_t)k * d + j];               \
            oracle_kronmult_naive_##SFX(d, n, mp, lda, in_slab + in_off[k], out_slab + out_off[k]);  \
        }                                                                                            \
        free(mp);                                                                                    \
    }

ORACLE_DEFINE(double, f64)
ORACLE_DEFINE(float, f32)

/* reference: tests/utils/batch_size.h:8-21 -- the ASGarD-like batch count used by every
 * reference test/bench case: min(memory cap, 2^level * level^min(1, dimension-1)). */
int oracle_compute_batch_size(int degree, int dimension, int grid_level, int nb_distinct_outputs)
{
    const int n = degree, d = dimension;
    const int N = oracle_pow_int(n, d);
    const long long cap_elems = 395000000000LL;
    /* note: like the reference, the products below are evaluated in int before widening */
    const long long cap = (cap_elems - nb_distinct_outputs * N) / (long long)(N * (2 + d * n * n));
    const long long formula =
        (long long)(oracle_pow_int(2, grid_level) * oracle_pow_int(grid_level, d - 1 < 1 ? d - 1 : 1));
    return (int)(cap < formula ? cap : formula);
}
