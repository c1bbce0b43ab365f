/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (build recipe for oracle/_ref/).
 *
 * This file contains NO algorithm.  It pulls in the UNMODIFIED reference header where it lies
 * (-I$(REF)/kronmult_omp, i.e. /root/reference/kronmult_omp/kronmult.hpp:77-104) and
 * the reference's own naive oracle ($(REF)/tests/utils/kronmult_naive.h:108-121) and gives
 * their template instantiations C names so that Python (ctypes) can call them.
 * The result, oracle/_ref/libkronmult_ref.so, is the real reference: it pins the C restatement in
 * oracle/kronmult_oracle.c and serves as the CPU baseline (`cpu_baseline.kind = "reference"`).
 *
 * The reference header defines a non-inline pow_int (kronmult.hpp:11-15), so it may be included in
 * exactly one translation unit -- this one.
 */
#include <kronmult.hpp>              // reference: kronmult_omp/kronmult.hpp
#include <utils/kronmult_naive.h>    // reference: tests/utils/kronmult_naive.h
#include <vector>

namespace
{
template<typename T>
void batched_slab(int d, int n, T const *mat_slab, long long const *mat_off, int lda, T *in_slab,
                  long long const *in_off, T *out_slab, long long const *out_off, T *ws_slab,
                  long long const *ws_off, int nb)
{
    std::vector<T const *> mats(static_cast<size_t>(nb) * d);
    std::vector<T *> in(nb), out(nb), ws(nb);
    for (size_t i = 0; i < mats.size(); i++) mats[i] = mat_slab + mat_off[i];
    for (int k = 0; k < nb; k++)
    {
        in[k]  = in_slab + in_off[k];
        out[k] = out_slab + out_off[k];
        ws[k]  = ws_slab + ws_off[k];
    }
    kronmult_batched<T>(d, n, mats.data(), lda, in.data(), out.data(), ws.data(), nb);
}

template<typename T>
void naive_slab(int d, int n, T *mat_slab, long long const *mat_off, int lda, T *in_slab,
                long long const *in_off, T *out_slab, long long const *out_off, int nb)
{
    std::vector<T *> mats(static_cast<size_t>(nb) * d);
    std::vector<T *> in(nb), out(nb);
    for (size_t i = 0; i < mats.size(); i++) mats[i] = mat_slab + mat_off[i];
    for (int k = 0; k < nb; k++)
    {
        in[k]  = in_slab + in_off[k];
        out[k] = out_slab + out_off[k];
    }
    // Two defects of the reference's naive oracle are worked around through its own injectable
    // allocator hooks (tests/utils/kronmult_naive.h:73-76), not by editing it:
    //  * :85 asks malloc_f for `size` elements but then fills a size x size matrix (heap overflow
    //    unless the allocator is page-granular like cudaMallocManaged) -> allocate size*size;
    //  * the default pairs new[] with free() (:8-11,:75-76) -> pass a matching deleter.
    kronmult_batched_naive<T>(d, n, mats.data(), lda, in.data(), out.data(), nullptr, nb,
                              [](size_t dim) { return new T[dim * dim]; },
                              [](void *p) { delete[] static_cast<T *>(p); });
}
} // namespace

extern "C"
{
int ref_pow_int(int a, int b) { return pow_int(a, b); }

void ref_kronmult_batched_f64(int d, int n, double const *const *mats, int lda, double **in, double **out,
                              double **ws, int nb)
{
    kronmult_batched<double>(d, n, mats, lda, in, out, ws, nb);
}
void ref_kronmult_batched_f32(int d, int n, float const *const *mats, int lda, float **in, float **out,
                              float **ws, int nb)
{
    kronmult_batched<float>(d, n, mats, lda, in, out, ws, nb);
}
void ref_kronmult_batched_slab_f64(int d, int n, double const *m, long long const *mo, int lda, double *i,
                                   long long const *io, double *o, long long const *oo, double *w,
                                   long long const *wo, int nb)
{
    batched_slab<double>(d, n, m, mo, lda, i, io, o, oo, w, wo, nb);
}
void ref_kronmult_batched_slab_f32(int d, int n, float const *m, long long const *mo, int lda, float *i,
                                   long long const *io, float *o, long long const *oo, float *w,
                                   long long const *wo, int nb)
{
    batched_slab<float>(d, n, m, mo, lda, i, io, o, oo, w, wo, nb);
}
void ref_kronmult_batched_naive_slab_f64(int d, int n, double *m, long long const *mo, int lda, double *i,
                                         long long const *io, double *o, long long const *oo, int nb)
{
    naive_slab<double>(d, n, m, mo, lda, i, io, o, oo, nb);
}
void ref_kronmult_batched_naive_slab_f32(int d, int n, float *m, long long const *mo, int lda, float *i,
                                         long long const *io, float *o, long long const *oo, int nb)
{
    naive_slab<float>(d, n, m, mo, lda, i, io, o, oo, nb);
}
}

// reference: tests/utils/batch_size.h:8-21 (needs pow_int declared first, which the header above did)
#include <algorithm>
#include <utils/batch_size.h>
extern "C" int ref_compute_batch_size(int degree, int dimension, int level, int nb_distinct)
{
    return compute_batch_size(degree, dimension, level, nb_distinct);
}
