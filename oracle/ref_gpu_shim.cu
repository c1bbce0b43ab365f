/*
 * oracle/ref_gpu_shim.cu -- TEST/BASELINE INFRASTRUCTURE ONLY (build recipe for oracle/_ref/).
 *
 * Contains NO algorithm.  oracle/Makefile compiles the UNMODIFIED reference CUDA source where it
 * lies ($(REF)/kronmult_gpu/kronmult.cu, launcher at :173-197) for sm_100a next to this shim, which
 * only renames its two exported specialisations (kronmult.cu:202-211, :216-224) to C symbols.
 * The resulting oracle/_ref/libkronmult_refgpu.so is the "reference kernel recompiled for B200"
 * timing baseline.  It is never a parity oracle: its float path has a divergent-barrier hazard
 * (kronmult.cu:59,74).  Linked -Bsymbolic so that its kronmult_batched<T> can coexist in one
 * process with the product library, which exports the same C++ symbols by design.
 */
#include <kronmult.cuh> // reference: kronmult_gpu/kronmult.cuh

extern "C"
{
int refgpu_kronmult_batched_f64(int d, int n, double const *const *mats, int lda, double **in,
                                double **out, double **ws, int nb)
{
    return static_cast<int>(kronmult_batched<double>(d, n, mats, lda, in, out, ws, nb));
}
int refgpu_kronmult_batched_f32(int d, int n, float const *const *mats, int lda, float **in, float **out,
                                float **ws, int nb)
{
    return static_cast<int>(kronmult_batched<float>(d, n, mats, lda, in, out, ws, nb));
}
}
