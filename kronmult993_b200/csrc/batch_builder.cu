// batch_builder.cu -- device-side construction of the pointer arrays of an ASGarD-style batch.
//
// No reference counterpart: the reference library only CONSUMES the four pointer arrays (kronmult_gpu/kronmult.cu:
// 148-151); its harness documents where they come from -- ASGarD (README.md:8, tests/README.md:33-37) builds one
// batch item per (row element, column element, operator term), tests/utils/batch_size.h:16-20 -- and fills them on
// the host, one allocation per vector (tests/utils/utils_gpu.h:58-65).  SURVEY.md section 8(f) rank 3 asks for the
// caller side on the device: a sparse-grid element is a d-tuple of 1-D cell indices; the term t of the operator
// is the Kronecker product of d one-dimensional coefficient matrices C[t][dim] (column-major, leading dimension
// lda, made of n x n blocks); the item (i, j, t) multiplies the blocks (cell_i[dim], cell_j[dim]) of those d
// matrices with the column element's vector x_j and adds into the row element's vector y_i:
//     A[k*d + dim] = C[t][dim] + n * cell_i[dim] + n * cell_j[dim] * lda        (a window, nothing is copied)
//     in[k]        = x + j * n^d          (shared by every row element and term: use kronmult_batched_const_*)
//     out[k]       = y + i * n^d          (repeats for the (col1-col0)*nterms consecutive items of row i)
// Items are ordered row-major (i, then j, then t), so the runs of equal output pointers are contiguous and every
// kernel family sums them on chip.  One thread per item; 8 * (d + 2) bytes written per item.
#include "../../include/kronmult_b200.h"
#include "common.cuh"

namespace kron
{

template<typename T>
__global__ void build_asgard_batch_kernel(int d, int n, int lda, const int *__restrict__ cells,
                                          const T *const *__restrict__ coeff, int nterms, int row0, int col0, int ncols,
                                          long long nb, long long N, const T *x, T *y, const T **__restrict__ A,
                                          const T **__restrict__ in, T **__restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    const int t       = (int)(k % nterms);
    const long long q = k / nterms;
    const int j       = col0 + (int)(q % ncols);
    const int i       = row0 + (int)(q / ncols);
    in[k]  = x + (long long)j * N;
    out[k] = y + (long long)i * N;
    for (int dim = 0; dim < d; ++dim)
    {
        const long long r = cells[(long long)i * d + dim], c = cells[(long long)j * d + dim];
        A[k * d + dim] = coeff[(long long)t * d + dim] + n * r + n * c * (long long)lda;
    }
}

template<typename T>
static int build_batch(int d, int n, int lda, const int *cells, const T *const *coeff, int nterms, int row0, int row1,
                       int col0, int col1, const T *x, T *y, const T **A, const T **in, T **out, long long *nb_out,
                       cudaStream_t st)
{
    if (d < 1 || n < 1 || lda < n || nterms < 1 || row1 < row0 || col1 < col0 || row0 < 0 || col0 < 0)
        return (int)cudaErrorInvalidValue;
    const long long nb = (long long)(row1 - row0) * (col1 - col0) * nterms;
    if (nb_out) *nb_out = nb;
    if (nb >= (1LL << 31)) return (int)cudaErrorInvalidValue; // nb_batch of kronmult_batched is an int
    if (nb == 0) return 0;
    if (!cells || !coeff || !x || !y || !A || !in || !out) return (int)cudaErrorInvalidValue;
    long long N = 1;
    for (int i = 0; i < d; ++i)
    {
        N *= n;
        if (N >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    }
    const int threads = 256;
    const long long blocks = (nb + threads - 1) / threads;
    build_asgard_batch_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(d, n, lda, cells, coeff, nterms, row0, col0,
                                                                      col1 - col0, nb, N, x, y, A, in, out);
    return (int)cudaGetLastError();
}

} // namespace kron

extern "C"
{
int kronmult_build_batch_f64(int d, int n, int lda, const int *cells, const double *const *coeff, int nterms, int row0,
                             int row1, int col0, int col1, const double *x, double *y, const double **A,
                             const double **in, double **out, long long *nb, void *stream)
{
    return kron::build_batch<double>(d, n, lda, cells, coeff, nterms, row0, row1, col0, col1, x, y, A, in, out, nb,
                                     static_cast<cudaStream_t>(stream));
}
int kronmult_build_batch_f32(int d, int n, int lda, const int *cells, const float *const *coeff, int nterms, int row0,
                             int row1, int col0, int col1, const float *x, float *y, const float **A, const float **in,
                             float **out, long long *nb, void *stream)
{
    return kron::build_batch<float>(d, n, lda, cells, coeff, nterms, row0, row1, col0, col1, x, y, A, in, out, nb,
                                    static_cast<cudaStream_t>(stream));
}
}
