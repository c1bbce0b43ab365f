// kronmult.cuh -- C++ drop-in boundary of the B200-native kronmult library.
//
// This header declares exactly the two entry points that the reference's CUDA flavour exposes
// (project-asgard/kronmult993, kronmult_gpu/kronmult.cuh:10 and :28-32), with the same names,
// parameter types and return type, so that their mangled symbols
//     _Z7pow_intii
//     _Z16kronmult_batchedIdE9cudaErroriiPKPKT_iPPS1_S7_S7_i   (T = double)
//     _Z16kronmult_batchedIfE9cudaErroriiPKPKT_iPPS1_S7_S7_i   (T = float)
// are the ones a consumer such as ASGarD already links against.  Only the float and double
// specialisations exist (as in kronmult_gpu/kronmult.cu:202-224); any other T is a link error.
// The implementation behind them (kronmult993_b200/csrc) is new; see DESIGN.md.
#pragma once
#include <cuda_runtime.h>

// n^p in int arithmetic (reference: kronmult_gpu/kronmult.cu:11-15).  Exported because the
// reference's own test programs call it (tests/utils/batch_size.h:13, tests/kronmult_test_gpu.cpp:20).
__host__ int pow_int(int const number, int const power);

// For every k in [0, nb_batch):
//     output_batched[k][0:N] += ( A[k,0] (x) A[k,1] (x) ... (x) A[k,d-1] ) * input_batched[k][0:N]
// with d = matrix_count, n = matrix_size, N = n^d, A[k,j] = matrix_list_batched[k*d + j] an n x n
// column-major matrix with leading dimension matrix_stride; the last factor acts on the fastest
// index of the vector.
//
// Contract (same as the reference, kronmult_gpu/kronmult.cuh:12-26):
//  * every array, and the pointer arrays themselves, must be device-accessible (cudaMalloc or
//    managed memory);
//  * output pointers may repeat: the accumulation is thread-safe for any aliasing pattern;
//  * input_batched[k] and workspace_batched[k] MAY be overwritten (this implementation leaves them
//    untouched whenever N fits on-chip; workspace is never needed);
//  * sizes are not validated; the call blocks until the result is visible and returns the CUDA
//    status (cudaSuccess, or the launch/synchronisation error).
// Limits of this implementation (the reference accepts any size and then overflows its int size_input silently,
// kronmult.cu:180): matrix_size <= 32 and matrix_size^matrix_count < 2^31, else cudaErrorInvalidValue.  The
// reference's own envelope is matrix_size <= 10, matrix_count <= 6 (tests/kronmult_fullbench_gpu.cpp:70-74).
template<typename T>
__host__ cudaError kronmult_batched(int const matrix_count, int const matrix_size,
                                    T const *const matrix_list_batched[], int const matrix_stride,
                                    T *input_batched[], T *output_batched[], T *workspace_batched[],
                                    int const nb_batch);

// ---- extension (not in the reference): read-only, shareable inputs -------------------------------------------------
// Same operation, but input_batched[k] is NEVER written, so entries may repeat (an ASGarD-style caller keeps one copy
// of every x_j instead of one per batch item, kronmult_gpu/kronmult.cuh:23 / README.md:49 force the copies today).
// workspace_batched[k] must be a distinct n^d-element scratch vector per item when
// kronmult_b200_needs_workspace(matrix_count, matrix_size, sizeof(T)) (include/kronmult_b200.h) returns 1, and may be
// nullptr otherwise.  float and double only; blocking like kronmult_batched.
template<typename T>
__host__ cudaError kronmult_batched_const(int const matrix_count, int const matrix_size,
                                          T const *const matrix_list_batched[], int const matrix_stride,
                                          T const *const input_batched[], T *output_batched[],
                                          T *workspace_batched[], int const nb_batch);
