/* kronmult_b200.h -- C ABI of the B200-native kronmult library (libkronmult_b200.so).
 *
 * The reference (project-asgard/kronmult993) has no C ABI: its GPU flavour exports two C++ template
 * specialisations (kronmult_gpu/kronmult.cu:202-211 double, :216-224 float; declared at
 * kronmult_gpu/kronmult.cuh:28-32) and pow_int (kronmult.cu:11-15, kronmult.cuh:10).  The functions
 * below are what an FFI for that path binds; the C++ specialisations in include/kronmult.cuh are thin
 * wrappers over them.  Plain pointers and sizes only; every function returns a cudaError_t value as
 * int (0 = cudaSuccess) and never throws.
 *
 * Common arguments (identical meaning to kronmult.cuh:12-26):
 *   d    matrix_count          number of Kronecker factors
 *   n    matrix_size           each factor is n x n, column-major
 *   A    matrix_list_batched   DEVICE array of nb*d DEVICE pointers, item k uses A[k*d .. k*d+d)
 *   lda  matrix_stride         leading dimension of every factor (>= n)
 *   in   input_batched         DEVICE array of nb DEVICE pointers to n^d-element vectors (may be clobbered)
 *   out  output_batched        DEVICE array of nb DEVICE pointers; out[k] += kron(A_k) * in[k]; may repeat
 *   ws   workspace_batched     DEVICE array of nb DEVICE pointers (may be clobbered; may be NULL here:
 *                              this implementation never dereferences it)
 *   nb   nb_batch              number of batch items (0 is a no-op)
 * Limits: n <= 32 and n^d < 2^31 (cudaErrorInvalidValue otherwise; the reference's envelope is n <= 10, d <= 6).
 */
#ifndef KRONMULT_B200_H
#define KRONMULT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* replaces pow_int, kronmult_gpu/kronmult.cuh:10 (kronmult.cu:11-15) */
int kronmult_pow_int(int number, int power);

/* replace kronmult_batched<double> / <float>, kronmult_gpu/kronmult.cuh:28-32
 * (kronmult.cu:202-211 / :216-224): blocking, legacy default stream, returns after
 * cudaDeviceSynchronize() like kronmult.cu:196. */
int kronmult_batched_f64(int d, int n, const double *const *A, int lda, double **in, double **out,
                         double **ws, int nb);
int kronmult_batched_f32(int d, int n, const float *const *A, int lda, float **in, float **out,
                         float **ws, int nb);

/* stream-ordered variants of the same call (no reference counterpart: the reference is blocking,
 * kronmult.cu:191-196).  `stream` is a cudaStream_t; no host synchronisation is performed. */
int kronmult_batched_f64_async(int d, int n, const double *const *A, int lda, double **in, double **out,
                               double **ws, int nb, void *stream);
int kronmult_batched_f32_async(int d, int n, const float *const *A, int lda, float **in, float **out,
                               float **ws, int nb, void *stream);

/* Read-only-input variants (no reference counterpart).  The reference's contract lets
 * the call clobber `input` (kronmult.cuh:23, kronmult.cu:115-121), which forces an ASGarD-style caller to give every
 * batch item its own copy of its input vector.  Here `in[k]` is never written, so items may share input vectors
 * (in[k] may repeat, like out[k]): a vector that many items multiply is fetched from HBM once and served from L2.
 * `ws[k]` must be a distinct n^d-element scratch vector per item when kronmult_b200_needs_workspace(d, n, sizeof(T))
 * returns 1 (vectors too long for shared memory take a multi-pass route through global memory: the first pass writes
 * the scratch vector, the rest works in place there); otherwise `ws` is not dereferenced and may be NULL.  With a
 * needed but missing `ws` the call fails with cudaErrorInvalidValue.  Everything else as kronmult_batched_*. */
int kronmult_batched_const_f64(int d, int n, const double *const *A, int lda, const double *const *in, double **out,
                               double **ws, int nb);
int kronmult_batched_const_f32(int d, int n, const float *const *A, int lda, const float *const *in, float **out,
                               float **ws, int nb);
int kronmult_batched_const_f64_async(int d, int n, const double *const *A, int lda, const double *const *in,
                                     double **out, double **ws, int nb, void *stream);
int kronmult_batched_const_f32_async(int d, int n, const float *const *A, int lda, const float *const *in, float **out,
                                     float **ws, int nb, void *stream);
/* 1 if the read-only-input entry points need `ws` for this shape, 0 if not, -1 for invalid arguments (the answer
 * follows knob 12 of kronmult_b200_set_tuning for double-precision n = 8, d = 5 / 6) */
int kronmult_b200_needs_workspace(int d, int n, int elem_size);

/* Host-buffer entry points with the signature of the reference's CPU flavour
 * (kronmult_omp/kronmult.hpp:77-80): every pointer array and every pointee lives in HOST memory.
 * The library stages chunks of items through pinned buffers, runs the device path on `device`
 * (-1 = current device) and writes the accumulated outputs back; input/workspace are not modified. */
int kronmult_batched_host_f64(int d, int n, const double *const *A, int lda, double **in, double **out,
                              double **ws, int nb, int device);
int kronmult_batched_host_f32(int d, int n, const float *const *A, int lda, float **in, float **out,
                              float **ws, int nb, int device);

/* Multi-GPU partitioner (no reference counterpart; BASELINE.json north_star): assign every item to one
 * of n_ranks owners such that all items sharing an output pointer get the same owner and the item
 * counts are balanced (longest-processing-time greedy over output groups).  `out` is a HOST array of
 * nb output pointers (any integer key works); owner[k] in [0, n_ranks) is written for every item.
 * Groups larger than `split_threshold` items (0 = never) are instead split evenly over all ranks and
 * flagged in needs_reduce[k] = 1: their partial outputs must be summed across ranks afterwards. */
int kronmult_partition_by_output(const void *const *out, int nb, int n_ranks, long long split_threshold,
                                 int *owner, unsigned char *needs_reduce);

/* Multi-GPU execution of one rank's shard, with the ONE collective of the design (no reference counterpart;
 * BASELINE.json north_star).  One process per GPU.  A kronmult_comm wraps an NCCL communicator over the ranks
 * (NCCL is loaded lazily with dlopen: the single-GPU entry points have no NCCL dependency):
 *   kronmult_comm_unique_id   rank 0 obtains the 128-byte ncclUniqueId and hands it to the other ranks by any
 *                             means (bench.py / the tests broadcast it with torch.distributed);
 *   kronmult_comm_create      collective over all ranks (ncclCommInitRank on the CURRENT device);
 *   kronmult_comm_adopt       wrap a communicator the caller already has (ncclComm_t; not destroyed by _destroy);
 *   kronmult_comm_destroy.
 * kronmult_batched_sharded_*: like kronmult_batched_*_async on this rank's items (A, in, out, nb as above), plus
 * `shared_out`: HOST array of n_shared DEVICE pointers = this rank's own copies of the output vectors whose items
 * were split across ranks (needs_reduce of kronmult_partition_by_output) -- the same vectors in the same order on
 * every rank -- and `owner` (HOST array, n_shared ranks, or NULL).  Items that write a shared vector accumulate
 * into zero-initialised scratch; the scratch is summed over all ranks with ONE ncclAllReduce over NVLink and the
 * total is added into this rank's copy of every shared vector with owner[j] == rank (NULL: into every rank's copy).
 * Every rank must make the call (also with nb = 0).  Stream-ordered; returns cudaErrorNotSupported if more than one
 * rank is asked for and libnccl.so.2 cannot be loaded.
 * kronmult_comm_last_collective_ms: device time of the most recent all-reduce (cudaEvents on the stream; waits for
 * it) and the number of collectives issued so far. */
typedef struct kronmult_comm kronmult_comm;
int kronmult_comm_unique_id(void *id128);
int kronmult_comm_create(const void *id128, int world, int rank, kronmult_comm **comm);
int kronmult_comm_adopt(void *nccl_comm, int world, int rank, kronmult_comm **comm);
int kronmult_comm_destroy(kronmult_comm *comm);
int kronmult_comm_last_collective_ms(kronmult_comm *comm, float *ms, long long *count);
int kronmult_batched_sharded_f64(int d, int n, const double *const *A, int lda, double **in, double **out, double **ws,
                                 int nb, double *const *shared_out, int n_shared, const int *owner, kronmult_comm *comm,
                                 void *stream);
int kronmult_batched_sharded_f32(int d, int n, const float *const *A, int lda, float **in, float **out, float **ws,
                                 int nb, float *const *shared_out, int n_shared, const int *owner, kronmult_comm *comm,
                                 void *stream);

/* Device-side batch builder for ASGarD-style callers (no reference counterpart; SURVEY.md section 8(f) rank 3: the
 * reference's harness fills the pointer arrays on the host, tests/utils/utils_gpu.h:58-65; the batch shape is
 * ASGarD's, tests/utils/batch_size.h:16-20).  Builds the pointer arrays of the batch
 *     { (i, j, t) : row element i in [row0,row1), column element j in [col0,col1), term t in [0,nterms) }
 * ordered (i, j, t) so that equal output pointers are consecutive:
 *     A[k*d + dim] = coeff[t*d + dim] + n*cells[i*d+dim] + n*cells[j*d+dim]*lda    (an n x n window, nothing copied)
 *     in[k] = x + j*n^d  (shared by rows and terms -> call kronmult_batched_const_*),   out[k] = y + i*n^d.
 * cells: DEVICE array [num_elements*d] of 1-D cell indices; coeff: DEVICE array [nterms*d] of DEVICE pointers to the
 * one-dimensional coefficient matrices (column-major, leading dimension lda); A / in / out: DEVICE arrays of
 * nb*d / nb / nb pointers with nb = (row1-row0)*(col1-col0)*nterms (also returned in *nb).  Stream-ordered. */
int kronmult_build_batch_f64(int d, int n, int lda, const int *cells, const double *const *coeff, int nterms, int row0,
                             int row1, int col0, int col1, const double *x, double *y, const double **A,
                             const double **in, double **out, long long *nb, void *stream);
int kronmult_build_batch_f32(int d, int n, int lda, const int *cells, const float *const *coeff, int nterms, int row0,
                             int row1, int col0, int col1, const float *x, float *y, const float **A, const float **in,
                             float **out, long long *nb, void *stream);

/* Batch / aliasing planner (no reference counterpart; BASELINE.json north_star).  Every kernel sums runs of
 * consecutive items that share an output pointer on chip; a plan sorts the batch by output pointer on the
 * device (stable, so equal pointers keep their batch order) and keeps sorted copies of the three pointer
 * arrays when that at least halves the number of runs.  The pointer arrays given to kronmult_plan_create_*
 * (DEVICE arrays, same meaning as above) must stay alive and unchanged while the plan is used; the vectors
 * and factors they point to may change freely between executions (ASGarD: same pointers every time step).
 * create blocks until the plan is built; execute is stream-ordered like the _async entry points.  The blocking
 * entry points above build and cache such plans on their own (knob 1 of kronmult_b200_set_tuning, default on). */
typedef struct kronmult_plan kronmult_plan;
int kronmult_plan_create_f64(int d, int n, const double *const *A, int lda, double **in, double **out, int nb,
                             void *stream, kronmult_plan **plan);
int kronmult_plan_create_f32(int d, int n, const float *const *A, int lda, float **in, float **out, int nb,
                             void *stream, kronmult_plan **plan);
int kronmult_plan_execute(const kronmult_plan *plan, void *stream);
/* runs of equal consecutive output pointers before / after sorting; permuted = 1 if the sorted order is used */
int kronmult_plan_stats(const kronmult_plan *plan, long long *runs_before, long long *runs_after, int *permuted);
int kronmult_plan_destroy(kronmult_plan *plan);
/* implicit plans of the blocking entry points: cache hits / plans built so far in this process */
long long kronmult_b200_plan_cache_hits(void);
long long kronmult_b200_plan_cache_builds(void);

/* Introspection used by the test-suite and bench.py. */
const char *kronmult_b200_version(void);
/* number of kernels this library has launched in this process so far */
long long kronmult_b200_launch_count(void);
/* name of the kernel family chosen by the most recent call on this thread ("tiny", "generic", ...) */
const char *kronmult_b200_last_path(void);
/* 0 = automatic dispatch; otherwise force a kernel family (the kron::Path codes in kronmult993_b200/csrc/common.cuh).
 * Unsupported combinations make the next call return cudaErrorInvalidValue.  For tests. */
int kronmult_b200_force_path(int path);
/* knobs.  0: regtile operand staging (0 = TMA into shared memory, 1 = L1 prefetch; development).
 *         1: implicit planning in the blocking entry points (1 = on, default; 0 = off).
 *         2: ablation variants of the n=4, d=5 kernel (only in builds with -DKRON_WSPEC5_EXPERIMENTS).
 *         3: largest vector, in KiB, that the shape-agnostic path keeps resident in shared memory in one pass
 *            (default 56); longer vectors take the tiled multi-pass route, which works in place in `in`.
 *         4: largest vector, in KiB, that the pairtile family keeps resident (default 227 = whatever fits); smaller
 *            values send long vectors through the pairtile multi-pass route (development: resident measured faster).
 *         5: variant of the n=4, d=5 kernel (development; only in builds with -DKRON_SYM5_VARIANTS).
 *         6: multi-pass routes (vectors beyond shared memory): MiB of vectors per chunk of items whose passes run back
 *            to back so that the intermediate stays in L2 (0 = pass by pass over the whole batch; -1 = automatic,
 *            the default: 32 MiB for routes of three or more passes, where it measured 29 % faster, else 0).
 *         7: 1 = drop the dead intermediate from L2 with discard.global.L2 after a chunk's last pass (default 0).
 *         8: internal streams the chunks of knob 6 are spread over (default 3; 1 = the caller's stream only).
 *         9: 1 (default) = the one-thread-per-item kernels stage items of 128 / 256 / 512 bytes through shared memory with
 *            coalesced 16-byte cp.async copies, and five more fp64 shapes with element-wise copies; 2 = the former
 *            only; 0 = every thread loads its own item (round-1 kernel).
 *        10: 1 / 2 = single-precision n = 4, d = 5 runs on the half-warp-per-item kernel (kernel_symh.cuh, 8 / 12 CTAs
 *            per SM); 0 (default, measured faster) = on the warp-per-item kernel of kernel_sym5.cuh.
 *        11: warp-per-item DMMA kernel (n = 5..8, d = 2, 3, both precisions): 1 (default) = the shapes where it measured
 *            faster, 2 = every shape it supports, 0 = off (tiny / pair-tile kernels).
 *        12: double-precision n = 8, d = 5 / 6 (vectors of 256 KiB / 2 MiB): 2 (default) = one persistent kernel whose
 *            intermediate lives in a library-owned 64 MiB ring that stays in L2 (kernel_dmma_l2.cuh; `in` is only
 *            read, no scratch vectors needed; successive calls on one device share the ring and are ordered by an
 *            event, also across streams), 1 = the same without the L2 evict-first hint on the input loads,
 *            0 = the multi-kernel routes that work in place through `in` (round-1 behaviour).
 *        13 / 14: ring slots in use (2..8, default 6) and queue lag in blocks (default 3) of knob 12's kernel for d = 6.
 *        15 / 16: the same for d = 5 (2..32, default 24; default 12).
 *        17: lane-per-fibre kernel for d = 2 (kernel_rows2.cuh, n = 5..10, both precisions): 1 (default) = the shapes where
 *            it measured faster (double n = 6, 9, 10; single n = 7 .. 10), 2 = every shape it is built for, 0 = off.
 *        18: its pipeline depth: 0 / 1 = two / three stages of items per warp, -1 (default) = per shape. */
int kronmult_b200_set_tuning(int knob, int value);

#ifdef __cplusplus
}
#endif
#endif /* KRONMULT_B200_H */
