// pairpass_f32.cu -- float instantiations of the pairtile multi-pass route (own translation unit).
#include "kernel_pairpass.cuh"
namespace kron
{
template<>
cudaError_t run_pairpass<float>(int sms, int d, int n, const float *const *A, int lda, float *const *in, float *const *out,
                              int nb, cudaStream_t st, std::atomic<long long> &launches, float *const *scratch)
{
    switch (n)
    {
#define KRON_PP(NN) case NN: return run_pairpass_n<float, NN>(sms, d, A, lda, in, out, nb, st, launches, scratch);
        KRON_PP(5) KRON_PP(6) KRON_PP(7) KRON_PP(8) KRON_PP(9) KRON_PP(10)
#undef KRON_PP
    default: return cudaErrorNotSupported;
    }
}
} // namespace kron
