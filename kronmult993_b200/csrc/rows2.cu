// rows2.cu -- instantiations of the d = 2 lane-per-fibre kernel (own translation unit: compiled in parallel).
#define KRON_ROWS2_DEFINE
#include "kernel_rows2.cuh"
namespace kron
{
KRON_ROWS2_DEFINE_RUN(double)
KRON_ROWS2_DEFINE_RUN(float)
}
