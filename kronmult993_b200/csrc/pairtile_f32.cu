// pairtile_f32.cu -- fp32 instantiations of the pairtile family (own translation unit: compiled in parallel).
#include "kernel_pairtile.cuh"
#define KRON_PT_TYPE float
namespace kron
{
KRON_PAIRTILE_DEFINE(float)
}
