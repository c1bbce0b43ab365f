// pairtile_f64.cu -- fp64 instantiations of the pairtile family (own translation unit: compiled in parallel).
#include "kernel_pairtile.cuh"
#define KRON_PT_TYPE double
namespace kron
{
KRON_PAIRTILE_DEFINE(double)
}
