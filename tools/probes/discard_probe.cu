// Does discard.global.L2 on B200 drop dirty L2 lines without writing them back?  (round-2 design probe for the
// fused multi-pass route: the intermediate vector is written once, read once, and never needed again.)
// Kernels, each over the same 48 MiB buffer (fits L2), run under ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum:
//   fill      : write the buffer (dirty lines in L2)
//   read      : read it back (sum)                      -> expect ~0 DRAM reads if it stayed in L2
//   read_disc : read it back, then discard every line   -> lines dropped?
//   evict     : write a different 256 MiB buffer        -> forces the 48 MiB out: DRAM writes show whether dirty data was still there
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill(double *p, size_t n, double v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v + i;
}
__global__ void readk(const double *p, size_t n, double *sink, int disc)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        s += p[i];
        if (disc && (i % 16 == 0)) // one thread per 128-byte line, after its own read (the others of the line are in the same warp instruction)
        {
            __syncwarp();
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + i) : "memory");
        }
    }
    if (s == 1.2345) *sink = s;
}
int main()
{
    const size_t n = (48u << 20) / 8, m = (256u << 20) / 8;
    double *a, *b, *sink;
    cudaMalloc(&a, n * 8); cudaMalloc(&b, m * 8); cudaMalloc(&sink, 8);
    for (int rep = 0; rep < 2; ++rep)
    {
        const int disc = rep;
        fill<<<148 * 4, 256>>>(b, m, 0.0);          // flush L2 with other data
        fill<<<148 * 4, 256>>>(a, n, 1.0);          // dirty lines of a
        readk<<<148 * 4, 256>>>(a, n, sink, disc);  // read (and discard)
        fill<<<148 * 4, 256>>>(b, m, 2.0);          // evict a: write-backs appear here (or in the kernels above)
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
