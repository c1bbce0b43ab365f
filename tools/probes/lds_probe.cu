// How many shared-memory wavefronts does a warp-wide load cost when GROUPS of lanes read the same address?
// (round-2 design probe for kernel_rows2.cuh: lane (s, j) of item slot s reads a factor column with a slot-uniform
// 128-bit load, i.e. a warp instruction with IPW distinct addresses.)  One CTA of 8 warps on one SM issues `iters` x 16
// independent loads per warp (XOR-accumulated, so that arithmetic does not bound the loop); cycles per warp instruction at saturation = wavefronts per instruction (the pipe
// delivers one wavefront per clock).  Patterns: lane l reads address (l / G) * stride + (l % G) * lane_step.
#include <cstdio>
#include <cuda_runtime.h>

template<int WIDTH> // bytes per lane: 8 or 16
__global__ void __launch_bounds__(256) probe(const int G, const int stride, const int lane_step, const int iters, long long *clk,
                                             double *sink)
{
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 48 * 1024 / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (lane / G) * stride + (lane % G) * lane_step;
    unsigned a0 = 0, a1 = 0; // two independent XOR chains: the loop is bound by the shared-memory pipe, not by arithmetic
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 16; ++u)
        {
            const unsigned a = base + ((it * 16 + u) & 15) * 1024; // 16 rotating rows: same bank pattern, different data
            if constexpr (WIDTH == 16)
            {
                unsigned x, y, z, w;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
                a0 ^= x ^ y; a1 ^= z ^ w;
            }
            else
            {
                unsigned x, y;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
                a0 ^= x; a1 ^= y;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) *clk = t1 - t0;
    if ((a0 ^ a1) == 0x12345u) *sink = a0;
}

int main()
{
    long long *clk; double *sink;
    cudaMalloc(&clk, 8); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    struct P { int width, G, stride, step; const char *what; };
    const P pats[] = {
        {16, 32, 0, 0, "128-bit, all lanes one address"},
        {16, 1, 16, 0, "128-bit, every lane its own consecutive 16 B (512 B)"},
        {16, 16, 16, 0, "128-bit, 2 groups of 16 lanes, adjacent 16 B"},
        {16, 16, 2416, 0, "128-bit, 2 groups of 16, stride 2416 B"},
        {16, 10, 2416, 0, "128-bit, rows2 n=10: groups of 10 lanes, slot stride 2416 B"},
        {16, 10, 16, 0, "128-bit, groups of 10 lanes, adjacent 16 B"},
        {16, 8, 16, 0, "128-bit, 4 groups of 8 lanes, adjacent 16 B"},
        {16, 8, 2416, 0, "128-bit, 4 groups of 8, stride 2416 B"},
        {16, 5, 16, 0, "128-bit, groups of 5 lanes (6 + 2 lanes), adjacent 16 B"},
        {16, 5, 2416, 0, "128-bit, groups of 5 lanes, stride 2416 B"},
        {16, 5, 1232, 0, "128-bit, groups of 5 lanes, stride 1232 B"},
        {16, 4, 16, 0, "128-bit, 8 groups of 4 lanes, adjacent 16 B (128 B)"},
        {16, 2, 16, 0, "128-bit, 16 groups of 2 lanes, adjacent 16 B (256 B)"},
        {16, 9, 2000, 0, "128-bit, rows2 n=9: groups of 9 lanes, slot stride 2000 B"},
        {16, 9, 2064, 0, "128-bit, groups of 9 lanes, slot stride 2064 B"},
        {8, 32, 0, 0, "64-bit, all lanes one address"},
        {8, 1, 8, 0, "64-bit, every lane its own consecutive 8 B (256 B)"},
        {8, 10, 2416, 0, "64-bit, groups of 10 lanes, stride 2416 B"},
        {8, 9, 2000, 0, "64-bit, groups of 9 lanes, stride 2000 B"},
        {8, 5, 1232, 0, "64-bit, groups of 5 lanes, stride 1232 B"},
        {8, 1, 80, 0, "64-bit, lane stride 80 B (rows2 n=10 row reads as 64-bit)"},
        {8, 1, 72, 0, "64-bit, lane stride 72 B (rows2 n=9 row reads)"},
        {16, 1, 80, 0, "128-bit, lane stride 80 B (rows2 n=10 row reads)"},
    };
    for (const P &p : pats)
    {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep)
        {
            if (p.width == 16) probe<16><<<1, 256, 64 * 1024>>>(p.G, p.stride, p.step, iters, clk, sink);
            else probe<8><<<1, 256, 64 * 1024>>>(p.G, p.stride, p.step, iters, clk, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        }
        printf("{\"pattern\": \"%s\", \"cycles_per_warp_load\": %.2f}\n", p.what, (double)h / (iters * 16.0 * 8));
    }
    return 0;
}
