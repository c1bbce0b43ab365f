// microbench.cu -- measures the B200 roofline denominators MEASURED_PEAKS.json does not carry and
// the primitive costs that decide the kernel design ("P_fp64, P_fp32 measured on the
// box by a DFMA/FFMA (and DMMA) saturation micro-kernel").  Standalone program; prints one JSON
// object per line.  Not part of the library.
//   dfma / ffma / dmma   sustained FMA-pipe and FP64 tensor-pipe peaks (TFLOP/s), ~1 s each so that
//                        the clocks settle under load
//   read_bw              read-only streaming bandwidth (128-bit loads)
//   red_f64_*            RED.ADD.F64 throughput: coalesced / 128-byte strided lanes, HBM-sized array
//   bulkred_f64          cp.reduce.async.bulk.add.f64 shared->global throughput
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <string>
#include <vector>

#define CK(x)                                                                                           \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

template<typename T, int ILP>
__global__ void __launch_bounds__(256) fma_kernel(T *out, int iters, T a, T b)
{
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == T(-12345.678)) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == -12345.678) out[0] = s;
}

// Dependent-issue behaviour of the FP64 tensor pipe: every warp runs CH independent chains in which the D fragment of one
// DMMA is the B fragment of the next (the chaining of kernel_dmma.cuh).  One block of `blockDim.x / 32` warps per SM.
// clocks[0] = cycles of block 0.
template<int CH>
__global__ void dmma_chain_kernel(double *out, long long *clocks, int iters)
{
    double a = 1e-3 * (threadIdx.x & 31), x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = 1e-3 * (threadIdx.x + i); y[i] = 0.0; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                         : "=d"(x[i]), "=d"(y[i]) : "d"(a), "d"(x[i]), "d"(0.0), "d"(0.0));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + y[i];
    if (s == -12345.678) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}
template<int CH>
static void dmma_chain(int sms, double *sink, long long *clk, int warps)
{
    const int iters = 1 << 15;
    dmma_chain_kernel<CH><<<sms, warps * 32>>>(sink, clk, iters);
    CK(cudaDeviceSynchronize());
    dmma_chain_kernel<CH><<<sms, warps * 32>>>(sink, clk, iters);
    CK(cudaDeviceSynchronize());
    long long c = 0;
    CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
    const double per_warp = (double)c / ((double)iters * CH);           // cycles between DMMAs of one warp
    const double per_smsp = per_warp / ((warps + 3) / 4 > 0 ? (double)((warps + 3) / 4) : 1.0); // ... of one SM sub-partition
    printf("{\"bench\": \"dmma_chain\", \"warps_per_sm\": %d, \"chains_per_warp\": %d, \"cycles_per_dmma_per_warp\": %.1f, "
           "\"cycles_per_dmma_per_smsp\": %.1f}\n", warps, CH, per_warp, per_smsp);
}

// DFMA and DMMA in ONE instruction stream: per trip 8 DMMA (8 x 256 MACs) and NF x 8 independent DFMA per thread.
// If the FP64 tensor pipe and the FP64 FMA pipe are separate units the rates add up; if they share the datapath
// the time is the sum of the two.
template<int NF>
__global__ void __launch_bounds__(256) mixed_kernel(double *out, int iters, double fa, double fb)
{
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[8][2], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; acc[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[(i + f) & 7] = acc[(i + f) & 7] * fa + fb;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + acc[i];
    if (s == -12345.678) out[0] = s;
}

template<int NF>
static void mixed(const char *name, int sms, double *sink);

__global__ void __launch_bounds__(256) read_kernel(const int4 *__restrict__ p, size_t n16, int *sink)
{
    int acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride)
    {
        int4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n16; i += stride) acc += __ldcs(p + i).x;
    if (acc == 0x7fffffff) *sink = acc;
}

// every thread RED-adds to its own element; stride_elems = 1 -> coalesced, 16 -> one 128-byte line per lane
__global__ void __launch_bounds__(256) red_kernel(double *p, size_t n, int stride_elems)
{
    const size_t total = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total)
    {
        size_t j = (stride_elems == 1) ? i : ((i * stride_elems) % n + (i * stride_elems) / n);
        atomicAdd(p + j, 1.0);
    }
}

// plain read-modify-write for comparison (distinct addresses, coalesced 128-bit)
__global__ void __launch_bounds__(256) rmw_kernel(double2 *p, size_t n2)
{
    const size_t total = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += total)
    {
        double2 v = p[i];
        v.x += 1.0; v.y += 1.0;
        p[i] = v;
    }
}

// one CTA repeatedly bulk-reduces a 32 KiB shared buffer into consecutive 32 KiB global segments
__global__ void __launch_bounds__(128) bulkred_kernel(double *p, size_t nseg, int seg_bytes)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double *s = reinterpret_cast<double *>(sm);
    for (int i = threadIdx.x; i < seg_bytes / 8; i += blockDim.x) s[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
        for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x)
        {
            double *dst = p + seg * (size_t)(seg_bytes / 8);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                         ::"l"(dst), "r"(sa), "r"(seg_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template<typename F>
static float time_ms(F &&launch, int reps = 3)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r)
    {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// FMA issue rate of FEW warps: one CTA of `warps` warps per SM, ILP independent chains per thread.  Reports warp
// instructions per clock per SM sub-partition (4 warps -> one per sub-partition); peak is 0.5 (DFMA) / 1.0 (FFMA).
template<typename T, int ILP>
__global__ void fma_few_kernel(T *out, int iters, T a, T b, long long *clk)
{
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    const long long t1 = clock64();
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == T(-12345.678)) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) *clk = t1 - t0;
}

template<typename T, int ILP>
static void few_warps(const char *name, int sms, double *sink)
{
    long long *d_clk, h_clk = 0;
    CK(cudaMalloc(&d_clk, 8));
    for (int warps : {1, 4, 8, 12, 16, 32})
    {
        const int iters = 1 << 13;
        fma_few_kernel<T, ILP><<<sms, 32 * warps>>>((T *)sink, iters, (T)1.0000001, (T)1e-9, d_clk);
        CK(cudaDeviceSynchronize());
        fma_few_kernel<T, ILP><<<sms, 32 * warps>>>((T *)sink, iters, (T)1.0000001, (T)1e-9, d_clk);
        CK(cudaMemcpy(&h_clk, d_clk, 8, cudaMemcpyDeviceToHost));
        const double per_smsp = (double)iters * ILP * ((warps + 3) / 4) / (double)h_clk;
        printf("{\"bench\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"clk\": %lld, \"warp_inst_per_clk_per_smsp\": %.3f}\n",
               name, ILP, warps, h_clk, per_smsp);
    }
    CK(cudaFree(d_clk));
}

template<int NF>
static void mixed(const char *name, int sms, double *sink)
{
    const int iters = 1 << 13, grid = sms * 8;
    float ms = time_ms([&] { mixed_kernel<NF><<<grid, 256>>>(sink, iters, 1.0000001, 1e-9); });
    const double warps = (double)grid * 8, trips = (double)iters * 8;
    const double fl_mma = 2.0 * 256.0 * trips * warps, fl_fma = 2.0 * 32.0 * NF * trips * warps;
    printf("{\"bench\": \"%s\", \"dfma_per_dmma\": %d, \"ms\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f, \"tflops_total\": %.2f}\n",
           name, NF, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9, (fl_mma + fl_fma) / ms * 1e-9);
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "mixed")
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, 0));
        double *sink;
        CK(cudaMalloc(&sink, 1024));
        mixed<0>("dmma_dfma_mixed", prop.multiProcessorCount, sink);
        mixed<2>("dmma_dfma_mixed", prop.multiProcessorCount, sink);
        mixed<4>("dmma_dfma_mixed", prop.multiProcessorCount, sink);
        mixed<8>("dmma_dfma_mixed", prop.multiProcessorCount, sink);
        mixed<16>("dmma_dfma_mixed", prop.multiProcessorCount, sink);
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "dmmalat")
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, 0));
        double *sink;
        long long *clk;
        CK(cudaMalloc(&sink, 1024));
        CK(cudaMalloc(&clk, 64));
        const int sms = prop.multiProcessorCount;
        for (int warps : {4, 8, 12, 16})
        {
            dmma_chain<1>(sms, sink, clk, warps);
            dmma_chain<2>(sms, sink, clk, warps);
            dmma_chain<4>(sms, sink, clk, warps);
            dmma_chain<8>(sms, sink, clk, warps);
        }
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "issue")
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, 0));
        double *sink;
        CK(cudaMalloc(&sink, 1024));
        few_warps<double, 1>("dfma_few", prop.multiProcessorCount, sink);
        few_warps<double, 2>("dfma_few", prop.multiProcessorCount, sink);
        few_warps<double, 4>("dfma_few", prop.multiProcessorCount, sink);
        few_warps<double, 8>("dfma_few", prop.multiProcessorCount, sink);
        few_warps<double, 16>("dfma_few", prop.multiProcessorCount, sink);
        few_warps<float, 1>("ffma_few", prop.multiProcessorCount, sink);
        few_warps<float, 4>("ffma_few", prop.multiProcessorCount, sink);
        few_warps<float, 8>("ffma_few", prop.multiProcessorCount, sink);
        return 0;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"smem_optin\": %zu, \"l2_bytes\": %d}\n", prop.name, sms,
           prop.sharedMemPerBlockOptin, prop.l2CacheSize);
    double *sink;
    CK(cudaMalloc(&sink, 1024));

    {   // FP64 FMA peak: 8 CTAs x 256 threads per SM, 8 independent chains per thread
        const int iters = 1 << 16, ILP = 8, grid = sms * 8;
        float ms = time_ms([&] { fma_kernel<double, ILP><<<grid, 256>>>(sink, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * ILP * (double)iters * grid * 256;
        printf("{\"bench\": \"dfma\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    }
    {
        const int iters = 1 << 17, ILP = 8, grid = sms * 8;
        float ms = time_ms([&] { fma_kernel<float, ILP><<<grid, 256>>>((float *)sink, iters, 1.0000001f, 1e-9f); });
        double fl = 2.0 * ILP * (double)iters * grid * 256;
        printf("{\"bench\": \"ffma\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    }
    {   // FP64 tensor pipe: m8n8k4 = 256 MACs per warp instruction
        const int iters = 1 << 14, grid = sms * 8;
        float ms = time_ms([&] { dmma_kernel<<<grid, 256>>>(sink, iters); });
        double fl = 2.0 * 256.0 * 8 * (double)iters * grid * 8;
        printf("{\"bench\": \"dmma_m8n8k4\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    }
    const size_t bytes = size_t(8) << 30; // 8 GiB >> 126 MB L2
    double *buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    {
        int *isink = (int *)sink;
        float ms = time_ms([&] { read_kernel<<<sms * 16, 256>>>((const int4 *)buf, bytes / 16, isink); });
        printf("{\"bench\": \"read_bw\", \"ms\": %.3f, \"gbs\": %.1f}\n", ms, bytes / ms * 1e-6);
    }
    {
        const size_t n = bytes / 8 / 4; // 2 GiB of doubles
        float ms = time_ms([&] { red_kernel<<<sms * 16, 256>>>(buf, n, 1); });
        printf("{\"bench\": \"red_f64_coalesced\", \"ms\": %.3f, \"gred_per_s\": %.2f, \"rmw_gbs\": %.1f}\n", ms,
               n / ms * 1e-6, 2.0 * n * 8 / ms * 1e-6);
        ms = time_ms([&] { red_kernel<<<sms * 16, 256>>>(buf, n, 16); });
        printf("{\"bench\": \"red_f64_stride128B\", \"ms\": %.3f, \"gred_per_s\": %.2f, \"rmw_gbs\": %.1f}\n", ms,
               n / ms * 1e-6, 2.0 * n * 8 / ms * 1e-6);
        ms = time_ms([&] { rmw_kernel<<<sms * 16, 256>>>((double2 *)buf, n / 2); });
        printf("{\"bench\": \"rmw_f64_plain\", \"ms\": %.3f, \"gelem_per_s\": %.2f, \"rmw_gbs\": %.1f}\n", ms,
               n / ms * 1e-6, 2.0 * n * 8 / ms * 1e-6);
        // L2-resident variant: 32 MiB array hit repeatedly
        const size_t nsmall = (size_t(32) << 20) / 8;
        ms = time_ms([&] { for (int r = 0; r < 8; ++r) red_kernel<<<sms * 16, 256>>>(buf, nsmall, 1); });
        printf("{\"bench\": \"red_f64_coalesced_L2\", \"ms\": %.3f, \"gred_per_s\": %.2f}\n", ms, 8.0 * nsmall / ms * 1e-6);
    }
    {
        const int seg = 32 * 1024;
        const size_t nseg = (bytes / 4) / seg;
        CK(cudaFuncSetAttribute(bulkred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, seg));
        float ms = time_ms([&] { bulkred_kernel<<<sms * 4, 128, seg>>>(buf, nseg, seg); });
        printf("{\"bench\": \"bulkred_f64\", \"ms\": %.3f, \"gred_per_s\": %.2f, \"rmw_gbs\": %.1f}\n", ms,
               (double)nseg * seg / 8 / ms * 1e-6, 2.0 * nseg * seg / ms * 1e-6);
    }
    CK(cudaFree(buf));
    CK(cudaFree(sink));
    return 0;
}
