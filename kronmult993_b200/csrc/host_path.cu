// host_path.cu -- kronmult_batched on HOST buffers (end-to-end entry points).
//
// Signature of the reference's CPU flavour (kronmult_omp/kronmult.hpp:77-80): all pointer arrays and
// all pointees are host memory.  Instead of the OpenMP loop (kronmult.hpp:86-102) the batch is
// streamed through the GPU in chunks of items:
//   host vectors/factors --(direct cudaMemcpyAsync when they are contiguous in host memory, else a
//   multi-threaded gather into pinned staging)--> device chunk buffers --> the device path of this
//   library (kronmult_batched_*_async) --> one device copy of every distinct output RANGE (overlapping or
//   touching output vectors are merged into spans, so partial overlaps accumulate like the device path), which
//   is seeded with the caller's current output values and copied back at the end.
// Two chunk buffer sets on two streams overlap the copies of chunk c+1 with the kernel of chunk c.
// Input and workspace arrays are never written (the contract would allow it, kronmult.hpp:72).
#include "../../include/kronmult_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace kron
{

template<typename T>
__global__ void build_chunk_pointers(T *d_in, T *d_A, T *d_out, const unsigned long long *__restrict__ out_off,
                                     int count, int d, int N, int nn, T **pin, const T **pA, T **pout)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count)
    {
        pin[i]  = d_in + (size_t)i * N;
        pout[i] = d_out + out_off[i]; // element offset of the item's output window in the device copy
    }
    if (i < count * d) pA[i] = d_A + (size_t)i * nn;
}

struct Buffers
{
    void *dev = nullptr, *pinned = nullptr;
    size_t dev_bytes = 0, pinned_bytes = 0;
    cudaError_t ensure(size_t db, size_t pb)
    {
        if (db > dev_bytes)
        {
            if (dev) cudaFree(dev);
            dev = nullptr; dev_bytes = 0;
            cudaError_t e = cudaMalloc(&dev, db);
            if (e != cudaSuccess) return e;
            dev_bytes = db;
        }
        if (pb > pinned_bytes)
        {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr; pinned_bytes = 0;
            cudaError_t e = cudaHostAlloc(&pinned, pb, cudaHostAllocDefault);
            if (e != cudaSuccess) return e;
            pinned_bytes = pb;
        }
        return cudaSuccess;
    }
};

// How the batch's output vectors map onto device memory.  Output vectors that overlap or touch in host memory
// are merged into one SPAN (a maximal contiguous host range); every item's output is a window into its span's
// device copy, so partially overlapping outputs accumulate correctly (the device path is atomic-class for any
// aliasing, like kronmult.cu:126-129) and a contiguous output slab is one span = one copy each way.
// Cached across calls: ASGarD passes the same pointer arrays every time step.
struct OutputMap
{
    const void *out_array = nullptr; // identity of the caller's array ...
    int nb = -1, N = 0, elem = 0;
    unsigned long long hash = 0;     // ... and of its contents
    std::vector<size_t> item_off;    // per item: element offset of its output in the device copy
    struct Span { char *host; size_t elems, dev_off; };
    std::vector<Span> spans;
    size_t total_elems = 0;
};

struct HostContext
{
    std::mutex mtx;
    Buffers set[2], outbuf;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t done[2]    = {nullptr, nullptr};
    int device             = -1;
    OutputMap omap;
};
constexpr int MAX_DEVICES = 64;
static HostContext g_ctx[MAX_DEVICES];

// restores the caller's current device on every exit path (the CPU flavour this mirrors has no such side effect)
struct DeviceGuard
{
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static unsigned long long hash_pointers(const void *const *p, size_t count)
{
    unsigned long long h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < count; ++i)
    {
        h ^= reinterpret_cast<unsigned long long>(p[i]) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    return h;
}

template<typename T>
static void build_output_map(OutputMap &om, T *const *out, int nb, int N)
{
    const unsigned long long h = hash_pointers(reinterpret_cast<const void *const *>(out), (size_t)nb);
    if (om.out_array == out && om.nb == nb && om.N == N && om.elem == (int)sizeof(T) && om.hash == h) return; // cached
    om.out_array = out; om.nb = nb; om.N = N; om.elem = (int)sizeof(T); om.hash = h;
    // distinct output pointers (consecutive repeats are the common case)
    std::vector<char *> uniq;
    {
        std::unordered_map<const void *, int> seen;
        const T *prev = nullptr;
        for (int k = 0; k < nb; ++k)
        {
            if (out[k] == prev) continue;
            prev = out[k];
            if (seen.emplace(out[k], 1).second) uniq.push_back(reinterpret_cast<char *>(out[k]));
        }
    }
    std::sort(uniq.begin(), uniq.end());
    // merge overlapping / touching ranges [p, p + N*s) into spans; an overlap that is not a whole number of
    // elements apart cannot be expressed as element offsets -> kept as separate spans only if disjoint
    const size_t bytes = (size_t)N * sizeof(T);
    om.spans.clear();
    std::unordered_map<const void *, size_t> dev_off; // element offset of every distinct output
    size_t total = 0;
    for (size_t u = 0; u < uniq.size(); ++u)
    {
        char *p = uniq[u];
        if (!om.spans.empty())
        {
            OutputMap::Span &sp = om.spans.back();
            char *end = sp.host + sp.elems * sizeof(T);
            if (p <= end && (size_t)(p - sp.host) % sizeof(T) == 0)
            {
                const size_t off = (size_t)(p - sp.host) / sizeof(T);
                if (off + N > sp.elems) { total += off + N - sp.elems; sp.elems = off + N; }
                dev_off[p] = sp.dev_off + off;
                continue;
            }
        }
        om.spans.push_back({p, (size_t)N, total});
        dev_off[p] = total;
        total += N;
        (void)bytes;
    }
    om.total_elems = total;
    om.item_off.resize(nb);
    const T *prev = nullptr; size_t prev_off = 0;
    for (int k = 0; k < nb; ++k)
    {
        if (out[k] != prev) { prev = out[k]; prev_off = dev_off[out[k]]; }
        om.item_off[k] = prev_off;
    }
}

template<typename F>
static void parallel_for(size_t count, size_t min_per_thread, F &&fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt   = std::min<size_t>(hw ? hw : 4, 32);
    nt          = std::min(nt, std::max<size_t>(1, count / std::max<size_t>(1, min_per_thread)));
    if (nt <= 1) { fn(0, count); return; }
    std::vector<std::thread> pool;
    const size_t per = (count + nt - 1) / nt;
    for (size_t i = 0; i < nt; ++i)
    {
        const size_t a = i * per, b = std::min(count, a + per);
        if (a >= b) break;
        pool.emplace_back([&fn, a, b] { fn(a, b); });
    }
    for (auto &th : pool) th.join();
}

#define KRON_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

template<typename T>
static int async_call(int d, int n, const T *const *A, int lda, T **in, T **out, int nb, cudaStream_t st);
template<>
int async_call<double>(int d, int n, const double *const *A, int lda, double **in, double **out, int nb, cudaStream_t st)
{
    return kronmult_batched_f64_async(d, n, A, lda, in, out, nullptr, nb, st);
}
template<>
int async_call<float>(int d, int n, const float *const *A, int lda, float **in, float **out, int nb, cudaStream_t st)
{
    return kronmult_batched_f32_async(d, n, A, lda, in, out, nullptr, nb, st);
}

template<typename T>
static int host_call(int d, int n, const T *const *A, int lda, T **in, T **out, int nb, int device)
{
    if (nb <= 0) return 0;
    if (d < 0 || n < 1 || lda < n || (!A && d > 0) || !in || !out) return (int)cudaErrorInvalidValue; // as dispatch()
    DeviceGuard guard;
    if (device >= 0)
    {
        int cur = -1;
        KRON_TRY(cudaGetDevice(&cur));
        if (cur != device)
        {
            KRON_TRY(cudaSetDevice(device));
            guard.prev = cur;
        }
    }
    int dev = 0;
    KRON_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return (int)cudaErrorInvalidDevice;
    long long N64 = 1;
    for (int i = 0; i < d; ++i)
    {
        N64 *= n;
        if (N64 >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    }
    const int N = (int)N64, nn = n * n;
    const size_t s = sizeof(T);

    HostContext &cx = g_ctx[dev];
    std::lock_guard<std::mutex> lock(cx.mtx);
    for (int i = 0; i < 2; ++i)
    {
        if (!cx.stream[i]) KRON_TRY(cudaStreamCreateWithFlags(&cx.stream[i], cudaStreamNonBlocking));
        if (!cx.done[i]) KRON_TRY(cudaEventCreateWithFlags(&cx.done[i], cudaEventDisableTiming));
    }

    // ---- output vectors -> spans of the device copy (cached across calls with the same pointer array)
    build_output_map<T>(cx.omap, out, nb, N);
    const OutputMap &om = cx.omap;
    const size_t U_el   = om.total_elems;
    const bool few_spans = om.spans.size() <= 64;
    KRON_TRY(cx.outbuf.ensure(U_el * s, few_spans ? 0 : U_el * s));
    T *d_out = static_cast<T *>(cx.outbuf.dev);
    if (few_spans)
    {
        for (const auto &sp : om.spans)
            KRON_TRY(cudaMemcpyAsync(d_out + sp.dev_off, sp.host, sp.elems * s, cudaMemcpyHostToDevice, cx.stream[0]));
    }
    else
    {
        T *stage = static_cast<T *>(cx.outbuf.pinned);
        parallel_for(om.spans.size(), 64, [&](size_t a, size_t b) {
            for (size_t u = a; u < b; ++u) std::memcpy(stage + om.spans[u].dev_off, om.spans[u].host, om.spans[u].elems * s);
        });
        KRON_TRY(cudaMemcpyAsync(d_out, stage, U_el * s, cudaMemcpyHostToDevice, cx.stream[0]));
    }
    KRON_TRY(cudaStreamSynchronize(cx.stream[0]));

    // ---- chunks of items, double-buffered
    size_t C = (size_t(128) << 20) / std::max<size_t>(1, N * s);
    C        = std::max<size_t>(1, std::min<size_t>(C, (size_t)nb));
    auto up  = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_in = 0, o_A = o_in + up(C * N * s), o_pin = o_A + up(C * d * nn * s), o_pA = o_pin + up(C * 8),
                 o_pout = o_pA + up(C * d * 8), o_slot = o_pout + up(C * 8), dev_total = o_slot + up(C * 8);
    const size_t h_in = 0, h_A = h_in + up(C * N * s), h_slot = h_A + up(C * d * nn * s), pin_total = h_slot + up(C * 8);
    for (int i = 0; i < 2; ++i) KRON_TRY(cx.set[i].ensure(dev_total, pin_total));

    int which = 0;
    for (size_t k0 = 0; k0 < (size_t)nb; k0 += C, which ^= 1)
    {
        const size_t cnt = std::min(C, (size_t)nb - k0);
        Buffers &bs      = cx.set[which];
        cudaStream_t st  = cx.stream[which];
        char *dv = static_cast<char *>(bs.dev), *hp = static_cast<char *>(bs.pinned);
        KRON_TRY(cudaEventSynchronize(cx.done[which])); // staging of this set is free again

        // inputs
        bool in_contig = true;
        for (size_t i = 1; i < cnt && in_contig; ++i) in_contig = (in[k0 + i] == in[k0] + i * N);
        if (in_contig) KRON_TRY(cudaMemcpyAsync(dv + o_in, in[k0], cnt * N * s, cudaMemcpyHostToDevice, st));
        else
        {
            T *stage = reinterpret_cast<T *>(hp + h_in);
            parallel_for(cnt, 64, [&](size_t a, size_t b) {
                for (size_t i = a; i < b; ++i) std::memcpy(stage + i * N, in[k0 + i], N * s);
            });
            KRON_TRY(cudaMemcpyAsync(dv + o_in, stage, cnt * N * s, cudaMemcpyHostToDevice, st));
        }
        // factors -> compact n x n blocks
        const size_t nm = cnt * d;
        bool A_contig   = (lda == n);
        for (size_t m = 1; m < nm && A_contig; ++m) A_contig = (A[k0 * d + m] == A[k0 * d] + m * nn);
        if (nm > 0)
        {
            if (A_contig) KRON_TRY(cudaMemcpyAsync(dv + o_A, A[k0 * d], nm * nn * s, cudaMemcpyHostToDevice, st));
            else
            {
                T *stage = reinterpret_cast<T *>(hp + h_A);
                parallel_for(nm, 256, [&](size_t a, size_t b) {
                    for (size_t m = a; m < b; ++m)
                        for (int c = 0; c < n; ++c)
                            std::memcpy(stage + m * nn + (size_t)c * n, A[k0 * d + m] + (size_t)c * lda, n * s);
                });
                KRON_TRY(cudaMemcpyAsync(dv + o_A, stage, nm * nn * s, cudaMemcpyHostToDevice, st));
            }
        }
        static_assert(sizeof(size_t) == 8, "output offsets travel as 64-bit values");
        std::memcpy(hp + h_slot, om.item_off.data() + k0, cnt * 8);
        KRON_TRY(cudaMemcpyAsync(dv + o_slot, hp + h_slot, cnt * 8, cudaMemcpyHostToDevice, st));

        const int total = (int)std::max(cnt, nm);
        build_chunk_pointers<T><<<(total + 255) / 256, 256, 0, st>>>(
            reinterpret_cast<T *>(dv + o_in), reinterpret_cast<T *>(dv + o_A), d_out,
            reinterpret_cast<const unsigned long long *>(dv + o_slot), (int)cnt, d, N, nn, reinterpret_cast<T **>(dv + o_pin),
            reinterpret_cast<const T **>(dv + o_pA), reinterpret_cast<T **>(dv + o_pout));
        KRON_TRY(cudaGetLastError());
        int rc = async_call<T>(d, n, reinterpret_cast<const T *const *>(dv + o_pA), n,
                               reinterpret_cast<T **>(dv + o_pin), reinterpret_cast<T **>(dv + o_pout), (int)cnt, st);
        if (rc != 0) return rc;
        KRON_TRY(cudaEventRecord(cx.done[which], st));
    }
    KRON_TRY(cudaStreamSynchronize(cx.stream[0]));
    KRON_TRY(cudaStreamSynchronize(cx.stream[1]));

    // ---- outputs back
    if (few_spans)
    {
        for (const auto &sp : om.spans)
            KRON_TRY(cudaMemcpyAsync(sp.host, d_out + sp.dev_off, sp.elems * s, cudaMemcpyDeviceToHost, cx.stream[0]));
        KRON_TRY(cudaStreamSynchronize(cx.stream[0]));
    }
    else
    {
        T *stage = static_cast<T *>(cx.outbuf.pinned);
        KRON_TRY(cudaMemcpy(stage, d_out, U_el * s, cudaMemcpyDeviceToHost));
        parallel_for(om.spans.size(), 64, [&](size_t a, size_t b) {
            for (size_t u = a; u < b; ++u) std::memcpy(om.spans[u].host, stage + om.spans[u].dev_off, om.spans[u].elems * s);
        });
    }
    return 0;
}

} // namespace kron

extern "C"
{
int kronmult_batched_host_f64(int d, int n, const double *const *A, int lda, double **in, double **out, double **ws,
                              int nb, int device)
{
    (void)ws;
    return kron::host_call<double>(d, n, A, lda, in, out, nb, device);
}
int kronmult_batched_host_f32(int d, int n, const float *const *A, int lda, float **in, float **out, float **ws,
                              int nb, int device)
{
    (void)ws;
    return kron::host_call<float>(d, n, A, lda, in, out, nb, device);
}
}
