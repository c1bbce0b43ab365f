// planner.cuh -- the batch / aliasing planner (BASELINE.json north_star: "sorting or grouping batch items by
// output pointer ... so they reduce in-block").  No reference counterpart: the reference adds every element
// of every item with atomicAdd (kronmult_gpu/kronmult.cu:126-129) whatever the order of the batch.
//
// Every kernel family of this library sums RUNS of consecutive items with the same output pointer on chip and
// flushes one atomic-class add per element per run.  ASGarD hands over batches that are already grouped; for a
// batch whose equal output pointers are scattered (tests: alias = "shuffled") every item is its own run and the
// output vectors are read-modify-written in HBM once per item instead of once per group -- 3x the traffic on
// BASELINE config 5.  A plan makes the runs long again:
//   1. stable LSD radix sort of (output pointer, item index) on the device (CUB; pointer bits [2, 48));
//   2. count runs before / after; keep the permutation only if it at least halves the number of runs;
//   3. gather the three pointer arrays in sorted order into plan-owned device arrays.
// Executing a plan is the ordinary dispatch on the gathered arrays: no kernel knows about plans, and because
// every flush stays an atomic-class add the result is correct for any aliasing either way.
//
// The blocking drop-in entry points consult a small cache of plans keyed on the pointer arrays' addresses,
// sizes and content hashes (ASGarD calls kronmult with the same pointer arrays every time step), so the sort is
// paid once; the stream-ordered entry points never plan implicitly (planning needs a host round trip) -- use
// kronmult_plan_create / kronmult_plan_execute there.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <list>

namespace kron
{

struct PlanStats
{
    unsigned long long runs;     // number of runs of equal consecutive output pointers
    unsigned long long hash_out; // order-sensitive content hashes of the three pointer arrays
    unsigned long long hash_in;
    unsigned long long hash_A;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// one pass over the pointer arrays: run count of `out` and position-dependent hashes of all three
__global__ void __launch_bounds__(256) plan_scan_kernel(const unsigned long long *__restrict__ out,
                                                        const unsigned long long *__restrict__ in,
                                                        const unsigned long long *__restrict__ A, int nb, int d,
                                                        PlanStats *stats)
{
    unsigned long long runs = 0, ho = 0, hi = 0, ha = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += stride)
    {
        const unsigned long long o = out[k];
        runs += (k == 0 || out[k - 1] != o) ? 1 : 0;
        ho += mix64(o + (unsigned long long)k * 0x632BE59BD9B4E019ull);
        if (in) hi += mix64(in[k] + (unsigned long long)k * 0x632BE59BD9B4E019ull);
    }
    if (A)
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)nb * d; e += stride)
            ha += mix64(A[e] + (unsigned long long)e * 0x632BE59BD9B4E019ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        runs += __shfl_xor_sync(0xffffffffu, runs, o);
        ho += __shfl_xor_sync(0xffffffffu, ho, o);
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
        ha += __shfl_xor_sync(0xffffffffu, ha, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicAdd(&stats->runs, runs);
        atomicAdd(&stats->hash_out, ho);
        atomicAdd(&stats->hash_in, hi);
        atomicAdd(&stats->hash_A, ha);
    }
}

__global__ void __launch_bounds__(256) plan_iota_kernel(unsigned *v, int nb)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nb) v[k] = (unsigned)k;
}

// sorted pointer arrays: item i of the plan is item perm[i] of the caller
__global__ void __launch_bounds__(256) plan_gather_kernel(const unsigned *__restrict__ perm,
                                                          const unsigned long long *__restrict__ in,
                                                          const unsigned long long *__restrict__ A, int nb, int d,
                                                          unsigned long long *in_s, unsigned long long *A_s)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nb) in_s[e] = in[perm[e]];
    if (e < (long long)nb * d)
    {
        const long long i = e / d;
        const int j       = (int)(e - i * d);
        A_s[e]            = A[(long long)perm[i] * d + j];
    }
}

struct Plan
{
    int elem = 0, d = 0, n = 0, lda = 0, nb = 0, device = 0;
    const void *A0 = nullptr, *in0 = nullptr, *out0 = nullptr; // the caller's pointer arrays
    PlanStats before{};                                        // of the caller's arrays
    unsigned long long runs_after = 0;
    bool permuted = false;
    void *slab = nullptr; // one allocation: out_s | in_s | A_s
    unsigned long long *out_s = nullptr, *in_s = nullptr, *A_s = nullptr;
    ~Plan() { if (slab) cudaFree(slab); }
};

static cudaError_t plan_scan(const void *A, const void *in, const void *out, int nb, int d, int sms, cudaStream_t st,
                             PlanStats &host)
{
    static thread_local PlanStats *d_stats_dev[64] = {};
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    PlanStats *&d_stats = d_stats_dev[dev];
    if (!d_stats)
    {
        e = cudaMalloc(&d_stats, sizeof(PlanStats));
        if (e != cudaSuccess) { d_stats = nullptr; return e; }
    }
    e = cudaMemsetAsync(d_stats, 0, sizeof(PlanStats), st);
    if (e != cudaSuccess) return e;
    long long blocks = ((long long)nb * (d > 0 ? d : 1) + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    plan_scan_kernel<<<(int)blocks, 256, 0, st>>>(static_cast<const unsigned long long *>(out),
                                                  static_cast<const unsigned long long *>(in),
                                                  static_cast<const unsigned long long *>(A), nb, d, d_stats);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(&host, d_stats, sizeof(PlanStats), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

// Builds the plan on `st` and waits for it (the run counts travel to the host).  `known` may carry the scan
// of the caller's arrays when the caller has already done it.
static cudaError_t plan_build(Plan &p, int sms, cudaStream_t st, const PlanStats *known)
{
    cudaError_t e;
    if (known) p.before = *known;
    else
    {
        e = plan_scan(p.A0, p.in0, p.out0, p.nb, p.d, sms, st, p.before);
        if (e != cudaSuccess) return e;
    }
    p.runs_after = p.before.runs;
    p.permuted   = false;
    if (p.nb < 2 || p.before.runs < 2) return cudaSuccess;

    const size_t nb = (size_t)p.nb;
    unsigned *vals = nullptr, *vals_tmp = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    const size_t o_in = nb * 8, o_A = o_in + nb * 8, total = o_A + nb * (size_t)(p.d > 0 ? p.d : 1) * 8;
    auto fail = [&](cudaError_t err) {
        cudaFree(vals); cudaFree(vals_tmp); cudaFree(cub_tmp);
        if (p.slab) { cudaFree(p.slab); p.slab = nullptr; }
        return err;
    };
    if ((e = cudaMalloc(&p.slab, total)) != cudaSuccess) return fail(e);
    p.out_s = static_cast<unsigned long long *>(p.slab);
    p.in_s  = p.out_s + nb;
    p.A_s   = p.in_s + nb;
    if ((e = cudaMalloc(&vals, nb * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&vals_tmp, nb * 4)) != cudaSuccess) return fail(e);
    plan_iota_kernel<<<(p.nb + 255) / 256, 256, 0, st>>>(vals_tmp, p.nb);
    const unsigned long long *keys_in = static_cast<const unsigned long long *>(p.out0);
    // device addresses fit 48 bits and vectors are at least 4-byte aligned
    if ((e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys_in, p.out_s, vals_tmp, vals, p.nb, 2, 48, st)) != cudaSuccess)
        return fail(e);
    if ((e = cudaMalloc(&cub_tmp, cub_bytes ? cub_bytes : 16)) != cudaSuccess) return fail(e);
    if ((e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, p.out_s, vals_tmp, vals, p.nb, 2, 48, st)) != cudaSuccess)
        return fail(e);
    PlanStats after{};
    if ((e = plan_scan(nullptr, nullptr, p.out_s, p.nb, 0, sms, st, after)) != cudaSuccess) return fail(e);
    p.runs_after = after.runs;
    if (after.runs * 2 <= p.before.runs)
    {
        const long long elems = (long long)nb * (p.d > 1 ? p.d : 1);
        plan_gather_kernel<<<(int)((elems + 255) / 256), 256, 0, st>>>(
            vals, static_cast<const unsigned long long *>(p.in0), static_cast<const unsigned long long *>(p.A0), p.nb, p.d,
            p.in_s, p.A_s);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
        p.permuted = true;
    }
    cudaFree(vals); cudaFree(vals_tmp); cudaFree(cub_tmp);
    if (!p.permuted) { cudaFree(p.slab); p.slab = nullptr; p.out_s = p.in_s = p.A_s = nullptr; }
    return cudaSuccess;
}

template<typename T>
static cudaError_t plan_execute(const Plan &p, cudaStream_t st)
{
    if (p.permuted)
        return dispatch<T>(p.d, p.n, reinterpret_cast<const T *const *>(p.A_s), p.lda, reinterpret_cast<T *const *>(p.in_s),
                           reinterpret_cast<T *const *>(p.out_s), p.nb, st);
    return dispatch<T>(p.d, p.n, static_cast<const T *const *>(p.A0), p.lda, static_cast<T *const *>(p.in0),
                       static_cast<T *const *>(p.out0), p.nb, st);
}

// ---- implicit planning for the blocking entry points -------------------------------------------------------------
static std::atomic<int> g_autoplan{1};
static std::atomic<long long> g_plan_hits{0}, g_plan_builds{0};

struct PlanCache
{
    std::mutex mtx;
    std::list<Plan *> lru; // front = most recent
    static constexpr size_t CAP = 4;
    ~PlanCache() { /* process exit: the context may already be gone, leak on purpose */ }
};
static PlanCache g_cache;

// the blocking call: plan when the batch is large, its vectors are long enough for flushes to matter, and the
// scan says that most items start a new run
template<typename T>
static cudaError_t autoplan_call(int d, int n, const T *const *A, int lda, T *const *in, T *const *out, int nb,
                                 cudaStream_t st, bool &handled)
{
    handled = false;
    if (!g_autoplan.load(std::memory_order_relaxed) || g_force.load(std::memory_order_relaxed) != PATH_AUTO) return cudaSuccess;
    if (nb < 4096 || d < 1 || n < 2 || lda < n || !A || !in || !out) return cudaSuccess;
    long long N = 1;
    for (int i = 0; i < d; ++i) { N *= n; if (N >= (1LL << 31)) return cudaSuccess; }
    if (N < 256) return cudaSuccess;
    DeviceInfo di;
    if (device_info(di) != cudaSuccess) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    PlanStats now{};
    if (plan_scan(A, in, out, nb, d, di.sms, st, now) != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }
    if (now.runs * 2 <= (unsigned long long)nb) return cudaSuccess; // already grouped: average run >= 2

    std::lock_guard<std::mutex> lk(g_cache.mtx);
    Plan *hit = nullptr;
    for (auto it = g_cache.lru.begin(); it != g_cache.lru.end(); ++it)
    {
        Plan *p = *it;
        if (p->elem == (int)sizeof(T) && p->d == d && p->n == n && p->lda == lda && p->nb == nb && p->device == dev &&
            p->A0 == A && p->in0 == in && p->out0 == out && p->before.hash_out == now.hash_out &&
            p->before.hash_in == now.hash_in && p->before.hash_A == now.hash_A && p->before.runs == now.runs)
        {
            hit = p;
            g_cache.lru.erase(it);
            break;
        }
    }
    if (hit) g_plan_hits.fetch_add(1, std::memory_order_relaxed);
    else
    {
        hit = new Plan;
        hit->elem = (int)sizeof(T); hit->d = d; hit->n = n; hit->lda = lda; hit->nb = nb; hit->device = dev;
        hit->A0 = A; hit->in0 = in; hit->out0 = out;
        if (plan_build(*hit, di.sms, st, &now) != cudaSuccess)
        {
            cudaGetLastError(); // e.g. out of memory for the sorted copies: run unplanned
            delete hit;
            return cudaSuccess;
        }
        g_plan_builds.fetch_add(1, std::memory_order_relaxed);
    }
    g_cache.lru.push_front(hit);
    while (g_cache.lru.size() > PlanCache::CAP)
    {
        delete g_cache.lru.back();
        g_cache.lru.pop_back();
    }
    handled = true;
    return plan_execute<T>(*hit, st);
}

} // namespace kron
