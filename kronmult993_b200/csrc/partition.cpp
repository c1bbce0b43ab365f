// partition.cpp -- multi-GPU partitioner: shard a batch by output-pointer ownership.
//
// No reference counterpart (the reference is single-GPU: kronmult_gpu/kronmult.cu:185,191 use the
// current device only).  Batch items are independent except for the `+=` into shared outputs
// (kronmult.cu:126-129), so giving every output vector exactly one owning rank removes all
// cross-GPU communication: each rank runs kronmult_batched on its own items and its own outputs.
// Only when one output group is too large to balance (e.g. the reference harness' 5 distinct
// outputs on 8 GPUs, tests/kronmult_bench_gpu.cpp:15) are its items split over all ranks and flagged,
// and the caller must then sum the per-rank partial outputs (one NCCL reduce).
#include "../../include/kronmult_b200.h"

#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

extern "C" int kronmult_partition_by_output(const void *const *out, int nb, int n_ranks, long long split_threshold,
                                            int *owner, unsigned char *needs_reduce)
{
    if (nb < 0 || n_ranks < 1 || (nb > 0 && (!out || !owner))) return 1; // cudaErrorInvalidValue
    if (nb == 0) return 0;
    try
    {
        // groups in order of first appearance; consecutive equal pointers are the common case
        std::unordered_map<const void *, int> index;
        std::vector<int> group_of(nb);
        std::vector<long long> size;
        const void *prev = nullptr;
        int prev_g       = -1;
        for (int k = 0; k < nb; ++k)
        {
            int g;
            if (prev_g >= 0 && out[k] == prev) g = prev_g;
            else
            {
                auto it = index.find(out[k]);
                if (it == index.end())
                {
                    g = (int)size.size();
                    index.emplace(out[k], g);
                    size.push_back(0);
                }
                else g = it->second;
            }
            ++size[g];
            group_of[k] = g;
            prev = out[k]; prev_g = g;
        }
        const int ng = (int)size.size();
        std::vector<int> rank_of(ng, -1);
        std::vector<long long> load(n_ranks, 0);

        // oversized groups are split evenly over all ranks
        std::vector<char> split(ng, 0);
        long long total_whole = 0;
        for (int g = 0; g < ng; ++g)
        {
            if (split_threshold > 0 && size[g] > split_threshold && n_ranks > 1) split[g] = 1;
            else total_whole += size[g];
        }

        long long max_whole = 0;
        for (int g = 0; g < ng; ++g)
            if (!split[g]) max_whole = std::max(max_whole, size[g]);
        const long long ideal = (total_whole + n_ranks - 1) / n_ranks;

        if (max_whole * 8 <= ideal)
        {
            // fine-grained groups: contiguous blocks of groups per rank (keeps each rank's items and
            // outputs contiguous in the original order), balanced to within one group
            long long prefix = 0;
            for (int g = 0; g < ng; ++g)
            {
                if (split[g]) continue;
                const long long mid = prefix + size[g] / 2;
                int r = (int)((mid * n_ranks) / total_whole);
                if (r > n_ranks - 1) r = n_ranks - 1;
                rank_of[g] = r;
                load[r] += size[g];
                prefix += size[g];
            }
        }
        else
        {
            // longest-processing-time greedy: largest group first onto the least loaded rank
            std::vector<int> order;
            for (int g = 0; g < ng; ++g)
                if (!split[g]) order.push_back(g);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return size[a] > size[b]; });
            for (int g : order)
            {
                int best = 0;
                for (int r = 1; r < n_ranks; ++r)
                    if (load[r] < load[best]) best = r;
                rank_of[g] = best;
                load[best] += size[g];
            }
        }

        std::vector<long long> seen(ng, 0);
        for (int k = 0; k < nb; ++k)
        {
            const int g = group_of[k];
            if (split[g])
            {
                // contiguous slices of the group's items, in order of appearance
                const long long i = seen[g]++;
                owner[k]          = (int)((i * n_ranks) / size[g]);
                if (needs_reduce) needs_reduce[k] = 1;
            }
            else
            {
                owner[k] = rank_of[g];
                if (needs_reduce) needs_reduce[k] = 0;
            }
        }
    }
    catch (...)
    {
        return 2; // cudaErrorMemoryAllocation
    }
    return 0;
}
