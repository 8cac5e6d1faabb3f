// oracle/nm_cpu_rule.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// torch.topk(largest=False) as torch-CPU evaluates it for the tiny slices the
// reference feeds it (bfp_ops.py:84, slices of length M, k = M-N).  ATen's
// topk_impl_loop (ATen/native/TopKImpl.h:46-93) copies the slice into
// (value,index) pairs and, because k*64 <= n is false for these sizes, runs
// std::nth_element(begin, begin+k-1, end) with a comparator on the value only
// (NaN sorts last), then reports the first k slots.  Which of several equal
// values land in those slots is whatever libstdc++'s introselect leaves there,
// so the restatement calls the same algorithm with the same comparator.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

extern "C" void oracle_topk_smallest_cpu_rule(const float* absv, int M, int k, int* idx_out) {
    using elem_t = std::pair<float, int64_t>;
    std::vector<elem_t> queue(M);
    for (int j = 0; j < M; ++j) { queue[j].first = absv[j]; queue[j].second = j; }
    auto cmp = [](const elem_t& x, const elem_t& y) -> bool {
        return ((!std::isnan(x.first) && std::isnan(y.first)) || (x.first < y.first));
    };
    if (k * 64 <= M) std::partial_sort(queue.begin(), queue.begin() + k, queue.end(), cmp);
    else std::nth_element(queue.begin(), queue.begin() + k - 1, queue.end(), cmp);
    for (int j = 0; j < k; ++j) idx_out[j] = (int)queue[j].second;
}
