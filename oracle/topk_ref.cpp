// TEST INFRASTRUCTURE ONLY (see oracle/sgpr_oracle.py header): the CPU `topk` tie behaviour of the reference.
//
// /root/reference/dgcnn.py:19 calls Tensor.topk; on the CPU that is ATen's TopKImpl.h (un-vendored third-party
// dependency, torch==1.6 pinned in requirements.txt:4, torch 2.11 here — same code in both):
//     std::nth_element(queue.begin(), queue.begin() + k - 1, queue.end(), cmp) over (value, index) pairs,
//     cmp(x, y) = (isnan(x.first) && !isnan(y.first)) || x.first > y.first,   then queue[0..k) is the result.
// This file calls the REAL libstdc++ algorithm (no restatement), so that
//   (1) tests/test_tie_rule.py can pin it against torch's own CPU topk on tie-heavy rows, and
//   (2) the product's step-for-step restatement for the device (sg_pr_b200/csrc/topk_nth.cuh) can be checked against it,
//       including the heap-select fallback (forced through libstdc++'s internal __introselect with a small depth budget).
// Built by oracle/build_oracle.py into oracle/_build/libtopk_ref.so (git-ignored).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace {
using elem_t = std::pair<float, int64_t>;
inline bool cmp(const elem_t& x, const elem_t& y) {
    return (std::isnan(x.first) && !std::isnan(y.first)) || (x.first > y.first);
}
}  // namespace

extern "C" {

// rows: [num_rows][n] fp32; idx_out: [num_rows][k] (nth_element order).  depth_limit < 0: plain std::nth_element.
int topk_ref_rows(const float* rows, int num_rows, int n, int k, int depth_limit, int32_t* idx_out) {
    if (n < 1 || k < 1 || k > n) return -1;
    std::vector<elem_t> queue(n);
    for (int r = 0; r < num_rows; ++r) {
        for (int j = 0; j < n; ++j) queue[j] = elem_t(rows[static_cast<size_t>(r) * n + j], j);
        if (depth_limit < 0)
            std::nth_element(queue.begin(), queue.begin() + k - 1, queue.end(), cmp);
        else
            std::__introselect(queue.begin(), queue.begin() + k - 1, queue.end(), static_cast<long>(depth_limit),
                               __gnu_cxx::__ops::__iter_comp_iter(cmp));
        for (int j = 0; j < k; ++j) idx_out[static_cast<size_t>(r) * k + j] = static_cast<int32_t>(queue[j].second);
    }
    return 0;
}

}  // extern "C"
