// k-NN selection with the tie behaviour of the reference's CPU path.
//
// `dgcnn.knn` ends in `pairwise_distance.topk(k=k, dim=-1)` (/root/reference/dgcnn.py:19).  On the CPU, ATen's topk
// (aten/src/ATen/native/cpu/TopKImpl.h, the branch taken whenever k*64 > n — always, for graphs of <= 128 nodes) is
//     queue[j] = (value[j], j);  std::nth_element(queue, queue + k - 1, queue + n, cmp);  take queue[0..k)
//     cmp(x, y) = (isnan(x) && !isnan(y)) || x.first > y.first
// so WHICH of several equal values at the k-th position survive is an artefact of libstdc++'s introselect
// (bits/stl_algo.h: median-of-3 to the front, unguarded Hoare partition, insertion sort below 4 elements, heap-select
// once the depth budget 2*floor(log2 n) is spent).  On the one-hot semantic branch (sg_net.py:94) most rows hold such
// a tie, and for graphs with fewer than k zero pads the tied candidates are different nodes — the choice moves the
// score by up to ~0.2.  This header restates that algorithm step for step, so that `knn_ties = cpu`
// selects exactly the reference's CPU index set; the default rule (`cuda`: lowest index first) is what ATen's CUDA
// radix-select does on the reference's native device (profiles/r02_tie_rule_probe.json).
//
// The functions work in place on a value array and a parallel index array; the first k entries afterwards are the
// selection (in nth_element's order; the reference then sorts them, which the max over neighbours does not see).
#pragma once
#include <stdint.h>

#ifndef SGPR_HD
#if defined(__CUDACC__)
#define SGPR_HD __host__ __device__ __forceinline__
#else
#define SGPR_HD inline
#endif
#endif

namespace sgpr {
namespace nth {

SGPR_HD bool before(float a, float b) { return ((a != a) && !(b != b)) || a > b; }   // TopKImpl.h comparator, largest=true

template <typename IdxT>
struct Seq {
    float* v;
    IdxT* ix;
    SGPR_HD bool cmp(int p, int q) const { return before(v[p], v[q]); }
    SGPR_HD void swap(int p, int q) const {
        const float tv = v[p]; v[p] = v[q]; v[q] = tv;
        const IdxT ti = ix[p]; ix[p] = ix[q]; ix[q] = ti;
    }
    SGPR_HD void move(int dst, int src) const { v[dst] = v[src]; ix[dst] = ix[src]; }
};

// std::__move_median_to_first(result, a, b, c)
template <typename IdxT>
SGPR_HD void median_to_first(const Seq<IdxT>& s, int result, int a, int b, int c) {
    if (s.cmp(a, b)) {
        if (s.cmp(b, c)) s.swap(result, b);
        else if (s.cmp(a, c)) s.swap(result, c);
        else s.swap(result, a);
    } else if (s.cmp(a, c)) s.swap(result, a);
    else if (s.cmp(b, c)) s.swap(result, c);
    else s.swap(result, b);
}

// std::__unguarded_partition(first, last, pivot)
template <typename IdxT>
SGPR_HD int unguarded_partition(const Seq<IdxT>& s, int first, int last, int pivot) {
    for (;;) {
        while (s.cmp(first, pivot)) ++first;
        --last;
        while (s.cmp(pivot, last)) --last;
        if (!(first < last)) return first;
        s.swap(first, last);
        ++first;
    }
}

// std::__insertion_sort(first, last)
template <typename IdxT>
SGPR_HD void insertion_sort(const Seq<IdxT>& s, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        const float val = s.v[i];
        const IdxT vi = s.ix[i];
        if (before(val, s.v[first])) {
            for (int j = i; j > first; --j) s.move(j, j - 1);            // std::move_backward(first, i, i + 1)
            s.v[first] = val; s.ix[first] = vi;
        } else {                                                          // std::__unguarded_linear_insert
            int hole = i, next = i - 1;
            while (before(val, s.v[next])) { s.move(hole, next); hole = next; --next; }
            s.v[hole] = val; s.ix[hole] = vi;
        }
    }
}

// std::__push_heap / __adjust_heap on the range [first, first + len)
template <typename IdxT>
SGPR_HD void adjust_heap(const Seq<IdxT>& s, int first, int hole, int len, float val, IdxT vi) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (s.cmp(first + child, first + (child - 1))) --child;
        s.move(first + hole, first + child);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        s.move(first + hole, first + (child - 1));
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;                                          // std::__push_heap
    while (hole > top && before(s.v[first + parent], val)) {
        s.move(first + hole, first + parent);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    s.v[first + hole] = val; s.ix[first + hole] = vi;
}

// std::__heap_select(first, middle, last)
template <typename IdxT>
SGPR_HD void heap_select(const Seq<IdxT>& s, int first, int middle, int last) {
    const int len = middle - first;
    if (len >= 2) {                                                       // std::__make_heap
        for (int parent = (len - 2) / 2;; --parent) {
            adjust_heap(s, first, parent, len, s.v[first + parent], s.ix[first + parent]);
            if (parent == 0) break;
        }
    }
    for (int i = middle; i < last; ++i) {
        if (s.cmp(i, first)) {                                            // std::__pop_heap(first, middle, i)
            const float val = s.v[i];
            const IdxT vi = s.ix[i];
            s.move(i, first);
            adjust_heap(s, first, 0, len, val, vi);
        }
    }
}

// std::__introselect(first, nth, last, depth_limit)
template <typename IdxT>
SGPR_HD void introselect(const Seq<IdxT>& s, int first, int nth, int last, int depth_limit) {
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(s, first, nth + 1, last);
            s.swap(first, nth);
            return;
        }
        --depth_limit;
        const int mid = first + (last - first) / 2;                      // std::__unguarded_partition_pivot
        median_to_first(s, first, first + 1, mid, last - 1);
        const int cut = unguarded_partition(s, first + 1, last, first);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(s, first, last);
}

SGPR_HD int floor_log2(int n) {
    int l = 0;
    while (n > 1) { n >>= 1; ++l; }
    return l;
}

// ATen CPU topk(k, largest=True) selection over v[0..n), ix[j] = j on entry: afterwards ix[0..k) is the selected set.
template <typename IdxT>
SGPR_HD void topk_cpu_rule(float* v, IdxT* ix, int n, int k) {
    if (n == 0 || k < 1 || k > n) return;                                 // k == n: nth == last - 1, still runs
    const Seq<IdxT> s{v, ix};
    introselect(s, 0, k - 1, n, 2 * floor_log2(n));                       // std::nth_element(first, first + k - 1, last)
}

}  // namespace nth
}  // namespace sgpr
