// fe_sort.cuh -- replica of libstdc++'s std::sort (introsort + final insertion sort) for setMask.
//
// FeatureTracker::setMask (reference feature_tracker.cpp:66-67) sorts (track_cnt, (pt, id)) with the
// comparator a.first > b.first using std::sort, which is unstable: the order of equal track counts is
// an artefact of the algorithm.  Feature order decides which of two close features survives and the
// order of prev_pts, so the replica must perform the same comparisons and moves as libstdc++
// (bits/stl_algo.h, bits/stl_heap.h; unchanged since GCC 4.x: threshold 16, median-of-3 moved to
// first, unguarded Hoare partition, heapsort fallback at depth 2*floor(log2 n)).
// Elements are packed as (track_cnt << 32 | index); only track_cnt takes part in comparisons.
// Verified against the real std::sort by tests/test_host_logic.py (host build of this header).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GF_HD __host__ __device__
#else
#define GF_HD
#endif

namespace gf {

typedef unsigned long long sort_elem;  // (uint32 track_cnt << 32) | uint32 index

GF_HD inline bool sm_comp(sort_elem a, sort_elem b) { return (int)(a >> 32) > (int)(b >> 32); }

GF_HD inline void sm_swap(sort_elem* a, sort_elem* b) { sort_elem t = *a; *a = *b; *b = t; }

GF_HD inline void sm_unguarded_linear_insert(sort_elem* last)
{
    sort_elem val = *last;
    sort_elem* next = last - 1;
    while (sm_comp(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}

GF_HD inline void sm_insertion_sort(sort_elem* first, sort_elem* last)
{
    if (first == last) return;
    for (sort_elem* i = first + 1; i != last; ++i) {
        if (sm_comp(*i, *first)) {
            sort_elem val = *i;
            for (sort_elem* p = i; p != first; --p) *p = *(p - 1);   // move_backward(first, i, i+1)
            *first = val;
        } else
            sm_unguarded_linear_insert(i);
    }
}

GF_HD inline void sm_push_heap(sort_elem* first, long hole, long top, sort_elem value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && sm_comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

GF_HD inline void sm_adjust_heap(sort_elem* first, long hole, long len, sort_elem value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (sm_comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    sm_push_heap(first, hole, top, value);
}

GF_HD inline void sm_heap_sort(sort_elem* first, sort_elem* last)  // __partial_sort(first, last, last)
{
    long len = last - first;
    if (len >= 2) {                                                 // __make_heap
        long parent = (len - 2) / 2;
        while (true) {
            sort_elem value = first[parent];
            sm_adjust_heap(first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {                                      // __sort_heap / __pop_heap
        --last;
        sort_elem value = *last;
        *last = *first;
        sm_adjust_heap(first, 0, last - first, value);
    }
}

GF_HD inline sort_elem* sm_partition_pivot(sort_elem* first, sort_elem* last)
{
    sort_elem* mid = first + (last - first) / 2;
    sort_elem *a = first + 1, *b = mid, *c = last - 1;
    if (sm_comp(*a, *b)) {                                          // __move_median_to_first
        if (sm_comp(*b, *c)) sm_swap(first, b);
        else if (sm_comp(*a, *c)) sm_swap(first, c);
        else sm_swap(first, a);
    } else if (sm_comp(*a, *c)) sm_swap(first, a);
    else if (sm_comp(*b, *c)) sm_swap(first, c);
    else sm_swap(first, b);
    sort_elem* lo = first + 1;                                      // __unguarded_partition
    sort_elem* hi = last;
    while (true) {
        while (sm_comp(*lo, *first)) ++lo;
        --hi;
        while (sm_comp(*first, *hi)) --hi;
        if (!(lo < hi)) return lo;
        sm_swap(lo, hi);
        ++lo;
    }
}

// std::sort(v, v+n, [](a,b){ return a.first > b.first; })
GF_HD inline void setmask_sort(sort_elem* v, int n)
{
    if (n <= 0) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    // explicit stack instead of the recursion on the right part (disjoint ranges: order is immaterial)
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[sp] = 0; stk_last[sp] = n; stk_depth[sp] = 2 * lg; sp++;
    while (sp > 0) {
        sp--;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { sm_heap_sort(v + first, v + last); break; }
            --depth;
            int cut = (int)(sm_partition_pivot(v + first, v + last) - v);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    if (n > 16) {                                                   // __final_insertion_sort
        sm_insertion_sort(v, v + 16);
        for (sort_elem* i = v + 16; i != v + n; ++i) sm_unguarded_linear_insert(i);
    } else
        sm_insertion_sort(v, v + n);
}


#ifdef __CUDACC__
// Parallel replay of setmask_sort by one CTA (call from all threads; v and the work arrays live in shared memory).
// __introsort_loop recurses on disjoint ranges [first, cut) and [cut, last) with the same decremented depth limit,
// so all ranges of one recursion level can be partitioned concurrently, one thread each, performing exactly the
// comparisons and swaps of the sequential code.  __final_insertion_sort then only moves an element within its
// leaf range (every element left of a cut is >= every element right of it and the comparator is strict), so each
// leaf is finished by one thread: the leaf that contains index 0 with the guarded __insertion_sort logic, the others
// with __unguarded_linear_insert (their left neighbour is the sentinel, as in the sequential run).
constexpr int SORT_PAR_MAX = 512;   // larger inputs use the sequential replay (leaf/range tables are sized for this)
struct SortWork { int first[2][128], last[2][128], depth[2][128], n[2]; short leaf_first[SORT_PAR_MAX], leaf_last[SORT_PAR_MAX]; int nleaf; };

__device__ inline void setmask_sort_parallel(sort_elem* v, int n, SortWork& W)
{
    const int tid = threadIdx.x;
    if (n <= 0) return;
    if (n > SORT_PAR_MAX) { if (tid == 0) setmask_sort(v, n); __syncthreads(); return; }
    if (tid == 0) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        W.first[0][0] = 0; W.last[0][0] = n; W.depth[0][0] = 2 * lg; W.n[0] = 1; W.n[1] = 0; W.nleaf = 0;
    }
    __syncthreads();
    int cur = 0;
    while (true) {
        const int nr = W.n[cur];
        if (nr == 0) break;
        if (tid < nr) {
            int first = W.first[cur][tid], last = W.last[cur][tid], depth = W.depth[cur][tid];
            if (last - first > 16) {
                if (depth == 0) {
                    sm_heap_sort(v + first, v + last);                 // __partial_sort(first, last, last): range is final
                    int k = atomicAdd(&W.nleaf, 1); W.leaf_first[k] = (short)first; W.leaf_last[k] = (short)last;
                } else {
                    --depth;
                    int cut = (int)(sm_partition_pivot(v + first, v + last) - v);
                    int k = atomicAdd(&W.n[cur ^ 1], 2);
                    W.first[cur ^ 1][k] = first; W.last[cur ^ 1][k] = cut; W.depth[cur ^ 1][k] = depth;
                    W.first[cur ^ 1][k + 1] = cut; W.last[cur ^ 1][k + 1] = last; W.depth[cur ^ 1][k + 1] = depth;
                }
            } else if (last > first) {
                int k = atomicAdd(&W.nleaf, 1); W.leaf_first[k] = (short)first; W.leaf_last[k] = (short)last;
            }
        }
        __syncthreads();
        if (tid == 0) W.n[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    // final insertion sort, leaf by leaf
    const int nl = W.nleaf;
    if (tid < nl) {
        const int first = W.leaf_first[tid], last = W.leaf_last[tid];
        for (int i = first; i < last; i++) {
            if (i == 0) continue;
            // __final_insertion_sort: elements 1..15 use the guarded form (compare with *begin first), the rest the unguarded one
            if (i < 16 && n > 16 ? true : (n <= 16)) {
                if (sm_comp(v[i], v[0])) {
                    sort_elem val = v[i];
                    for (int p = i; p != 0; --p) v[p] = v[p - 1];
                    v[0] = val;
                    continue;
                }
            }
            sm_unguarded_linear_insert(v + i);
        }
    }
    __syncthreads();
}
#endif

}  // namespace gf
