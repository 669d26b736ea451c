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
// Parallel replay of setmask_sort by one CTA (call from all threads, blockDim.x >= n and a multiple of 32; v and the
// work arrays live in shared memory).  Three observations make the sequential algorithm parallel without changing a
// single comparison outcome or element position:
//  (1) __introsort_loop recurses on the disjoint ranges [first, cut) and [cut, last) with the same decremented depth
//      limit, so all ranges of one recursion level are independent: one warp per range.
//  (2) __unguarded_partition only ever looks at elements it has not moved yet: before swap k, lo scans originals in
//      (L[k-1], R[k-1]) and hi scans originals below R[k-1].  With L = ascending positions where lo stops
//      (!comp(x, pivot)) and R = descending positions where hi stops (!comp(pivot, x)) in the ORIGINAL range, swap k
//      exchanges L[k] and R[k] for every k with L[k] < R[k] (a monotone condition, k* swaps in total), and the
//      returned cut is min(L[k*], R[k*-1]) (lo runs into the first element hi has already moved).  Ranks in L and R are
//      warp prefix sums, the swaps touch disjoint positions.
//  (3) __final_insertion_sort never moves an element across a cut (left of a cut everything is >= pivot >= everything
//      right of it, and the comparator is strict), and insertion sort is stable: its result is the stable sort of every
//      leaf range, i.e. element i of a leaf goes to leaf_first + #{j: key_j > key_i} + #{j < i: key_j == key_i}.
constexpr int SORT_PAR_MAX = 512;   // larger inputs use the sequential replay
constexpr int SORT_QCAP = 64;       // ranges longer than 16 on one level: at most n/17
struct SortWork {
    short first[2][SORT_QCAP], last[2][SORT_QCAP], depth[2][SORT_QCAP];
    int n[2];
    short lpos[SORT_PAR_MAX], rasc[SORT_PAR_MAX];      // per range, at offset first: stop positions of lo / hi (ascending)
    short leaf_first[SORT_PAR_MAX], leaf_last[SORT_PAR_MAX];   // per position: its leaf range
};

// one warp: libstdc++ __unguarded_partition_pivot on [first, last), last - first > 16; returns the cut
__device__ inline int sm_warp_partition(sort_elem* v, int first, int last, SortWork& W, int lane)
{
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    if (lane == 0) {                                                // __move_median_to_first
        sort_elem *f = v + first, *a = f + 1, *b = f + (last - first) / 2, *c = v + last - 1;
        if (sm_comp(*a, *b)) {
            if (sm_comp(*b, *c)) sm_swap(f, b);
            else if (sm_comp(*a, *c)) sm_swap(f, c);
            else sm_swap(f, a);
        } else if (sm_comp(*a, *c)) sm_swap(f, a);
        else if (sm_comp(*b, *c)) sm_swap(f, c);
        else sm_swap(f, b);
    }
    __syncwarp();
    const int pv = (int)(v[first] >> 32);
    short* lpos = W.lpos + first;
    short* rasc = W.rasc + first;
    int nL = 0, nR = 0;
    for (int base = first + 1; base < last; base += 32) {
        const int p = base + lane;
        const int c = (p < last) ? (int)(v[p] >> 32) : 0;
        const bool sl = (p < last) && !(c > pv), sr = (p < last) && !(pv > c);
        const unsigned bl = __ballot_sync(full, sl), br = __ballot_sync(full, sr);
        if (sl) lpos[nL + __popc(bl & lt)] = (short)p;
        if (sr) rasc[nR + __popc(br & lt)] = (short)p;
        nL += __popc(bl);
        nR += __popc(br);
    }
    __syncwarp();
    const int m = min(nL, nR);
    int ks = 0;                                                     // number of swaps
    for (int base = 0; base < m; base += 32) {
        const int k = base + lane;
        const bool ok = (k < m) && (lpos[k] < rasc[nR - 1 - k]);
        ks += __popc(__ballot_sync(full, ok));
    }
    for (int k = lane; k < ks; k += 32) sm_swap(v + lpos[k], v + rasc[nR - 1 - k]);
    int cut = 0x7fffffff;
    if (ks < nL) cut = lpos[ks];
    if (ks >= 1) cut = min(cut, (int)rasc[nR - ks]);
    __syncwarp();
    return cut;
}

__device__ inline void setmask_sort_parallel(sort_elem* v, int n, SortWork& W)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (n <= 0) return;
    if (n > SORT_PAR_MAX) { if (tid == 0) setmask_sort(v, n); __syncthreads(); return; }
    if (tid == 0) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        W.first[0][0] = 0; W.last[0][0] = (short)n; W.depth[0][0] = (short)(2 * lg); W.n[0] = (n > 16) ? 1 : 0; W.n[1] = 0;
    }
    if (n <= 16 && tid < n) { W.leaf_first[tid] = 0; W.leaf_last[tid] = (short)n; }
    __syncthreads();
    int cur = 0;
    while (true) {
        const int nr = W.n[cur];
        if (nr == 0) break;
        for (int r = wid; r < nr; r += nw) {
            const int first = W.first[cur][r], last = W.last[cur][r], depth = W.depth[cur][r];
            if (depth == 0) {                                      // __partial_sort(first, last, last): the range is final
                if (lane == 0) sm_heap_sort(v + first, v + last);
                for (int p = first + lane; p < last; p += 32) { W.leaf_first[p] = (short)first; W.leaf_last[p] = (short)last; }
                __syncwarp();
                continue;
            }
            const int cut = sm_warp_partition(v, first, last, W, lane);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int a = half ? cut : first, b = half ? last : cut;
                if (b - a > 16) {
                    if (lane == 0) {
                        const int k = atomicAdd(&W.n[cur ^ 1], 1);
                        W.first[cur ^ 1][k] = (short)a; W.last[cur ^ 1][k] = (short)b; W.depth[cur ^ 1][k] = (short)(depth - 1);
                    }
                } else
                    for (int p = a + lane; p < b; p += 32) { W.leaf_first[p] = (short)a; W.leaf_last[p] = (short)b; }
            }
        }
        __syncthreads();
        if (tid == 0) W.n[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    // __final_insertion_sort == stable sort of every leaf
    sort_elem e = 0;
    int dst = -1;
    if (tid < n) {
        const int a = W.leaf_first[tid], b = W.leaf_last[tid];
        e = v[tid];
        const int c = (int)(e >> 32);
        dst = a;
        for (int j = a; j < b; j++) {
            const int cj = (int)(v[j] >> 32);
            dst += (cj > c) || (cj == c && j < tid);
        }
    }
    __syncthreads();
    if (dst >= 0) v[dst] = e;
    __syncthreads();
}
#endif

}  // namespace gf
