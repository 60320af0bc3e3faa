// mp_stdsort.h -- std::sort as libstdc++ runs it, for host and device code.
//
// The reference orders the merged single-end seeds of a read with an UNSTABLE std::sort by seed length
// (singleMerge, DV-DPfunctions.cpp:330-336) and then keeps a prefix of that order (>= 60 % of the longest, at most 200 per read,
// DV-DPForSingleReads.cpp:186-199).  Which of several equally long seeds survive therefore depends on what std::sort does with ties.
// The reference is built with GCC/libstdc++ (its Makefile names g++), whose std::sort is: introsort (median-of-three to the first
// position, unguarded Hoare partition, recursion on the right part, heapsort after 2*floor(log2 n) levels) down to ranges of 16, then
// one final insertion sort (bits/stl_algo.h).  This header re-implements that algorithm step for step on a random-access range so that
// a kernel can order a read's seeds exactly as the reference's host code does; tests/test_stdsort.py checks it against the real
// std::sort on inputs full of ties (sizes 0 .. 3000).
#pragma once
#include <stdint.h>

#ifndef MP_HD
#ifdef __CUDACC__
#define MP_HD __host__ __device__
#else
#define MP_HD
#endif
#endif

namespace mp_stdsort {

template <class T> MP_HD inline void swap_(T &a, T &b) { T t = a; a = b; b = t; }

template <class T, class Less> MP_HD inline void unguarded_linear_insert(T *last, Less less)
{
    T val = *last;
    T *next = last - 1;
    while (less(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
template <class T, class Less> MP_HD inline void insertion_sort(T *first, T *last, Less less)
{
    if (first == last) return;
    for (T *i = first + 1; i != last; ++i) {
        if (less(*i, *first)) {
            T val = *i;
            for (T *p = i; p != first; --p) *p = *(p - 1);             // move_backward(first, i, i + 1)
            *first = val;
        } else unguarded_linear_insert(i, less);
    }
}
template <class T, class Less> MP_HD inline void final_insertion_sort(T *first, T *last, Less less)
{
    if (last - first > 16) {
        insertion_sort(first, first + 16, less);
        for (T *i = first + 16; i != last; ++i) unguarded_linear_insert(i, less);
    } else insertion_sort(first, last, less);
}
template <class T, class Less> MP_HD inline void move_median_to_first(T *result, T *a, T *b, T *c, Less less)
{
    if (less(*a, *b)) {
        if (less(*b, *c)) swap_(*result, *b);
        else if (less(*a, *c)) swap_(*result, *c);
        else swap_(*result, *a);
    } else if (less(*a, *c)) swap_(*result, *a);
    else if (less(*b, *c)) swap_(*result, *c);
    else swap_(*result, *b);
}
template <class T, class Less> MP_HD inline T *unguarded_partition(T *first, T *last, T *pivot, Less less)
{
    for (;;) {
        while (less(*first, *pivot)) ++first;
        --last;
        while (less(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swap_(*first, *last);
        ++first;
    }
}
// ---- heapsort fallback: __partial_sort(first, last, last) = __heap_select (make_heap; nothing beyond middle) + __sort_heap ----
template <class T, class Less> MP_HD inline void push_heap_(T *first, long hole, long top, T value, Less less)
{
    long parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
    first[hole] = value;
}
template <class T, class Less> MP_HD inline void adjust_heap(T *first, long hole, long len, T value, Less less)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, less);
}
template <class T, class Less> MP_HD inline void heap_sort(T *first, T *last, Less less)
{
    const long len = last - first;
    if (len >= 2) {                                                    // __make_heap
        long parent = (len - 2) / 2;
        for (;;) {
            T value = first[parent];
            adjust_heap(first, parent, len, value, less);
            if (parent == 0) break;
            --parent;
        }
    }
    while (last - first > 1) {                                         // __sort_heap: __pop_heap(first, last - 1, last - 1)
        --last;
        T value = *last;
        *last = *first;
        adjust_heap(first, 0L, (long)(last - first), value, less);
    }
}
MP_HD inline int lg_(long n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }

// std::sort(first, last, less)
template <class T, class Less> MP_HD inline void sort(T *first, T *last, Less less)
{
    if (first == last) return;
    // __introsort_loop with its recursion on the right part turned into an explicit stack (depth <= 2 lg n + 1 <= 130)
    struct Frame { T *first, *last; int depth; };
    Frame stack[132];
    int sp = 0;
    stack[sp].first = first; stack[sp].last = last; stack[sp].depth = lg_(last - first) * 2; ++sp;
    while (sp) {
        --sp;
        T *f = stack[sp].first, *l = stack[sp].last; int depth = stack[sp].depth;
        while (l - f > 16) {
            if (depth == 0) { heap_sort(f, l, less); break; }
            --depth;
            T *mid = f + (l - f) / 2;
            move_median_to_first(f, f + 1, mid, l - 1, less);
            T *cut = unguarded_partition(f + 1, l, f, less);
            // the reference recurses into [cut, l) first and then continues with [f, cut): both ranges are disjoint, so the order in
            // which they are processed does not change the result; push the right part, go on with the left one
            if (sp < 131) { stack[sp].first = cut; stack[sp].last = l; stack[sp].depth = depth; ++sp; }
            l = cut;
        }
    }
    final_insertion_sort(first, last, less);
}

}  // namespace mp_stdsort
