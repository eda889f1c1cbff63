// miso_b200/csrc/bm_sort.hpp -- the draw-order sort, one statement for host and device.
//
// The reference orders a gene's reads with its own qsort (a Bentley-McIlroy quicksort,
// /root/reference/pysplicing/src/qsort.c, called from splicing_order_matches, src/miso.c:988-993,
// with the column comparison of include/matrix.pmt:546-561).  The sort is UNSTABLE and the order of
// equal columns is observable -- the uniforms are dealt to the reads in this order -- so it has to be
// reproduced step by step: same pivot choice (median of three / ninther), same split-end
// partition, same insertion sort below 7 elements.  This is that algorithm as an index sort with an
// explicit stack instead of recursion (the sub-ranges are independent, so the order in which they
// are finished does not matter), compiled by the host plan stage (plan.cpp) and by the device
// kernel (match.cu): the two cannot disagree.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define MISOB200_HD_SORT __host__ __device__ __forceinline__
#else
#define MISOB200_HD_SORT inline
#endif

namespace misob200 {

constexpr int kSortStack = 64;      // pending sub-ranges; the smaller side is finished first: depth <= log2(n)

template <class Cmp>
struct BMSort {
  Cmp cmp;
  static MISOB200_HD_SORT void swp(int32_t *v, long a, long b) { int32_t t = v[a]; v[a] = v[b]; v[b] = t; }
  MISOB200_HD_SORT void insertion(int32_t *v, long n) const {
    for (long m = 1; m < n; m++)
      for (long l = m; l > 0 && cmp(v[l - 1], v[l]) > 0; l--) swp(v, l, l - 1);
  }
  MISOB200_HD_SORT long med3(const int32_t *v, long a, long b, long c) const {
    if (cmp(v[a], v[b]) < 0) {
      if (cmp(v[b], v[c]) < 0) return b;
      return cmp(v[a], v[c]) < 0 ? c : a;
    }
    if (cmp(v[b], v[c]) > 0) return b;
    return cmp(v[a], v[c]) < 0 ? a : c;
  }
  // returns false if the explicit stack overflowed (cannot happen for n < 2^63; kept as a guard)
  MISOB200_HD_SORT bool sort(int32_t *v0, long n0) const {
    long st_off[kSortStack], st_n[kSortStack];
    int sp = 0;
    st_off[sp] = 0; st_n[sp] = n0; sp++;
    while (sp > 0) {
      sp--;
      int32_t *v = v0 + st_off[sp];
      long n = st_n[sp];
      const long base = st_off[sp];
      long off = 0;                         // v = v0 + base + off
      while (true) {
        if (n < 7) { insertion(v, n); break; }
        long mid = n / 2;
        if (n > 7) {
          long lo = 0, hi = n - 1;
          if (n > 40) {
            const long d = n / 8;
            lo = med3(v, lo, lo + d, lo + 2 * d);
            mid = med3(v, mid - d, mid, mid + d);
            hi = med3(v, hi - 2 * d, hi - d, hi);
          }
          mid = med3(v, lo, mid, hi);
        }
        swp(v, 0, mid);                       // pivot parked at v[0]
        long a = 1, b = 1, c = n - 1, d = n - 1;
        bool moved = false;
        while (true) {
          int r;
          while (b <= c && (r = cmp(v[b], v[0])) <= 0) {
            if (r == 0) { moved = true; swp(v, a, b); a++; }
            b++;
          }
          while (b <= c && (r = cmp(v[c], v[0])) >= 0) {
            if (r == 0) { moved = true; swp(v, c, d); d--; }
            c--;
          }
          if (b > c) break;
          swp(v, b, c);
          moved = true;
          b++; c--;
        }
        if (!moved) { insertion(v, n); break; }
        long r = a < b - a ? a : b - a;       // equal-to-pivot runs to the middle
        for (long i = 0; i < r; i++) swp(v, i, b - r + i);
        r = d - c < n - d - 1 ? d - c : n - d - 1;
        for (long i = 0; i < r; i++) swp(v, b + i, n - r + i);
        const long left = b - a, right = d - c;
        // two independent sub-ranges: [0, left) and [n - right, n); finish the smaller now
        const bool left_now = left <= right;
        const long later_off = left_now ? n - right : 0, later_n = left_now ? right : left;
        const long now_off = left_now ? 0 : n - right, now_n = left_now ? left : right;
        if (later_n > 1) {
          if (sp >= kSortStack) return false;
          st_off[sp] = base + off + later_off; st_n[sp] = later_n; sp++;
        }
        if (now_n > 1) { v += now_off; off += now_off; n = now_n; } else break;
      }
    }
    return true;
  }
};

// One integer comparison per column pair: every code is replaced by the dense rank of its
// probability (equal probabilities -- the two tails of a symmetric insert model -- share a rank),
// isoform 0 most significant.  The comparison results are exactly those of the reference's
// lexicographic comparison of the probability columns.
template <class KeyT>
struct KeyCmp {
  const KeyT *key;
  MISOB200_HD_SORT int operator()(int32_t a, int32_t b) const {
    const KeyT x = key[a], y = key[b];
    return x < y ? -1 : (x > y ? 1 : 0);
  }
};

}  // namespace misob200
