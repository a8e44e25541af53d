// fft_core.h -- radix plans and per-thread line-FFT stages shared by the sm_100a kernels (fft.cu) and the
// CPU unit test (tests/cpu_harness/fft_core_test.cpp).  Plain C++: SGW_HD is __host__ __device__ under nvcc.
//
// A 1-D length-n transform (n = r1*r2, radices from fft_codelets.h) is done in place in two stages with
// no digit-reversal pass:
//   nat2perm  (used for G->r):  [strided DFT_r1 + twiddle]  ->  [contiguous DFT_r2]
//             natural-order input, output position p holds natural index  perm(p) = p / r2 + r1 * (p % r2)
//   perm2nat  (used for r->G):  [contiguous DFT_r2 + twiddle] -> [strided DFT_r1]
//             input in that same permuted order, natural-order output.
// Each butterfly touches only its own index set, so a stage needs no intra-stage synchronisation; real
// space therefore lives in "permuted" order everywhere inside the library (the local potential and all
// real-space fields are stored pre-permuted), which is invisible at the C-ABI.
#pragma once
#include "fft_codelets.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
struct double2 { double x, y; };
#endif

namespace sgw {

struct Plan1D {
  int n, r1, r2;
};

inline bool radix_ok(int r) {
  switch (r) {
    case 1: case 2: case 3: case 4: case 5: case 6: case 8: case 9: case 10: case 12: case 15: case 16:
    case 18: case 20: case 24: case 25: case 27: case 30: case 32:
    case 7: case 11: case 14: case 21: case 22: case 28: return true;
    default: return false;
  }
}

// balanced two-radix plan; returns false if n is not representable.  Every n = 2^a 3^b 5^c <= 960 is (radices up to 32), and
// so are the lengths with ONE factor 7 and/or 11 that QE's good_fft_order hands out (14, 21, 22, 28, 42, 44, 56, 63, 66, 70, 77,
// 84, 88, ...) through the radices 7, 11, 14, 21, 22, 28.  The radices above 16 and the 7/11 family only serve lengths that have
// no plan without them: they are dispatched separately (run_*_big) and may spill under the kernels' register caps --
// correct, slower per point.
inline bool make_plan(int n, Plan1D* p) {
  p->n = n;
  if (n <= 16 && radix_ok(n)) { p->r1 = n; p->r2 = 1; return true; }
  int best = -1, bestmax = 1 << 30;
  for (int r1 = 2; r1 <= 32; ++r1) {
    if (n % r1 || !radix_ok(r1)) continue;
    int r2 = n / r1;
    if (r2 > 32 || !radix_ok(r2) || r2 < 2) continue;
    int m = r1 > r2 ? r1 : r2;
    if (m < bestmax || (m == bestmax && r1 < r2)) { bestmax = m; best = r1; }
  }
  if (best < 0) return false;
  p->r1 = best; p->r2 = n / best;
  return true;
}

// plan made of the common radices only (the ones the kernels run without spills): what a grid the library is free to choose
// (the reduced Delta-rho box) should have
inline bool plan_is_fast(const Plan1D& p) {
  auto common = [](int r) { return r <= 16 && r != 7 && r != 11 && r != 13 && r != 14; };
  return common(p.r1) && common(p.r2);
}

SGW_HD int perm_index(int r1, int r2, int pos) { return pos / r2 + r1 * (pos % r2); }

// ---- one stage over a set of lines; thread `tid` of `nthreads` takes tasks tid, tid+nthreads, ...
// line l starts at x + (line_ids ? line_ids[l] : l) * ls ; element e of a line at + e * es.
// Task t = (line l = t % nlines, sub-index t / nlines); the pair is advanced incrementally (one division per
// stage call instead of two per task) and all offsets are 32-bit: the integer pipe, not FP64, was the top
// issuer in the first profile of the plane kernel.
struct TaskIter {
  int l, j, dl, dj, nlines;
  SGW_HD TaskIter(int tid, int nthreads, int nl) : l(tid % nl), j(tid / nl), dl(nthreads % nl), dj(nthreads / nl), nlines(nl) {}
  SGW_HD void next() {
    l += dl; j += dj;
    if (l >= nlines) { l -= nlines; ++j; }
  }
};

template <int R, int DIR>
SGW_HD void stage_strided(double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                          const double2* tw, bool do_tw, int tid, int nthreads) {
  const int ntasks = nlines * r_other;
  const int st = r_other * es;
  TaskIter it(tid, nthreads, nlines);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    const int j2 = it.j;
    double2* base = x + ((line_ids ? line_ids[it.l] : it.l) * ls + j2 * es);
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const double2 v = base[j * st]; re[j] = v.x; im[j] = v.y; }
    if (DIR < 0) dft_fwd<R>(re, im); else dft_fwd<R>(im, re);
    if (do_tw) {
#pragma unroll
      for (int k = 1; k < R; ++k) {
        const double2 w = tw[j2 * k];
        const double c = w.x, s = DIR < 0 ? w.y : -w.y;
        const double a = re[k], b = im[k];
        re[k] = a * c - b * s;
        im[k] = a * s + b * c;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) { double2 v; v.x = re[k]; v.y = im[k]; base[k * st] = v; }
  }
}

template <int R, int DIR>
SGW_HD void stage_contig(double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                         const double2* tw, bool do_tw, int tid, int nthreads) {
  const int ntasks = nlines * r_other;
  TaskIter it(tid, nthreads, nlines);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    const int a = it.j;
    double2* base = x + ((line_ids ? line_ids[it.l] : it.l) * ls + a * R * es);
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const double2 v = base[j * es]; re[j] = v.x; im[j] = v.y; }
    if (DIR < 0) dft_fwd<R>(re, im); else dft_fwd<R>(im, re);
    if (do_tw) {
#pragma unroll
      for (int k = 1; k < R; ++k) {
        const double2 w = tw[a * k];
        const double c = w.x, s = DIR < 0 ? w.y : -w.y;
        const double p = re[k], q = im[k];
        re[k] = p * c - q * s;
        im[k] = p * s + q * c;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) { double2 v; v.x = re[k]; v.y = im[k]; base[k * es] = v; }
  }
}

// Fused middle of the local-potential product along the contiguous axis: the LAST inverse stage (contiguous DFT_R,
// no twiddle), the point-wise multiplication by the potential, and the FIRST forward stage (contiguous DFT_R +
// twiddle) act on the same R elements of a row, so they are done in one register round trip.  Arithmetic per
// element is exactly that of stage_contig<R,+1> -> multiply -> stage_contig<R,-1>.
// v (real) or f (complex) are indexed like the plane without padding: row l at + l * vls.
template <int R, bool CPLX>
SGW_HD void stage_mid(double2* x, int nlines, int ls, int r_other, const double2* tw, bool do_tw, const double* v,
                      const double2* f, int vls, int tid, int nthreads) {
  const int ntasks = nlines * r_other;
  TaskIter it(tid, nthreads, nlines);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    const int l = it.l, a = it.j;
    double2* base = x + (l * ls + a * R);
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const double2 w = base[j]; re[j] = w.x; im[j] = w.y; }
    dft_fwd<R>(im, re);
    if (CPLX) {
      const double2* fr = f + (l * vls + a * R);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double2 q = fr[j];
        const double p = re[j], s = im[j];
        re[j] = q.x * p - q.y * s;
        im[j] = q.x * s + q.y * p;
      }
    } else {
      const double* vr = v + (l * vls + a * R);
#pragma unroll
      for (int j = 0; j < R; ++j) { const double q = vr[j]; re[j] *= q; im[j] *= q; }
    }
    dft_fwd<R>(re, im);
    if (do_tw) {
#pragma unroll
      for (int k = 1; k < R; ++k) {
        const double2 w = tw[a * k];
        const double c = w.x, s = w.y;
        const double p = re[k], q = im[k];
        re[k] = p * c - q * s;
        im[k] = p * s + q * c;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) { double2 w; w.x = re[k]; w.y = im[k]; base[k] = w; }
  }
}

// runtime radix dispatch (uniform across the block).  The radices above 16 live in their own functions: their codelets
// spill under the kernels' register caps, and keeping them out of the common dispatchers leaves those unchanged.
#ifdef __CUDACC__
#define SGW_DISPATCH __host__ __device__ __noinline__
#else
#define SGW_DISPATCH inline
#endif

template <int DIR>
SGW_DISPATCH void run_strided_big(int R, double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                                  const double2* tw, bool do_tw, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_strided<r, DIR>(x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
    SGW_FOR_EACH_BIG_RADIX(SGW_CASE)
#undef SGW_CASE
    default: break;
  }
}

template <int DIR>
SGW_DISPATCH void run_strided(int R, double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                              const double2* tw, bool do_tw, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_strided<r, DIR>(x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
    SGW_FOR_EACH_RADIX(SGW_CASE)
#undef SGW_CASE
    default: run_strided_big<DIR>(R, x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
  }
}

template <int DIR>
SGW_DISPATCH void run_contig_big(int R, double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                                 const double2* tw, bool do_tw, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_contig<r, DIR>(x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
    SGW_FOR_EACH_BIG_RADIX(SGW_CASE)
#undef SGW_CASE
    default: break;
  }
}

template <int DIR>
SGW_DISPATCH void run_contig(int R, double2* x, int nlines, const int* line_ids, int ls, int es, int r_other,
                             const double2* tw, bool do_tw, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_contig<r, DIR>(x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
    SGW_FOR_EACH_RADIX(SGW_CASE)
#undef SGW_CASE
    default: run_contig_big<DIR>(R, x, nlines, line_ids, ls, es, r_other, tw, do_tw, tid, nthreads); break;
  }
}

template <bool CPLX>
SGW_DISPATCH void run_mid_big(int R, double2* x, int nlines, int ls, int r_other, const double2* tw, bool do_tw, const double* v,
                              const double2* f, int vls, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_mid<r, CPLX>(x, nlines, ls, r_other, tw, do_tw, v, f, vls, tid, nthreads); break;
    SGW_FOR_EACH_BIG_RADIX(SGW_CASE)
#undef SGW_CASE
    default: break;
  }
}

template <bool CPLX>
SGW_DISPATCH void run_mid(int R, double2* x, int nlines, int ls, int r_other, const double2* tw, bool do_tw, const double* v,
                          const double2* f, int vls, int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_mid<r, CPLX>(x, nlines, ls, r_other, tw, do_tw, v, f, vls, tid, nthreads); break;
    SGW_FOR_EACH_RADIX(SGW_CASE)
#undef SGW_CASE
    default: run_mid_big<CPLX>(R, x, nlines, ls, r_other, tw, do_tw, v, f, vls, tid, nthreads); break;
  }
}

}  // namespace sgw
