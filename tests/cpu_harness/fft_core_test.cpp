// CPU simulation of the device line-FFT stages (fft_core.h): reads n, nlines, dir and complex data from stdin,
// runs nat2perm (dir=+1) or perm2nat (dir=-1) with `nthreads` simulated threads, prints the result.
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../sternheimergw_b200/csrc/fft_core.h"
using namespace sgw;
int main() {
  int n, nlines, dir, nthreads, strided_layout;
  if (scanf("%d %d %d %d %d", &n, &nlines, &dir, &nthreads, &strided_layout) != 5) return 1;
  Plan1D p;
  if (!make_plan(n, &p)) { printf("NOPLAN\n"); return 0; }
  std::vector<double2> tw(n);
  for (int m = 0; m < n; ++m) { tw[m].x = cos(-2.0 * M_PI * m / n); tw[m].y = sin(-2.0 * M_PI * m / n); }
  // layout: strided_layout=0: line l contiguous with pitch n+1 ; =1: element stride = nlines+1 (lines adjacent)
  int ls = strided_layout ? 1 : n + 1, es = strided_layout ? nlines + 1 : 1;
  std::vector<double2> x((size_t)(n + 1) * (nlines + 1));
  for (int l = 0; l < nlines; ++l)
    for (int e = 0; e < n; ++e) { double a, b; if (scanf("%lf %lf", &a, &b) != 2) return 2; x[(size_t)l * ls + (size_t)e * es] = {a, b}; }
  if (dir == 0) {  // fused local-potential product along a contiguous axis: inverse, x v (permuted order), forward
    std::vector<double> v((size_t)n * nlines);
    for (auto &q : v) if (scanf("%lf", &q) != 1) return 3;
    if (p.r2 > 1) {
      for (int t = 0; t < nthreads; ++t) run_strided<+1>(p.r1, x.data(), nlines, nullptr, ls, es, p.r2, tw.data(), true, t, nthreads);
      for (int t = 0; t < nthreads; ++t) run_mid<false>(p.r2, x.data(), nlines, ls, p.r1, tw.data(), true, v.data(), nullptr, n, t, nthreads);
      for (int t = 0; t < nthreads; ++t) run_strided<-1>(p.r1, x.data(), nlines, nullptr, ls, es, p.r2, tw.data(), false, t, nthreads);
    } else {
      for (int t = 0; t < nthreads; ++t) run_mid<false>(p.r1, x.data(), nlines, ls, 1, tw.data(), false, v.data(), nullptr, n, t, nthreads);
    }
  } else if (dir > 0) {   // nat2perm, inverse codelets
    for (int t = 0; t < nthreads; ++t) run_strided<+1>(p.r1, x.data(), nlines, nullptr, ls, es, p.r2, tw.data(), p.r2 > 1, t, nthreads);
    if (p.r2 > 1) for (int t = 0; t < nthreads; ++t) run_contig<+1>(p.r2, x.data(), nlines, nullptr, ls, es, p.r1, tw.data(), false, t, nthreads);
  } else {         // perm2nat, forward codelets
    if (p.r2 > 1) for (int t = 0; t < nthreads; ++t) run_contig<-1>(p.r2, x.data(), nlines, nullptr, ls, es, p.r1, tw.data(), true, t, nthreads);
    for (int t = 0; t < nthreads; ++t) run_strided<-1>(p.r1, x.data(), nlines, nullptr, ls, es, p.r2, tw.data(), false, t, nthreads);
  }
  printf("%d %d\n", p.r1, p.r2);
  for (int l = 0; l < nlines; ++l)
    for (int e = 0; e < n; ++e) printf("%.17g %.17g\n", x[(size_t)l * ls + (size_t)e * es].x, x[(size_t)l * ls + (size_t)e * es].y);
  return 0;
}
