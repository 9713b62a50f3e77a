// Two-variable Nelder-Mead simplex search, host+device.
//
// The reference finishes its 20x20 brute-force window search with `scipy.optimize.fmin`
// (utilities.py:1312, `brute(..., finish=fmin)`), i.e. scipy's Nelder-Mead with xtol=ftol=1e-4,
// maxiter=maxfun=200*N, reflection/expansion/contraction/shrink = 1, 2, 0.5, 0.5, initial simplex
// x0 and x0 with one coordinate scaled by 1.05 (0.00025 if that coordinate is 0), and a *stable*
// ascending sort of the three vertices after every iteration
// (scipy/optimize/_optimize.py:_minimize_neldermead).  Restated here for N = 2.
#pragma once
#include "lbfgsb_small.cuh"

namespace pwb {

struct Nm2Sort {
  // stable insertion sort of three (f, x, y) vertices by f ascending
  PWB_HD static void sort3(double* f, double (*s)[2]) {
    for (int i = 1; i < 3; i++) {
      double fi = f[i], xi = s[i][0], yi = s[i][1];
      int j = i - 1;
      while (j >= 0 && f[j] > fi) {
        f[j + 1] = f[j]; s[j + 1][0] = s[j][0]; s[j + 1][1] = s[j][1];
        j--;
      }
      f[j + 1] = fi; s[j + 1][0] = xi; s[j + 1][1] = yi;
    }
  }
};

// Obj must provide: double f2(double x, double y)   (warp-cooperative on device)
template <class Obj>
PWB_HDN void nelder_mead2(Obj& obj, double x0, double y0, double xatol, double fatol, int maxiter, int maxfun,
                          double* xy_out, double* f_out, int* nfev_out) {
  double sim[3][2], fs[3];
  sim[0][0] = x0; sim[0][1] = y0;
  sim[1][0] = (x0 != 0) ? (1 + 0.05) * x0 : 0.00025; sim[1][1] = y0;
  sim[2][0] = x0; sim[2][1] = (y0 != 0) ? (1 + 0.05) * y0 : 0.00025;
  int fcalls = 0;
  for (int k = 0; k < 3; k++) { fs[k] = obj.f2(sim[k][0], sim[k][1]); fcalls++; }
  Nm2Sort::sort3(fs, sim);
  int iterations = 1;
  while (fcalls < maxfun && iterations < maxiter) {
    double dx = fmax(fmax(fabs(sim[1][0] - sim[0][0]), fabs(sim[1][1] - sim[0][1])),
                     fmax(fabs(sim[2][0] - sim[0][0]), fabs(sim[2][1] - sim[0][1])));
    double df = fmax(fabs(fs[0] - fs[1]), fabs(fs[0] - fs[2]));
    if (dx <= xatol && df <= fatol) break;
    double xb0 = (sim[0][0] + sim[1][0]) / 2, xb1 = (sim[0][1] + sim[1][1]) / 2;
    double xr0 = (1 + 1.0) * xb0 - 1.0 * sim[2][0], xr1 = (1 + 1.0) * xb1 - 1.0 * sim[2][1];
    double fxr = obj.f2(xr0, xr1); fcalls++;
    bool doshrink = false, aborted = false;
    if (fxr < fs[0]) {
      if (fcalls >= maxfun) aborted = true;
      else {
        double xe0 = (1 + 2.0) * xb0 - 2.0 * sim[2][0], xe1 = (1 + 2.0) * xb1 - 2.0 * sim[2][1];
        double fxe = obj.f2(xe0, xe1); fcalls++;
        if (fxe < fxr) { sim[2][0] = xe0; sim[2][1] = xe1; fs[2] = fxe; }
        else { sim[2][0] = xr0; sim[2][1] = xr1; fs[2] = fxr; }
      }
    } else if (fxr < fs[1]) {
      sim[2][0] = xr0; sim[2][1] = xr1; fs[2] = fxr;
    } else {
      if (fcalls >= maxfun) aborted = true;
      else if (fxr < fs[2]) {
        double xc0 = (1 + 0.5) * xb0 - 0.5 * sim[2][0], xc1 = (1 + 0.5) * xb1 - 0.5 * sim[2][1];
        double fxc = obj.f2(xc0, xc1); fcalls++;
        if (fxc <= fxr) { sim[2][0] = xc0; sim[2][1] = xc1; fs[2] = fxc; }
        else doshrink = true;
      } else {
        double xc0 = (1 - 0.5) * xb0 + 0.5 * sim[2][0], xc1 = (1 - 0.5) * xb1 + 0.5 * sim[2][1];
        double fxcc = obj.f2(xc0, xc1); fcalls++;
        if (fxcc < fs[2]) { sim[2][0] = xc0; sim[2][1] = xc1; fs[2] = fxcc; }
        else doshrink = true;
      }
      if (doshrink) {
        for (int j = 1; j < 3; j++) {
          if (fcalls >= maxfun) { aborted = true; break; }
          sim[j][0] = sim[0][0] + 0.5 * (sim[j][0] - sim[0][0]);
          sim[j][1] = sim[0][1] + 0.5 * (sim[j][1] - sim[0][1]);
          fs[j] = obj.f2(sim[j][0], sim[j][1]); fcalls++;
        }
      }
    }
    if (!aborted) iterations++;
    Nm2Sort::sort3(fs, sim);
  }
  xy_out[0] = sim[0][0]; xy_out[1] = sim[0][1];
  *f_out = fmin(fs[0], fmin(fs[1], fs[2]));
  *nfev_out = fcalls;
}

}  // namespace pwb
