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
//
// Written as a state machine around ONE call site of the objective: on the device the objective is a
// few hundred inlined instructions, and the textbook layout (one call per reflection / expansion /
// contraction / shrink branch) multiplies that by nine in a kernel that already misses in the
// instruction cache.  The sequence of evaluations and comparisons is scipy's, unchanged.
template <class Obj>
PWB_HDN void nelder_mead2(Obj& obj, double x0, double y0, double xatol, double fatol, int maxiter, int maxfun,
                          double* xy_out, double* f_out, int* nfev_out) {
  enum { INIT, REFLECT, EXPAND, CONTRACT_OUT, CONTRACT_IN, SHRINK };
  double sim[3][2], fs[3];
  sim[0][0] = x0; sim[0][1] = y0;
  sim[1][0] = (x0 != 0) ? (1 + 0.05) * x0 : 0.00025; sim[1][1] = y0;
  sim[2][0] = x0; sim[2][1] = (y0 != 0) ? (1 + 0.05) * y0 : 0.00025;
  int fcalls = 0, iterations = 0, st = INIT, k = 0;
  double tx = sim[0][0], ty = sim[0][1];
  double xb0 = 0, xb1 = 0, xr0 = 0, xr1 = 0, fxr = 0;
  for (;;) {
    const double fv = obj.f2(tx, ty); fcalls++;
    bool end_iter = false, aborted = false, shrink = false;
    if (st == INIT) {
      if (k == 0) fs[0] = fv; else if (k == 1) fs[1] = fv; else fs[2] = fv;
      k++;
      if (k == 1) { tx = sim[1][0]; ty = sim[1][1]; }
      else if (k == 2) { tx = sim[2][0]; ty = sim[2][1]; }
      else { iterations = 0; end_iter = true; }          // falls into the sort + loop head below
    } else if (st == REFLECT) {
      fxr = fv;
      if (fxr < fs[0]) {
        if (fcalls >= maxfun) { aborted = true; end_iter = true; }
        else { tx = (1 + 2.0) * xb0 - 2.0 * sim[2][0]; ty = (1 + 2.0) * xb1 - 2.0 * sim[2][1]; st = EXPAND; }
      } else if (fxr < fs[1]) {
        sim[2][0] = xr0; sim[2][1] = xr1; fs[2] = fxr; end_iter = true;
      } else if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else if (fxr < fs[2]) {
        tx = (1 + 0.5) * xb0 - 0.5 * sim[2][0]; ty = (1 + 0.5) * xb1 - 0.5 * sim[2][1]; st = CONTRACT_OUT;
      } else {
        tx = (1 - 0.5) * xb0 + 0.5 * sim[2][0]; ty = (1 - 0.5) * xb1 + 0.5 * sim[2][1]; st = CONTRACT_IN;
      }
    } else if (st == EXPAND) {
      if (fv < fxr) { sim[2][0] = tx; sim[2][1] = ty; fs[2] = fv; }
      else { sim[2][0] = xr0; sim[2][1] = xr1; fs[2] = fxr; }
      end_iter = true;
    } else if (st == CONTRACT_OUT) {
      if (fv <= fxr) { sim[2][0] = tx; sim[2][1] = ty; fs[2] = fv; end_iter = true; }
      else shrink = true;
    } else if (st == CONTRACT_IN) {
      if (fv < fs[2]) { sim[2][0] = tx; sim[2][1] = ty; fs[2] = fv; end_iter = true; }
      else shrink = true;
    } else {   // SHRINK: vertex k was just evaluated
      if (k == 1) fs[1] = fv; else fs[2] = fv;
      k++;
      if (k == 3) end_iter = true;
      else if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else {
        sim[2][0] = sim[0][0] + 0.5 * (sim[2][0] - sim[0][0]);
        sim[2][1] = sim[0][1] + 0.5 * (sim[2][1] - sim[0][1]);
        tx = sim[2][0]; ty = sim[2][1];
      }
    }
    if (shrink) {
      if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else {
        sim[1][0] = sim[0][0] + 0.5 * (sim[1][0] - sim[0][0]);
        sim[1][1] = sim[0][1] + 0.5 * (sim[1][1] - sim[0][1]);
        tx = sim[1][0]; ty = sim[1][1]; k = 1; st = SHRINK;
      }
    }
    if (!end_iter) continue;
    if (!aborted) iterations++;
    Nm2Sort::sort3(fs, sim);
    if (!(fcalls < maxfun && iterations < maxiter)) break;
    double dx = fmax(fmax(fabs(sim[1][0] - sim[0][0]), fabs(sim[1][1] - sim[0][1])),
                     fmax(fabs(sim[2][0] - sim[0][0]), fabs(sim[2][1] - sim[0][1])));
    double df = fmax(fabs(fs[0] - fs[1]), fabs(fs[0] - fs[2]));
    if (dx <= xatol && df <= fatol) break;
    xb0 = (sim[0][0] + sim[1][0]) / 2; xb1 = (sim[0][1] + sim[1][1]) / 2;
    xr0 = (1 + 1.0) * xb0 - 1.0 * sim[2][0]; xr1 = (1 + 1.0) * xb1 - 1.0 * sim[2][1];
    tx = xr0; ty = xr1; st = REFLECT;
  }
  xy_out[0] = sim[0][0]; xy_out[1] = sim[0][1];
  *f_out = fmin(fs[0], fmin(fs[1], fs[2]));
  *nfev_out = fcalls;
}

}  // namespace pwb
