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

// stable ascending sort of three (f, x, y) vertices -- the insertion sort scipy's np.argsort(kind=stable)
// amounts to for three elements -- on scalars, so the simplex stays in registers on the device
#define PWB_NM_SWAP(a, b) { double t_; t_ = f##a; f##a = f##b; f##b = t_; t_ = x##a; x##a = x##b; x##b = t_; t_ = y##a; y##a = y##b; y##b = t_; }

// Obj must provide: double f2(double x, double y)   (warp-cooperative on device)
//
// Written as a state machine around ONE call site of the objective: on the device the objective is a
// few hundred inlined instructions, and the textbook layout (one call per reflection / expansion /
// contraction / shrink branch) multiplies that by nine in a kernel that already misses in the
// instruction cache.  The sequence of evaluations and comparisons is scipy's, unchanged.
template <class Obj>
PWB_HDN void nelder_mead2(Obj& obj, double xs, double ys, double xatol, double fatol, int maxiter, int maxfun,
                          double* xy_out, double* f_out, int* nfev_out) {
  enum { INIT, REFLECT, EXPAND, CONTRACT_OUT, CONTRACT_IN, SHRINK };
  double x0 = xs, y0 = ys, f0 = 0;                              // vertex 0 (best after sorting)
  double x1 = (xs != 0) ? (1 + 0.05) * xs : 0.00025, y1 = ys, f1 = 0;
  double x2 = xs, y2 = (ys != 0) ? (1 + 0.05) * ys : 0.00025, f2 = 0;
  int fcalls = 0, iterations = 0, st = INIT, k = 0;
  double tx = x0, ty = y0;
  double xb0 = 0, xb1 = 0, xr0 = 0, xr1 = 0, fxr = 0;
  for (;;) {
    const double fv = obj.f2(tx, ty); fcalls++;
    bool end_iter = false, aborted = false, shrink = false;
    if (st == INIT) {
      if (k == 0) f0 = fv; else if (k == 1) f1 = fv; else f2 = fv;
      k++;
      if (k == 1) { tx = x1; ty = y1; }
      else if (k == 2) { tx = x2; ty = y2; }
      else { iterations = 0; end_iter = true; }          // falls into the sort + loop head below
    } else if (st == REFLECT) {
      fxr = fv;
      if (fxr < f0) {
        if (fcalls >= maxfun) { aborted = true; end_iter = true; }
        else { tx = (1 + 2.0) * xb0 - 2.0 * x2; ty = (1 + 2.0) * xb1 - 2.0 * y2; st = EXPAND; }
      } else if (fxr < f1) {
        x2 = xr0; y2 = xr1; f2 = fxr; end_iter = true;
      } else if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else if (fxr < f2) {
        tx = (1 + 0.5) * xb0 - 0.5 * x2; ty = (1 + 0.5) * xb1 - 0.5 * y2; st = CONTRACT_OUT;
      } else {
        tx = (1 - 0.5) * xb0 + 0.5 * x2; ty = (1 - 0.5) * xb1 + 0.5 * y2; st = CONTRACT_IN;
      }
    } else if (st == EXPAND) {
      if (fv < fxr) { x2 = tx; y2 = ty; f2 = fv; }
      else { x2 = xr0; y2 = xr1; f2 = fxr; }
      end_iter = true;
    } else if (st == CONTRACT_OUT) {
      if (fv <= fxr) { x2 = tx; y2 = ty; f2 = fv; end_iter = true; }
      else shrink = true;
    } else if (st == CONTRACT_IN) {
      if (fv < f2) { x2 = tx; y2 = ty; f2 = fv; end_iter = true; }
      else shrink = true;
    } else {   // SHRINK: vertex k was just evaluated
      if (k == 1) f1 = fv; else f2 = fv;
      k++;
      if (k == 3) end_iter = true;
      else if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else {
        x2 = x0 + 0.5 * (x2 - x0);
        y2 = y0 + 0.5 * (y2 - y0);
        tx = x2; ty = y2;
      }
    }
    if (shrink) {
      if (fcalls >= maxfun) { aborted = true; end_iter = true; }
      else {
        x1 = x0 + 0.5 * (x1 - x0);
        y1 = y0 + 0.5 * (y1 - y0);
        tx = x1; ty = y1; k = 1; st = SHRINK;
      }
    }
    if (!end_iter) continue;
    if (!aborted) iterations++;
    if (f0 > f1) PWB_NM_SWAP(0, 1)
    if (f1 > f2) { PWB_NM_SWAP(1, 2) if (f0 > f1) PWB_NM_SWAP(0, 1) }
    if (!(fcalls < maxfun && iterations < maxiter)) break;
    double dx = fmax(fmax(fabs(x1 - x0), fabs(y1 - y0)), fmax(fabs(x2 - x0), fabs(y2 - y0)));
    double df = fmax(fabs(f0 - f1), fabs(f0 - f2));
    if (dx <= xatol && df <= fatol) break;
    xb0 = (x0 + x1) / 2; xb1 = (y0 + y1) / 2;
    xr0 = (1 + 1.0) * xb0 - 1.0 * x2; xr1 = (1 + 1.0) * xb1 - 1.0 * y2;
    tx = xr0; ty = xr1; st = REFLECT;
  }
  xy_out[0] = x0; xy_out[1] = y0;
  *f_out = fmin(f0, fmin(f1, f2));
  *nfev_out = fcalls;
}
#undef PWB_NM_SWAP

}  // namespace pwb
