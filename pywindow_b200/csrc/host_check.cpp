// Host build of the small optimisers (lbfgsb_small.cuh, neldermead2.cuh) with a scalar pore
// objective, so the restatement can be checked against scipy on a CPU-only machine
// (tests/test_host_optimisers.py).  Not part of the GPU product path: libpwb.so never links this.
#include "lbfgsb_small.cuh"
#include "neldermead2.cuh"

static double* g_trace = nullptr; static int g_trace_cap = 0, g_trace_n = 0;
extern "C" void pwbh_set_trace(double* buf, int cap) { g_trace = buf; g_trace_cap = cap; g_trace_n = 0; }
extern "C" int pwbh_trace_count() { return g_trace_n; }
static void trace3(const double* x, double f) {
  if (g_trace && g_trace_n < g_trace_cap) { double* p = g_trace + 4 * g_trace_n++; p[0] = x[0]; p[1] = x[1]; p[2] = x[2]; p[3] = f; }
}
namespace {
struct Atoms { const double* X; const double* r; int n; };

inline double pore_d(const Atoms& a, double cx, double cy, double cz) {
  double best = 1.0 / 0.0;
  for (int i = 0; i < a.n; i++) {
    double dx = a.X[3 * i] - cx, dy = a.X[3 * i + 1] - cy, dz = a.X[3 * i + 2] - cz;
    double v = sqrt(dx * dx + dy * dy + dz * dz) - a.r[i];
    if (v < best) best = v;
  }
  return 2 * best;
}

struct PoreObj3 {   // utilities.py:391-397 correct_pore_diameter
  Atoms a; const double *lo, *hi; const int* nbd; int nfev;
  void fg(const double* x, double& f, double* g) {
    f = -pore_d(a, x[0], x[1], x[2]); nfev++;
    trace3(x, f);
    for (int i = 0; i < 3; i++) {
      double h = pwb::lb_fd_step(x[i], nbd[i], lo[i], hi[i]);
      double y[3] = {x[0], x[1], x[2]};
      y[i] = x[i] + h;
      double f1 = -pore_d(a, y[0], y[1], y[2]); nfev++;
      g[i] = (f1 - f) / ((x[i] + h) - x[i]);
    }
  }
};
struct ZObj1 {      // utilities.py:1174-1188 optimise_z
  Atoms a; double lo; int nfev;
  void fg(const double* x, double& f, double* g) {
    f = pore_d(a, 0, 0, x[0]); nfev++;
    double h = pwb::lb_fd_step(x[0], 1, lo, 0);
    double f1 = pore_d(a, 0, 0, x[0] + h); nfev++;
    g[0] = (f1 - f) / ((x[0] + h) - x[0]);
  }
};
struct XYObj {      // utilities.py:1164-1171 optimise_xy
  Atoms a; double z;
  double f2(double x, double y) { return -pore_d(a, x, y, z); }
};
}  // namespace

extern "C" {
// out: x[3], f, nit, status, nfev
int pwbh_opt_pore(const double* X, const double* r, int n, const double* com, double* out) {
  Atoms a{X, r, n};
  double pr = pore_d(a, com[0], com[1], com[2]) / 2;
  double lo[3], hi[3], x[3];
  int nbd[3] = {2, 2, 2};
  for (int i = 0; i < 3; i++) { lo[i] = com[i] - pr; hi[i] = com[i] + pr; x[i] = com[i]; }
  if (pr < 0) return 1;
  PoreObj3 o{a, lo, hi, nbd, 0};
  double f; int nit;
  double hist[LB_STORAGE(3)];
  int st = pwb::lbfgsb_minimize<3>(o, x, lo, hi, nbd, 2.220446049250313e-09, 1e-5, 20, 15000, &f, &nit, hist);
  out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = f; out[4] = nit; out[5] = st; out[6] = o.nfev;
  return 0;
}
// out: z, f, nit, status, nfev
int pwbh_opt_z(const double* X, const double* r, int n, double lb, double* out) {
  Atoms a{X, r, n};
  double lo[1] = {lb}, hi[1] = {0}, x[1] = {0.0};
  int nbd[1] = {1};
  ZObj1 o{a, lb, 0};
  double f; int nit;
  double hist[LB_STORAGE(1)];
  int st = pwb::lbfgsb_minimize<1>(o, x, lo, hi, nbd, 2.220446049250313e-09, 1e-5, 20, 15000, &f, &nit, hist);
  out[0] = x[0]; out[1] = f; out[2] = nit; out[3] = st; out[4] = o.nfev;
  return 0;
}
// out: x, y, f, nfev
int pwbh_fmin_xy(const double* X, const double* r, int n, double z, double x0, double y0, double* out) {
  Atoms a{X, r, n};
  XYObj o{a, z};
  double xy[2], f; int nfev;
  pwb::nelder_mead2(o, x0, y0, 1e-4, 1e-4, 400, 400, xy, &f, &nfev);
  out[0] = xy[0]; out[1] = xy[1]; out[2] = f; out[3] = nfev;
  return 0;
}
}
