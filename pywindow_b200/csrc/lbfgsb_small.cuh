// Bound-constrained limited-memory BFGS for tiny problems (n <= 3), host+device.
//
// The reference's pore-centre optimiser (utilities.py:400-426) and window z-stage (utilities.py:1301)
// are `scipy.optimize.minimize(..., bounds=...)`, i.e. L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; Zhu et
// al. 1997; Morales & Nocedal 2011) with scipy's defaults m=10, ftol=2.22e-9, gtol=1e-5, maxls=20 and
// 2-point forward-difference gradients with absolute step 1e-8 that flip direction at bounds
// (scipy/optimize/_numdiff.py:_adjust_scheme_to_bounds).  The reference's answer is *where this
// algorithm stops*, not the true optimum, so the algorithm is restated here:
// generalized Cauchy point -> subspace minimisation with projection -> More-Thuente line search
// (dcsrch/dcstep) -> BFGS pair update with the same skip / restart / termination rules.
//
// Design for n <= 3: the compact representation B = theta*I - W M W^T (2m x 2m middle matrix) is
// replaced by the *same* matrix formed densely: B0 = theta*I followed by the stored (s, y) BFGS
// updates, oldest first.  This is algebraically identical (Byrd-Nocedal-Schnabel 1994, Thm 2.3) and
// keeps all state in registers / a 480-byte ring instead of the 1100-double workspace.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PWB_HD __host__ __device__ __forceinline__
#define PWB_HDN __host__ __device__
#else
#define PWB_HD inline
#define PWB_HDN
#endif

namespace pwb {

// On the device the optimiser kernels are bound by INSTRUCTION SUPPLY and by the latency of their dependent FP64 /
// shared-memory chains (ncu: 65-78 % instruction-cache hit rate, 3-5 of 7 warps per scheduler waiting for
// instructions, one instruction per ~10 cycles per warp).  Two measures keep the code small without lengthening
// the executed path (profiles/r2_opt_pore_variants.md): the cold per-variable loops of the Cauchy / subspace /
// projection steps stay ROLLED (arrays live in shared memory anyway), while divisions and square roots stay INLINE
// -- one out-of-line copy of each (LB_NOINLINE_MATH) shrank one search from 80 KB to 58 KB of code and raised the
// I-cache hit rate, but the extra calls and moves cost 24 % more executed instructions and the kernel got slower.
// Same IEEE operations in the same order in every variant: results are unchanged bit for bit.
#if defined(__CUDA_ARCH__) && defined(LB_NOINLINE_MATH)
__device__ __noinline__ double lb_div(double a, double b) { return a / b; }
__device__ __noinline__ double lb_sqrt(double a) { return sqrt(a); }
#else
PWB_HD double lb_div(double a, double b) { return a / b; }
PWB_HD double lb_sqrt(double a) { return sqrt(a); }
#endif
#if defined(__CUDA_ARCH__) && !defined(LB_UNROLLED)
#define LB_ROLLED _Pragma("unroll 1")
#else
#define LB_ROLLED
#endif

constexpr int LB_M = 10;
// doubles of caller storage lbfgsb_minimize<N> needs: (s, y) ring, z / d / t / r, line-search record
// and the small work arrays of the Cauchy / subspace steps (x, g, iwhere: 3N; B: N*N; scratch: 4N + 3N*N)
#define LB_HISTORY(N) (2 * pwb::LB_M * (N))                                      /* the (s, y) ring */
#define LB_STATE(N) (4 * (N) + 20 + 10 * (N) + 4 * (N) * (N) + 8)                /* everything else = LbSolver<N>::STATE */
#define LB_STORAGE(N) (LB_STATE(N) + LB_HISTORY(N))                              /* = LbSolver<N>::SIZE */
constexpr double LB_EPSMCH = 2.220446049250313e-16;
constexpr double LB_FD_STEP = 1e-8;

enum LbStatus {
  LB_CONV_PGTOL = 0,
  LB_CONV_FTOL = 1,
  LB_ABNORMAL = 2,
  LB_MAXITER = 3,
};

// scipy _numdiff.py: absolute step h adjusted to the box so that x+h stays feasible.
PWB_HD double lb_fd_step(double x, int nbd, double lo, double hi) {
  double h = LB_FD_STEP;
  if (nbd == 0) return h;
  // boxed variable with room for the step on one side: the general rule below reduces to a sign flip
  if (nbd == 2 && hi - lo > 1e-6 && x >= lo && x <= hi) return (x + h > hi) ? -h : h;
  const double INF = 1.0 / 0.0;
  double lb = (nbd == 1 || nbd == 2) ? lo : -INF;
  double ub = (nbd == 2 || nbd == 3) ? hi : INF;
  double lower_dist = x - lb, upper_dist = ub - x;
  double xt = x + h;
  bool violated = (xt < lb) || (xt > ub);
  bool fitting = fabs(h) <= fmax(lower_dist, upper_dist);
  if (violated && fitting) h = -h;
  if (!fitting) h = (upper_dist >= lower_dist) ? upper_dist : -lower_dist;
  return h;
}

// ---- More-Thuente safeguarded step (MINPACK-2 dcstep; scipy/optimize/_dcsrch.py:dcstep)
// The four cases of dcstep each carry their own cubic (one sqrt, five to seven divisions).  Here the
// cubic is evaluated ONCE on case-selected operands -- on the GPU the four inlined copies were a
// quarter of each optimiser kernel's code, and those kernels are instruction-cache bound.  Every
// expression keeps the operation order of the original; case 4 uses 3(fy-fp)/(stp-sty), which is
// 3(fp-fy)/(sty-stp) with both signs flipped and therefore rounds identically.
PWB_HD void lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy,
                      double& stp, double fp, double dp, bool& brackt, double stpmin, double stpmax) {
  const double sgn_dp = (dp > 0) - (dp < 0), sgn_dx = (dx > 0) - (dx < 0);
  const double sgnd = sgn_dp * sgn_dx;
  const int cs = (fp > fx) ? 1 : (sgnd < 0.0) ? 2 : (fabs(dp) < fabs(dx)) ? 3 : 4;
  const bool far_end = cs == 4;                       // cubic through (sty, stp) instead of (stx, stp)
  const double fa = far_end ? fy : fx, sta = far_end ? sty : stx, da = far_end ? dy : dx;
  const double theta = lb_div(3.0 * (fa - fp), stp - sta) + da + dp;
  const double s = fmax(fabs(theta), fmax(fabs(da), fabs(dp)));
  const double ths = lb_div(theta, s);
  double rad = ths * ths - lb_div(da, s) * lb_div(dp, s);
  if (cs == 3) rad = fmax(0.0, rad);
  double gamma = s * lb_sqrt(rad);
  if (cs == 1 ? (stp < stx) : (stp > sta)) gamma = -gamma;
  const double e = cs == 1 ? dx : dp;
  const double p = (gamma - e) + theta;
  const double q = cs == 3 ? (gamma + (dx - dp)) + gamma : ((gamma - e) + gamma) + (cs == 1 ? dp : da);
  const double r = lb_div(p, q);
  double stpf;
  if (cs == 1) {
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + (lb_div(dx, lb_div(fx - fp, stp - stx) + dx) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) <= fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (cs == 4) {
    if (brackt) stpf = stp + r * (sty - stp);
    else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  } else {
    const double stpq = stp + lb_div(dp, dp - dx) * (stx - stp);
    double stpc = stp + r * (stx - stp);
    if (cs == 2) {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      brackt = true;
    } else {
      if (!(r < 0 && gamma != 0)) stpc = (stp > stx) ? stpmax : stpmin;
      if (brackt) {
        stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
        if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
        else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
      } else {
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        stpf = fmin(fmax(stpf, stpmin), stpmax);
      }
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// ---- dcsrch as a resumable state machine (MINPACK-2; scipy/optimize/_dcsrch.py:_iterate)
struct LbLineSearch {
  double ftol, gtol, xtol, stpmin, stpmax;
  bool brackt;
  int stage;
  double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
  enum Task { FG = 0, CONV = 1, WARN = 2, ERR = 3 };

  PWB_HD int start(double& stp, double f, double g) {
    if (stp < stpmin || stp > stpmax || g >= 0 || stpmax < stpmin) return ERR;
    brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
    width = stpmax - stpmin; width1 = width * 2.0;
    stx = 0; fx = finit; gx = ginit; sty = 0; fy = finit; gy = ginit;
    stmin = 0; stmax = stp + 4.0 * stp;
    return FG;
  }
  PWB_HD int iterate(double& stp, double f, double g) {
    double ftest = finit + stp * gtest;
    if (stage == 1 && f <= ftest && g >= 0) stage = 2;
    int task = FG;
    if (brackt && (stp <= stmin || stp >= stmax)) task = WARN;
    if (brackt && stmax - stmin <= xtol * stmax) task = WARN;
    if (stp == stpmax && f <= ftest && g <= gtest) task = WARN;
    if (stp == stpmin && (f > ftest || g >= gtest)) task = WARN;
    if (f <= ftest && fabs(g) <= gtol * -ginit) task = CONV;
    if (task != FG) return task;
    // stage 1 works on the modified function f - stp*gtest; one dcstep call site serves both forms
    const bool modified = stage == 1 && f <= fx && f > ftest;
    double fm = f, gm = g;
    if (modified) {
      fm = f - stp * gtest; fx = fx - stx * gtest; fy = fy - sty * gtest;
      gm = g - gtest; gx = gx - gtest; gy = gy - gtest;
    }
    lb_dcstep(stx, fx, gx, sty, fy, gy, stp, fm, gm, brackt, stmin, stmax);
    if (modified) { fx = fx + stx * gtest; fy = fy + sty * gtest; gx = gx + gtest; gy = gy + gtest; }
    if (brackt) {
      if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
      width1 = width;
      width = fabs(sty - stx);
    }
    if (brackt) { stmin = fmin(stx, sty); stmax = fmax(stx, sty); }
    else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
    stp = fmin(fmax(stp, stpmin), stpmax);
    if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
    return FG;
  }
};

static_assert(sizeof(LbLineSearch) <= 20 * sizeof(double), "LB_STORAGE reserves 20 doubles for the line-search record");

// The (s, y) ring lives in caller-provided storage (2*LB_M*N doubles): shared memory on the device, so
// the 32 lanes that run the optimiser redundantly share ONE copy instead of 32 local-memory copies.
template <int N>
struct LbHistory {
  double (*S)[N];
  double (*Y)[N];
  int col, head;   // head = slot of the oldest pair
};

// B = theta*I, then BFGS updates with the stored pairs (oldest -> newest).
template <int N>
PWB_HD void lb_build_B(const LbHistory<N>& H, double theta, double* Bm) {
  double B[N][N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) B[i][j] = (i == j) ? theta : 0.0;
#ifdef LB_READ_AHEAD
  // the next pair is read while the current one is applied (for a ring kept in global memory)
  double sn[N], yn[N];
  if (H.col > 0)
    for (int i = 0; i < N; i++) { sn[i] = H.S[H.head][i]; yn[i] = H.Y[H.head][i]; }
#endif
  for (int k = 0; k < H.col; k++) {
    double s[N], y[N];
#ifdef LB_READ_AHEAD
    for (int i = 0; i < N; i++) { s[i] = sn[i]; y[i] = yn[i]; }
    if (k + 1 < H.col) {
      const int pn = (H.head + k + 1) % LB_M;
      for (int i = 0; i < N; i++) { sn[i] = H.S[pn][i]; yn[i] = H.Y[pn][i]; }
    }
#else
    const int p = (H.head + k) % LB_M;
    for (int i = 0; i < N; i++) { s[i] = H.S[p][i]; y[i] = H.Y[p][i]; }
#endif
    double Bs[N], sBs = 0, ys = 0;
    for (int i = 0; i < N; i++) {
      double a = 0;
      for (int j = 0; j < N; j++) a += B[i][j] * s[j];
      Bs[i] = a;
    }
    for (int i = 0; i < N; i++) { sBs += s[i] * Bs[i]; ys += y[i] * s[i]; }
    const double iys = lb_div(1.0, ys), isBs = lb_div(1.0, sBs);
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) B[i][j] += y[i] * y[j] * iys - Bs[i] * Bs[j] * isBs;
  }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) Bm[i * N + j] = B[i][j];
}

// Solve A d = b for the leading nf x nf SPD system by Cholesky; false if not positive definite.
template <int N>
PWB_HD bool lb_spd_solve(int nf, const double* A, double* b, double* Lm) {   // A, Lm: [N][N] row-major
  LB_ROLLED
  for (int i = 0; i < nf; i++) {
    LB_ROLLED
    for (int j = 0; j <= i; j++) {
      double s = A[i * N + j];
      for (int k = 0; k < j; k++) s -= Lm[i * N + k] * Lm[j * N + k];
      if (i == j) {
        if (!(s > 0)) return false;
        Lm[i * N + i] = lb_sqrt(s);
      } else Lm[i * N + j] = lb_div(s, Lm[j * N + j]);
    }
  }
  LB_ROLLED
  for (int i = 0; i < nf; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= Lm[i * N + k] * b[k];
    b[i] = lb_div(s, Lm[i * N + i]);
  }
  LB_ROLLED
  for (int i = nf - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < nf; k++) s -= Lm[k * N + i] * b[k];
    b[i] = lb_div(s, Lm[i * N + i]);
  }
  return true;
}

// (NBD >= 0: every variable has that bound type and `nbd` is not read; else per-variable types in memory)
template <int N, int NBD>
PWB_HD double lb_projgr(const double* x, const double* g, const double* lo, const double* hi, const int* nbd) {
  double sb = 0;
  LB_ROLLED
  for (int i = 0; i < N; i++) {
    double gi = g[i];
    const int nbi = NBD >= 0 ? NBD : nbd[i];
    if (nbi != 0) {
      if (gi < 0) { if (nbi >= 2) gi = fmax(x[i] - hi[i], gi); }
      else { if (nbi <= 2) gi = fmin(x[i] - lo[i], gi); }
    }
    sb = fmax(sb, fabs(gi));
  }
  return sb;
}

// Generalized Cauchy point along the projected steepest-descent path (cauchy in L-BFGS-B 3.0).
template <int N, int NBD>
PWB_HD void lb_cauchy(const double* x, const double* lo, const double* hi, const int* nbd, const double* g,
                      int* iwhere, const double* B, double theta, double sbgnrm, double* z, double* scratch) {
  for (int i = 0; i < N; i++) z[i] = x[i];
  if (sbgnrm <= 0) return;
  bool bnded = true;
  int nbreak = 0, nunb = 0;
  // work arrays in caller storage: they are indexed by the run-time break-point index
  double* d = scratch; double* tb = d + N; double* zz = tb + N; double* Bz = zz + N;
  int* isb = reinterpret_cast<int*>(Bz + N); int* fixed = isb + N;
  double f1 = 0;
  LB_ROLLED
  for (int i = 0; i < N; i++) {
    double neggi = -g[i], tl = 0, tu = 0;
    const int nbi = NBD >= 0 ? NBD : nbd[i];
    isb[i] = false; fixed[i] = false; tb[i] = 0;
    if (iwhere[i] != 3 && iwhere[i] != -1) {
      if (nbi <= 2) tl = x[i] - lo[i];
      if (nbi >= 2) tu = hi[i] - x[i];
      bool xlower = nbi <= 2 && tl <= 0, xupper = nbi >= 2 && tu <= 0;
      iwhere[i] = 0;
      if (xlower) { if (neggi <= 0) iwhere[i] = 1; }
      else if (xupper) { if (neggi >= 0) iwhere[i] = 2; }
      else { if (fabs(neggi) <= 0) iwhere[i] = -3; }
    }
    if (iwhere[i] != 0 && iwhere[i] != -1) d[i] = 0;
    else {
      d[i] = neggi;
      f1 -= neggi * neggi;
      if (nbi <= 2 && nbi != 0 && neggi < 0) { isb[i] = true; tb[i] = lb_div(tl, -neggi); nbreak++; }
      else if (nbi >= 2 && neggi > 0) { isb[i] = true; tb[i] = lb_div(tu, neggi); nbreak++; }
      else { nunb++; if (fabs(neggi) > 0) bnded = false; }
    }
  }
  if (nbreak == 0 && nunb == 0) return;
  double f2_org = -theta * f1;
  double f2 = 0;
  LB_ROLLED
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) f2 += d[i] * B[i * N + j] * d[j];
  double dtm = lb_div(-f1, f2), tsum = 0;
  if (nbreak > 0) {
    int nleft = nbreak;
    double tj = 0;
    while (true) {
      double tj0 = tj;
      int ibp = -1;
      LB_ROLLED
      for (int i = 0; i < N; i++)
        if (isb[i] && (ibp < 0 || tb[i] < tb[ibp])) ibp = i;
      tj = tb[ibp];
      double dt = tj - tj0;
      if (dtm < dt) break;
      tsum += dt;
      nleft--;
      isb[ibp] = false;
      double dibp = d[ibp];
      d[ibp] = 0;
      fixed[ibp] = true;
      if (dibp > 0) { z[ibp] = hi[ibp]; iwhere[ibp] = 2; }
      else { z[ibp] = lo[ibp]; iwhere[ibp] = 1; }
      if (nleft == 0 && nbreak == N) return;   // every variable sits on a bound
      LB_ROLLED
      for (int i = 0; i < N; i++) zz[i] = fixed[i] ? z[i] - x[i] : tsum * d[i];
      f1 = 0; f2 = 0;
      LB_ROLLED
      for (int i = 0; i < N; i++) {
        double a = 0, b = 0;
        for (int j = 0; j < N; j++) { a += B[i * N + j] * zz[j]; b += B[i * N + j] * d[j]; }
        Bz[i] = a;
        f1 += d[i] * (g[i] + Bz[i]);
        f2 += d[i] * b;
      }
      f2 = fmax(LB_EPSMCH * f2_org, f2);
      if (nleft > 0) { dtm = lb_div(-f1, f2); continue; }
      if (bnded) { f1 = 0; f2 = 0; dtm = 0; }
      else dtm = lb_div(-f1, f2);
      break;
    }
  }
  if (dtm <= 0) dtm = 0;
  tsum += dtm;
  LB_ROLLED
  for (int i = 0; i < N; i++)
    if (!fixed[i]) z[i] = x[i] + tsum * d[i];
}

// Subspace minimisation over the variables free at the Cauchy point (cmprlb + subsm, 3.0 projection).
// Returns false if the reduced Hessian is not positive definite (caller refreshes the memory).
template <int N, int NBD>
PWB_HD bool lb_subsm(const double* x, const double* g, const double* lo, const double* hi, const int* nbd,
                     const int* iwhere, const double* B, double* z, double* scratch) {
  double* zx = scratch; double* rr = zx + N; double* xp = rr + N; double* A = xp + N; double* Lm = A + N * N;
  int* ind = reinterpret_cast<int*>(Lm + N * N);
  int nf = 0;
  for (int i = 0; i < N; i++) if (iwhere[i] <= 0) ind[nf++] = i;
  if (nf == 0) return true;
  for (int i = 0; i < N; i++) zx[i] = z[i] - x[i];
  LB_ROLLED
  for (int a = 0; a < nf; a++) {
    int i = ind[a];
    double s = g[i];
    for (int j = 0; j < N; j++) s += B[i * N + j] * zx[j];
    rr[a] = -s;
    for (int b = 0; b < nf; b++) A[a * N + b] = B[i * N + ind[b]];
  }
  if (!lb_spd_solve<N>(nf, A, rr, Lm)) return false;   // rr now holds d on the free set
  for (int i = 0; i < N; i++) xp[i] = z[i];
  bool iword = false;
  LB_ROLLED
  for (int a = 0; a < nf; a++) {
    int k = ind[a];
    double dk = rr[a], xk = z[k];
    const int nbk = NBD >= 0 ? NBD : nbd[k];
    if (nbk != 0) {
      if (nbk == 1) { z[k] = fmax(lo[k], xk + dk); if (z[k] == lo[k]) iword = true; }
      else if (nbk == 2) {
        xk = fmax(lo[k], xk + dk); z[k] = fmin(hi[k], xk);
        if (z[k] == lo[k] || z[k] == hi[k]) iword = true;
      } else { z[k] = fmin(hi[k], xk + dk); if (z[k] == hi[k]) iword = true; }
    } else z[k] = xk + dk;
  }
  if (!iword) return true;
  double ddp = 0;
  for (int i = 0; i < N; i++) ddp += (z[i] - x[i]) * g[i];
  if (ddp > 0) {
    for (int i = 0; i < N; i++) z[i] = xp[i];
    double alpha = 1, temp1 = 1;
    int ibd = -1;
    LB_ROLLED
    for (int a = 0; a < nf; a++) {
      int k = ind[a];
      double dk = rr[a];
      const int nbk = NBD >= 0 ? NBD : nbd[k];
      if (nbk != 0) {
        if (dk < 0 && nbk <= 2) {
          double t2 = lo[k] - z[k];
          if (t2 >= 0) temp1 = 0; else if (dk * alpha < t2) temp1 = lb_div(t2, dk);
        } else if (dk > 0 && nbk >= 2) {
          double t2 = hi[k] - z[k];
          if (t2 <= 0) temp1 = 0; else if (dk * alpha > t2) temp1 = lb_div(t2, dk);
        }
        if (temp1 < alpha) { alpha = temp1; ibd = a; }
      }
    }
    if (alpha < 1 && ibd >= 0) {
      double dk = rr[ibd];
      int k = ind[ibd];
      if (dk > 0) { z[k] = hi[k]; rr[ibd] = 0; }
      else if (dk < 0) { z[k] = lo[k]; rr[ibd] = 0; }
    }
    for (int a = 0; a < nf; a++) z[ind[a]] += alpha * rr[a];
  }
  return true;
}

// The solver as a RESUMABLE state machine: `init` projects the start point into the box, the caller evaluates
// (f, g) at x -- the ONE objective call site of a kernel -- and calls `advance`, which runs the bookkeeping up to
// the next point that needs an evaluation (a line-search trial point) or to termination.  Several solvers can
// therefore be interleaved by one warp (one per lane group), each pulling new problems when it finishes, and the
// objective code exists once.  Scalars live in the struct (registers); everything indexed by a run-time variable or
// ring slot lives in caller storage (LB_STORAGE(N) doubles: shared memory on the device, one copy per problem).
// NBD >= 0: every variable has that bound type, known at compile time (the bound-type tests fold away).
template <int N, int NBD = -1>
struct LbSolver {
  double* st;     // caller storage, LB_STATE(N) doubles: the state touched on every call lives there (on the device:
                  // shared memory, so a kernel that interleaves several searches keeps one pointer per search)
  double* hs;     // the (s, y) ring, LB_HISTORY(N) doubles: read once per iteration (lb_build_B), one pair written;
                  // behind the state by default, or anywhere else (a kernel short of shared memory keeps it in
                  // global memory: the ring is 36 % of the storage and is read ahead of its use)
  // layout: z d t r | line-search record (20) | x g | B | iwhere (N ints in N doubles) | scratch |
  //         lo hi | nbd (N ints in N doubles) | scalars [| (s, y) ring]
  static constexpr int O_Z = 0, O_LS = O_Z + 4 * N, O_X = O_LS + 20, O_G = O_X + N, O_B = O_G + N,
                       O_IW = O_B + N * N, O_SCR = O_IW + N, O_LO = O_SCR + 4 * N + 3 * N * N, O_HI = O_LO + N,
                       O_NBD = O_HI + N, O_SC = O_NBD + N, STATE = O_SC + 8, SIZE = STATE + 2 * LB_M * N;
  // scalars: doubles theta, sbgnrm, fold, gdold, stp, f; ints (in two doubles) iter, ifun, col+head, flags
  enum { F_CNSTND = 1, F_BOXED = 2, F_STARTED = 4, F_RUNNING = 8 };

  PWB_HD double* xs() const { return st + O_X; }        // current point: evaluate (f, g) here ...
  PWB_HD double* gs() const { return st + O_G; }        // ... store the gradient here and pass f to advance()
  PWB_HD const double* lo() const { return st + O_LO; }
  PWB_HD const double* hi() const { return st + O_HI; }
  PWB_HD int nbd(int i) const { return NBD >= 0 ? NBD : reinterpret_cast<const int*>(st + O_NBD)[i]; }
  PWB_HD double& theta() const { return st[O_SC]; }
  PWB_HD double& sbgnrm() const { return st[O_SC + 1]; }
  PWB_HD double& fold() const { return st[O_SC + 2]; }
  PWB_HD double& gdold() const { return st[O_SC + 3]; }
  PWB_HD double& stp() const { return st[O_SC + 4]; }
  PWB_HD double& fval() const { return st[O_SC + 5]; }  // objective at x (final value after termination)
  PWB_HD int* ints() const { return reinterpret_cast<int*>(st + O_SC + 6); }
  PWB_HD int& iter() const { return ints()[0]; }
  PWB_HD int& ifun() const { return ints()[1]; }
  PWB_HD int& colhead() const { return ints()[2]; }      // col | head << 8
  PWB_HD int& flags() const { return ints()[3]; }

  // x0 is projected into the box; the caller then evaluates (f, g) at xs()
  PWB_HD void init(double* storage, const double* x0, const double* lo_, const double* hi_, const int* nbd_in) {
    init(storage, storage + STATE, x0, lo_, hi_, nbd_in);
  }
  PWB_HD void init(double* state, double* history, const double* x0, const double* lo_, const double* hi_,
                   const int* nbd_in) {
    st = state;
    hs = history;
    double* const x = st + O_X;
    double* const lo = st + O_LO; double* const hi = st + O_HI;
    int* const nb = reinterpret_cast<int*>(st + O_NBD);
    int* const iwhere = reinterpret_cast<int*>(st + O_IW);
    for (int i = 0; i < N; i++) { x[i] = x0[i]; lo[i] = lo_[i]; hi[i] = hi_[i]; nb[i] = NBD >= 0 ? NBD : nbd_in[i]; }
    bool cnstnd = false, boxed = true;
    for (int i = 0; i < N; i++) {
      const int nbi = nbd(i);
      if (nbi > 0) {
        if (nbi <= 2 && x[i] <= lo[i]) { if (x[i] < lo[i]) x[i] = lo[i]; }
        else if (nbi >= 2 && x[i] >= hi[i]) { if (x[i] > hi[i]) x[i] = hi[i]; }
      }
    }
    for (int i = 0; i < N; i++) {
      const int nbi = nbd(i);
      if (nbi != 2) boxed = false;
      if (nbi == 0) iwhere[i] = -1;
      else { cnstnd = true; iwhere[i] = (nbi == 2 && hi[i] - lo[i] <= 0) ? 3 : 0; }
    }
    theta() = 1.0; sbgnrm() = 0; fold() = 0; gdold() = 0; stp() = 0; fval() = 0;
    iter() = 0; ifun() = 0; colhead() = 0;
    flags() = (cnstnd ? F_CNSTND : 0) | (boxed ? F_BOXED : 0);
  }

  // `f` and gs() hold the objective at xs().  Returns -1 if xs() now holds the next point to evaluate, otherwise
  // the termination status (LbStatus; xs(), fval(), iter() final).
  PWB_HD int advance(double f, double ftol, double pgtol, int maxls, int maxiter) {
    double* const x = st + O_X; double* const g = st + O_G; double* const B = st + O_B;
    int* const iwhere = reinterpret_cast<int*>(st + O_IW);
    double* const scratch = st + O_SCR;
    double* const z = st + O_Z; double* const d = z + N; double* const t = d + N; double* const r = t + N;
    const double* const lo = st + O_LO; const double* const hi = st + O_HI;
    LbLineSearch* const ls = reinterpret_cast<LbLineSearch*>(st + O_LS);
    const int* const nb = reinterpret_cast<const int*>(st + O_NBD);
    LbHistory<N> H;
    H.S = reinterpret_cast<double (*)[N]>(hs);
    H.Y = reinterpret_cast<double (*)[N]>(hs + LB_M * N);
    H.col = colhead() & 255; H.head = colhead() >> 8;
    int fl = flags();
    const bool cnstnd = fl & F_CNSTND, boxed = fl & F_BOXED;
    int it = iter(), nfun = ifun();
    double th = theta(), sb = sbgnrm(), step = stp();
    int status = -1;
    bool new_iter;
    if (!(fl & F_RUNNING)) {
      fl |= F_RUNNING;
      sb = lb_projgr<N, NBD>(x, g, lo, hi, nb);
      if (sb <= pgtol) status = LB_CONV_PGTOL;
      new_iter = true;
    } else new_iter = false;
    while (status < 0) {
      if (new_iter) {
        lb_build_B<N>(H, th, B);
        if (!cnstnd && H.col > 0) { for (int i = 0; i < N; i++) z[i] = x[i]; }
        else lb_cauchy<N, NBD>(x, lo, hi, nb, g, iwhere, B, th, sb, z, scratch);
        int nfree = 0;
        for (int i = 0; i < N; i++) if (iwhere[i] <= 0) nfree++;
        if (nfree != 0 && H.col != 0) {
          if (!lb_subsm<N, NBD>(x, g, lo, hi, nb, iwhere, B, z, scratch)) { H.col = 0; H.head = 0; th = 1.0; continue; }
        }
        // ---- line search set-up (lnsrlb)
        double dtd = 0;
        for (int i = 0; i < N; i++) { d[i] = z[i] - x[i]; dtd += d[i] * d[i]; }
        const double dnorm = lb_sqrt(dtd);
        double stpmx = 1e10;
        if (cnstnd) {
          if (it == 0) stpmx = 1.0;
          else {
            LB_ROLLED
            for (int i = 0; i < N; i++) {
              double a1 = d[i];
              const int nbi = nbd(i);
              if (nbi != 0) {
                if (a1 < 0 && nbi <= 2) {
                  double a2 = lo[i] - x[i];
                  if (a2 >= 0) stpmx = 0; else if (a1 * stpmx < a2) stpmx = lb_div(a2, a1);
                } else if (a1 > 0 && nbi >= 2) {
                  double a2 = hi[i] - x[i];
                  if (a2 <= 0) stpmx = 0; else if (a1 * stpmx > a2) stpmx = lb_div(a2, a1);
                }
              }
            }
          }
        }
        step = (it == 0 && !boxed) ? fmin(lb_div(1.0, dnorm), stpmx) : 1.0;
        for (int i = 0; i < N; i++) { t[i] = x[i]; r[i] = g[i]; }
        fold() = f;
        nfun = 0;
        ls->ftol = 1e-3; ls->gtol = 0.9; ls->xtol = 0.1; ls->stpmin = 0; ls->stpmax = stpmx;
        fl &= ~F_STARTED;
        new_iter = false;
      }
      // ---- one line-search step
      bool fail = false;
      double gd = 0;
      for (int i = 0; i < N; i++) gd += g[i] * d[i];
      if (nfun == 0) {
        gdold() = gd;
        if (gd >= 0) fail = true;
      }
      if (!fail) {
        int task;
        if (!(fl & F_STARTED)) { task = ls->start(step, f, gd); fl |= F_STARTED; }
        else task = ls->iterate(step, f, gd);
        if (!(task == LbLineSearch::CONV || task == LbLineSearch::WARN)) {
          nfun++;
          if (step == 1.0) { for (int i = 0; i < N; i++) x[i] = z[i]; }
          else { for (int i = 0; i < N; i++) x[i] = step * d[i] + t[i]; }
          if (nfun - 1 >= maxls) fail = true;
          else break;                                // evaluate (f, g) at x, then call advance() again
        }
      }
      if (fail) {
        for (int i = 0; i < N; i++) { x[i] = t[i]; g[i] = r[i]; }
        f = fold();
        if (H.col == 0) { status = LB_ABNORMAL; it++; break; }
        H.col = 0; H.head = 0; th = 1.0;
        new_iter = true;
        continue;
      }
      it++;
      sb = lb_projgr<N, NBD>(x, g, lo, hi, nb);
      if (sb <= pgtol) { status = LB_CONV_PGTOL; break; }
      const double fo = fold(), go = gdold();
      double ddum = fmax(fmax(fabs(fo), fabs(f)), 1.0);
      if ((fo - f) <= ftol * ddum) { status = LB_CONV_FTOL; break; }
      if (it >= maxiter) { status = LB_MAXITER; break; }
      double rr = 0, dr;
      for (int i = 0; i < N; i++) { r[i] = g[i] - r[i]; rr += r[i] * r[i]; }
      if (step == 1.0) { dr = gd - go; ddum = -go; }
      else {
        dr = (gd - go) * step;
        for (int i = 0; i < N; i++) d[i] *= step;
        ddum = -go * step;
      }
      new_iter = true;
      if (dr <= LB_EPSMCH * ddum) continue;       // skip the BFGS update
      int slot;
      if (H.col < LB_M) { slot = (H.head + H.col) % LB_M; H.col++; }
      else { slot = H.head; H.head = (H.head + 1) % LB_M; }
      for (int i = 0; i < N; i++) { H.S[slot][i] = d[i]; H.Y[slot][i] = r[i]; }
      th = lb_div(rr, dr);
    }
    theta() = th; sbgnrm() = sb; stp() = step; fval() = f;
    iter() = it; ifun() = nfun; colhead() = H.col | (H.head << 8); flags() = fl;
    return status;
  }
};

static_assert(LbSolver<3>::SIZE == LB_STORAGE(3) && LbSolver<1>::SIZE == LB_STORAGE(1) &&
              LbSolver<3>::STATE == LB_STATE(3), "LB_STORAGE / LB_STATE out of step with LbSolver");

// Obj must provide:  void fg(const double* x, double& f, double* g);
// (f and the bound-aware forward-difference gradient; on device it is warp-cooperative and every
// lane runs this routine redundantly with identical data, so control flow is warp-uniform.)
template <int N, class Obj, int NBD = -1>
PWB_HDN int lbfgsb_minimize(Obj& obj, double* x, const double* lo, const double* hi, const int* nbd_in,
                            double ftol, double pgtol, int maxls, int maxiter, double* f_out, int* nit_out,
                            double* hist_storage) {
  LbSolver<N, NBD> s;
  s.init(hist_storage, x, lo, hi, nbd_in);
  int status;
  double f;
  do {
    obj.fg(s.xs(), f, s.gs());
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
    status = s.advance(f, ftol, pgtol, maxls, maxiter);
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
  } while (status < 0);
  for (int i = 0; i < N; i++) x[i] = s.xs()[i];
  *f_out = s.fval();
  *nit_out = s.iter();
  return status;
}

}  // namespace pwb
