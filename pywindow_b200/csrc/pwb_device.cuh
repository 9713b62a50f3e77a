// Device helpers shared by the pore-geometry kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pwb.h"
#include "lbfgsb_small.cuh"
#include "neldermead2.cuh"

namespace pwb {

// Math-library heavy helpers (sincos / log10 / fmod / atan2 / acos expand to hundreds of instructions per call
// site).  Compiling them out of line (-DPWB_NOINLINE_MATH) shrinks the ray kernels by ~10 % but was measured
// SLOWER on B200 (k_rays 1.07 -> 1.27, k_cluster 0.48 -> 0.84 ms per 10 000 frames: the calls sit in per-ray
// loops), so they stay inline.
#ifdef PWB_NOINLINE_MATH
#define PWB_MATHFN static __device__ __noinline__
#else
#define PWB_MATHFN __device__ __forceinline__
#endif

constexpr int ITEM_D = 16;   // doubles of per-item scratch
constexpr int ITEM_I = 32;   // ints of per-item scratch
// item_d slots
constexpr int D_COM = 0, D_MAXD = 3, D_CENTRE = 4, D_RADIUS = 7, D_EPS = 8, D_ICOM = 9, D_SHIFT = 12;
// item_i slots
constexpr int I_STATUS = 0, I_P = 1, I_NSURV = 2, I_NCLEAR = 3, I_NWIN = 4, I_N = 5, I_CLS0 = 6;  // cls_start[17]
constexpr int I_QUEUE = 26;  // (first item of a launch only) work-queue head of k_opt_pore
constexpr int I_TILES_PREP = 28, I_TILES_RAYS = 29;   // CTAs of a multi-CTA item that have finished (self-resetting)
constexpr int I_DEFER = 24;   // cluster stage deferred to the large-shared-memory pass

constexpr double PWB_INF = 1e300;
constexpr double GOLDEN_ANGLE = 2.399963229728653;  // pi*(3-sqrt(5)), utilities.py:1412

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Python / numpy float floor division (fmod based), used for `norm // increment` utilities.py:1108
PWB_MATHFN double py_floordiv(double a, double b) {
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0 && ((b < 0) != (mod < 0))) div -= 1.0;
  if (div != 0.0) {
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
  }
  return 0.0;
}

// number of golden-spiral sampling points, utilities.py:1399-1409
PWB_MATHFN int n_sampling_points(double radius, double adjust) {
  // reference: 4*pi*r**2 -> (4*pi)*(r*r)
  double area = __dmul_rn(4.0 * 3.141592653589793, __dmul_rn(radius, radius));
  double v = __dmul_rn(__dmul_rn(log10(area), 250.0), adjust);
  return (int)v;
}

// point k of the P-point golden spiral scaled to radius R, utilities.py:1412-1423 (numpy linspace
// semantics: z_k = k*step + start with separate roundings, last element = stop exactly)
PWB_MATHFN void spiral_point(int k, int P, double R, double& px, double& py, double& pz) {
  double invP = 1.0 / (double)P;
  double start = 1.0 - invP, stop = invP - 1.0;
  double step = (stop - start) / (double)(P - 1);
  double z = (k == P - 1 && P > 1) ? stop : __dadd_rn(__dmul_rn((double)k, step), start);
  double rad = sqrt(__dadd_rn(1.0, -__dmul_rn(z, z)));
  double s, c;
  sincos(__dmul_rn(GOLDEN_ANGLE, (double)k), &s, &c);
  px = __dmul_rn(__dmul_rn(rad, c), R);
  py = __dmul_rn(__dmul_rn(rad, s), R);
  pz = __dmul_rn(z, R);
}

// ---- (value, index) reductions with first-occurrence tie-breaks --------------------------------
struct ValIdx { double v; int i; };
__device__ __forceinline__ ValIdx vi_min(ValIdx a, ValIdx b) {   // smaller v, then smaller i
  return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ValIdx vi_max(ValIdx a, ValIdx b) {   // larger v, then smaller i
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
// Warp-wide min / max of doubles with two 32-bit REDUX operations instead of five shuffle steps:
// doubles are mapped to order-isomorphic unsigned 64-bit keys (sign flip), the high words are
// reduced first, then the low words among the lanes that hold the winning high word.  Bit exact.
__device__ __forceinline__ unsigned long long dkey(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__device__ __forceinline__ double warp_min(double v) {
  const unsigned long long k = dkey(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return dunkey(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ double warp_max(double v) {
  const unsigned long long k = dkey(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return dunkey(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ ValIdx warp_vi_min(ValIdx a) {        // smaller v, then smaller i
  ValIdx r; r.v = warp_min(a.v);
  r.i = (int)__reduce_min_sync(0xffffffffu, a.v == r.v ? (unsigned)a.i : 0xffffffffu);
  return r;
}
__device__ __forceinline__ ValIdx warp_vi_max(ValIdx a) {        // larger v, then smaller i
  ValIdx r; r.v = warp_max(a.v);
  r.i = (int)__reduce_min_sync(0xffffffffu, a.v == r.v ? (unsigned)a.i : 0xffffffffu);
  return r;
}
__device__ __forceinline__ double warp_sum(double a) {
  for (int m = 16; m; m >>= 1) a += shfl_xor_d(a, m);
  return a;
}

// block-wide reductions (blockDim.x multiple of 32, <= 1024); `red` is >= 32 entries of scratch
__device__ __forceinline__ double block_sum(double a, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a);
  __syncthreads();
  if (lane == 0) red[w] = a;
  __syncthreads();
  double t = 0;
  for (int k = 0; k < nw; k++) t += red[k];
  return t;
}
__device__ __forceinline__ ValIdx block_vi(ValIdx a, bool want_max, double* redv, int* redi) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = want_max ? warp_vi_max(a) : warp_vi_min(a);
  __syncthreads();
  if (lane == 0) { redv[w] = a.v; redi[w] = a.i; }
  __syncthreads();
  ValIdx t; t.v = redv[0]; t.i = redi[0];
  for (int k = 1; k < nw; k++) {
    ValIdx b; b.v = redv[k]; b.i = redi[k];
    t = want_max ? vi_max(t, b) : vi_min(t, b);
  }
  return t;
}

// ---- type-bucketed atoms in shared memory -------------------------------------------------------
// Atoms of one item sorted (stably) by type so that min_i(|x_i-q| - r_i) = min_c(sqrt(min_{i in c}
// |x_i-q|^2) - r_c): the inner loop has no sqrt and the result is bit-identical because sqrt is
// monotone.
struct Bucketed {
  const double4* p;                                    // [n] shifted coordinates x, y, z and w = |x|^2
  const int* cls;                                      // [ncls+1] class starts
  const double* cr;                                    // [ncls] class radius
  int ncls;
};

// min_i(|x_i - q| - r_i) for K points at once.  |x-q|^2 = |x|^2 - 2 x.q + |q|^2: three FMAs and one
// min per atom and point, |q|^2 and the sqrt once per class.
template <int K>
__device__ __forceinline__ void min_clearance_k(const Bucketed& A, const double* qx, const double* qy,
                                                const double* qz, double* out) {
  double ax[K], ay[K], az[K], qq[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    qq[k] = qx[k] * qx[k] + qy[k] * qy[k] + qz[k] * qz[k];
    ax[k] = -2.0 * qx[k]; ay[k] = -2.0 * qy[k]; az[k] = -2.0 * qz[k];
    out[k] = PWB_INF;
  }
  for (int c = 0; c < A.ncls; c++) {
    const int b = A.cls[c], e = A.cls[c + 1];
    if (b == e) continue;
    double m[K];
#pragma unroll
    for (int k = 0; k < K; k++) m[k] = PWB_INF;
    for (int i = b; i < e; i++) {
      const double4 q = A.p[i];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = fmin(m[k], fma(ax[k], q.x, fma(ay[k], q.y, fma(az[k], q.z, q.w))));
    }
    const double rc = A.cr[c];
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = fmin(out[k], sqrt(fmax(m[k] + qq[k], 0.0)) - rc);
  }
}

__device__ __forceinline__ double min_clearance(const Bucketed& A, double qx, double qy, double qz) {
  double o[1];
  min_clearance_k<1>(A, &qx, &qy, &qz, o);
  return o[0];
}

// ---- warp-cooperative pore objective (lanes stride over atoms) ---------------------------------
struct WarpAtoms {
  const double4* p;      // [n] x, y, z (w belongs to the owner: radius in k_opt_pore, |x|^2 in k_windows)
  const double* r;       // vdW radius of atom i at r[i * rs]
  int rs;
  int n, lane;
  const int* cls;      // [ncls+1] class starts of the type-bucketed atoms
  int ncls;
};

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }

// Warp-wide min of NON-NEGATIVE doubles (squared distances, +inf): their bit patterns are already
// ordered as unsigned integers, so two REDUX operations suffice.
__device__ __forceinline__ double warp_min_nonneg(double v) {
  const unsigned long long k = (unsigned long long)__double_as_longlong(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}

// Same result as warp_min_clear, bit for bit, with ONE square root per warp instead of one per
// (lane, class, point): sqrt is monotone, so the squared distances are reduced over the warp first;
// the reduced value of (class c, point k) is parked in lane c*K + k, all lanes take the root of their
// own value at once, and a stride-K butterfly finishes the min over classes.
// `first(c)`, `last(c)` give the class's range, `atom(t)` maps a range position to an atom index.
template <int K, class First, class Last, class Atom>
__device__ __forceinline__ void warp_min_clear_rs(const WarpAtoms& A, First first, Last last, Atom atom,
                                                  const double* px, const double* py, const double* pz, double* out) {
  constexpr int CPG = 32 / K;                       // classes per group of 32 lanes
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = PWB_INF;
  for (int c0 = 0; c0 < A.ncls; c0 += CPG) {
    double mine = PWB_INF;
    const int cend = min(A.ncls, c0 + CPG);
    for (int c = c0; c < cend; c++) {
      const int cb = first(c), ce = last(c);
      if (cb == ce) continue;
      double m[K];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = PWB_INF;
      for (int t = cb + A.lane; t < ce; t += 32) {
        const double4 q = A.p[atom(t)];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
          m[k] = dmin(m[k], dx * dx + dy * dy + dz * dz);
        }
      }
      const int slot = (c - c0) * K;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double red = warp_min_nonneg(m[k]);
        if (A.lane == slot + k) mine = red;
      }
    }
    const int myc = c0 + A.lane / K;
    double v = PWB_INF;
    if (mine < PWB_INF) v = sqrt(mine) - A.r[A.cls[myc] * A.rs];
    if (K == 1) {
      out[0] = dmin(out[0], warp_min(v));
    } else {
#pragma unroll
      for (int s = K; s < 32; s <<= 1) v = dmin(v, shfl_xor_d(v, s));
#pragma unroll
      for (int k = 0; k < K; k++) out[k] = dmin(out[k], __shfl_sync(0xffffffffu, v, k));
    }
  }
}

// min_i(|x_i - p_k| - r_i) for K points; lanes stride over the atoms of one class at a time so the
// sqrt is taken once per (lane, class, point) instead of once per (atom, point).
template <int K>
__device__ __forceinline__ void warp_min_clear(const WarpAtoms& A, const double* px, const double* py,
                                               const double* pz, double* out) {
  if constexpr (K >= 4) {   // measured: pays for the 4-point finite-difference stencil, not for K = 1, 2
    warp_min_clear_rs<K>(A, [&](int c) { return A.cls[c]; }, [&](int c) { return A.cls[c + 1]; },
                         [](int t) { return t; }, px, py, pz, out);
  } else {
  double v[K];
#pragma unroll
  for (int k = 0; k < K; k++) v[k] = PWB_INF;
  for (int c = 0; c < A.ncls; c++) {
    const int cb = A.cls[c], ce = A.cls[c + 1];
    if (cb == ce) continue;
    double m[K];
#pragma unroll
    for (int k = 0; k < K; k++) m[k] = PWB_INF;
    for (int i = cb + A.lane; i < ce; i += 32) {
      const double4 q = A.p[i];
      const double ax = q.x, ay = q.y, az = q.z;
#pragma unroll
      for (int k = 0; k < K; k++) {
        double dx = ax - px[k], dy = ay - py[k], dz = az - pz[k];
        m[k] = dmin(m[k], dx * dx + dy * dy + dz * dz);
      }
    }
    const double rc = A.r[cb * A.rs];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = dmin(v[k], sqrt(m[k]) - rc);
  }
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = warp_min(v[k]);
  }
}

// ---- lane groups: G lanes (a power of two, aligned) of a warp work on one problem ---------------------------
// k_opt_pore runs 32/G independent pore-centre searches per warp, one per group of G lanes.  Every group-wide
// collective names the group's own lanes in its member mask, so groups may sit in different branches.
template <int G> __device__ __forceinline__ unsigned group_mask(int lane) {
  return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
}
template <int G> __device__ __forceinline__ double group_min_nonneg(unsigned mask, double v) {   // v >= 0 or +inf
  const unsigned long long k = (unsigned long long)__double_as_longlong(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(mask, hi);
  const unsigned ml = __reduce_min_sync(mask, hi == mh ? lo : 0xffffffffu);
  return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
template <int G> __device__ __forceinline__ ValIdx group_vi_min(unsigned mask, ValIdx a) {   // smaller v, then smaller i
  const unsigned long long k = dkey(a.v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(mask, hi);
  const unsigned ml = __reduce_min_sync(mask, hi == mh ? lo : 0xffffffffu);
  ValIdx r; r.v = dunkey(((unsigned long long)mh << 32) | ml);
  r.i = (int)__reduce_min_sync(mask, a.v == r.v ? (unsigned)a.i : 0xffffffffu);
  return r;
}

// |d|^2 with a FIXED operation order (one product, two FMAs): left to the compiler, the contraction of
// dx*dx + dy*dy + dz*dz depends on the surrounding code and two instantiations of the same source can differ in the
// last bit -- enough to send the chaotic pore-centre search elsewhere.
__device__ __forceinline__ double sq3(double dx, double dy, double dz) {
  return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

struct GroupAtoms {
  const double4* p;      // [n] x, y, z, vdW radius of the type-bucketed atoms (shared or global memory)
  int n, gl;             // atoms, lane index inside the group
  const int* cls;        // [ncls+1] class starts
  int ncls;
  unsigned mask;         // member mask of the group
};

// min_i(|x_i - p_k| - r_i) for K points by the G lanes of one group; same arithmetic and the same reduce-before-
// sqrt scheme as warp_min_clear_rs (bit-identical results for every G): the squared distances of (class c, point
// k) are reduced over the group, parked in group lane (c - c0)*K + k, rooted once, and a stride-K butterfly
// finishes the min over the G/K classes of a pass.
// G < K (one or two lanes per problem): every lane walks its share of each class and the class minima are combined
// with plain shuffles -- same values, same reduce-before-sqrt order.
template <int G, int K>
__device__ __forceinline__ void group_min_clear_narrow(const GroupAtoms& A, const double* px, const double* py,
                                                       const double* pz, double* out) {
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = PWB_INF;
  for (int c = 0; c < A.ncls; c++) {
    const int cb = A.cls[c], ce = A.cls[c + 1];
    if (cb == ce) continue;
    double m[K];
#pragma unroll
    for (int k = 0; k < K; k++) m[k] = PWB_INF;
    // four atoms per step, loaded before any of them is used: with one or two lanes per problem the tile comes from
    // L1 / L2 and the loop is bound by the latency of these loads
    int t = cb + A.gl;
    for (; t + 3 * G < ce; t += 4 * G) {
      const double4 q[4] = {A.p[t], A.p[t + G], A.p[t + 2 * G], A.p[t + 3 * G]};
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int k = 0; k < K; k++) {
          double dx = q[u].x - px[k], dy = q[u].y - py[k], dz = q[u].z - pz[k];
          m[k] = dmin(m[k], sq3(dx, dy, dz));
        }
    }
    for (; t < ce; t += G) {
      const double4 q = A.p[t];
#pragma unroll
      for (int k = 0; k < K; k++) {
        double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
        m[k] = dmin(m[k], sq3(dx, dy, dz));
      }
    }
    const double rc = A.p[cb].w;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
      for (int s = 1; s < G; s <<= 1) m[k] = dmin(m[k], __shfl_xor_sync(A.mask, m[k], s, G));
      out[k] = dmin(out[k], sqrt(m[k]) - rc);
    }
  }
}

template <int G, int K>
__device__ __forceinline__ void group_min_clear(const GroupAtoms& A, const double* px, const double* py,
                                                const double* pz, double* out) {
  if constexpr (G < K) { group_min_clear_narrow<G, K>(A, px, py, pz, out); return; }
  static_assert(G < K || (G % K) == 0, "a pass must hold at least one class");
  constexpr int CPG = G / K;
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = PWB_INF;
  for (int c0 = 0; c0 < A.ncls; c0 += CPG) {
    double mine = PWB_INF;
    const int cend = min(A.ncls, c0 + CPG);
    for (int c = c0; c < cend; c++) {
      const int cb = A.cls[c], ce = A.cls[c + 1];
      if (cb == ce) continue;
      double m[K];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = PWB_INF;
#ifdef OBJ_ROLLED
#pragma unroll 1
#endif
      for (int t = cb + A.gl; t < ce; t += G) {
        const double4 q = A.p[t];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
          m[k] = dmin(m[k], sq3(dx, dy, dz));
        }
      }
      const int slot = (c - c0) * K;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double red = group_min_nonneg<G>(A.mask, m[k]);
        if (A.gl == slot + k) mine = red;
      }
    }
    double v = PWB_INF;
    if (mine < PWB_INF) v = sqrt(mine) - A.p[A.cls[c0 + A.gl / K]].w;
#pragma unroll
    for (int s = K; s < G; s <<= 1) v = dmin(v, __shfl_xor_sync(A.mask, v, s, G));
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = dmin(out[k], __shfl_sync(A.mask, v, k, G));
  }
}

// The same with EXACT pruning against a reference evaluation.  a(p) = min_i(|x_i - p| - r_i) is 1-Lipschitz, so for a
// point within delta of a reference point p_ref the minimiser j obeys |x_j - p_ref| <= a(p_ref) + r_j + 2 delta;
// every other atom is provably not the minimum and is skipped after ONE comparison of its stored reference distance
// (squared, float, rounded down) with the class threshold (rounded up).  The surviving atoms are evaluated with the
// unchanged arithmetic, so the result is bit-identical to group_min_clear.  `store`: evaluate all atoms (thr = inf)
// and record their squared distances to point 0 as the new reference.  In a line search the trial points crowd
// around the iterate: ~90 % of the atoms drop out.
template <int G, int K>
__device__ __forceinline__ void group_min_clear_pruned(const GroupAtoms& A, const double* px, const double* py,
                                                       const double* pz, double* out, float* d2ref, double thr_base,
                                                       bool store) {
  static_assert(G >= K && (G % K) == 0, "a pass must hold at least one class");
  constexpr int CPG = G / K;
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = PWB_INF;
  for (int c0 = 0; c0 < A.ncls; c0 += CPG) {
    double mine = PWB_INF;
    const int cend = min(A.ncls, c0 + CPG);
    for (int c = c0; c < cend; c++) {
      const int cb = A.cls[c], ce = A.cls[c + 1];
      if (cb == ce) continue;
      const double rc = A.p[cb].w;
      const double th = thr_base + rc;
      const float thr2 = store ? 3.0e38f : __double2float_ru(th * th);
      double m[K];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = PWB_INF;
      for (int t = cb + A.gl; t < ce; t += G) {
        if (!store && d2ref[t] > thr2) continue;
        const double4 q = A.p[t];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
          const double d2 = sq3(dx, dy, dz);
          if (k == 0 && store) d2ref[t] = __double2float_rd(d2);
          m[k] = dmin(m[k], d2);
        }
      }
      const int slot = (c - c0) * K;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double red = group_min_nonneg<G>(A.mask, m[k]);
        if (A.gl == slot + k) mine = red;
      }
    }
    double v = PWB_INF;
    if (mine < PWB_INF) v = sqrt(mine) - A.p[A.cls[c0 + A.gl / K]].w;
#pragma unroll
    for (int s = K; s < G; s <<= 1) v = dmin(v, __shfl_xor_sync(A.mask, v, s, G));
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = dmin(out[k], __shfl_sync(A.mask, v, k, G));
  }
}

// ---- survivor lists: the objective of a search that stays near a reference point ------------------------------
// a(p) = min_i(|x_i - p| - r_i) is 1-Lipschitz: for every p within R of a reference point p_ref with clearance
// a_ref, an atom with |x_i - p_ref| - r_i > a_ref + 2R has |x_i - p| - r_i > a_ref + R >= a(p) and cannot attain the
// minimum.  group_build_list keeps the indices of the other atoms (ascending); group_min_list evaluates a stencil of
// K points over them.  Each lane takes sqrt(min d2) - r over its own RUNS of equal radius instead of one reduction per
// class over the group: the same doubles, because sqrt is monotone -- so the result is bit-identical to
// group_min_clear for every point of the ball.
template <int G>
__device__ __forceinline__ int group_build_list(const GroupAtoms& A, double cx, double cy, double cz, double a_ref,
                                                double two_r, unsigned short* lst, unsigned short* lcls, int cap) {
  const int gbase_bit = (threadIdx.x & 31) & ~(G - 1);
  int n = 0;
  for (int c = 0; c < A.ncls; c++) {
    if (A.gl == 0) lcls[c] = (unsigned short)min(n, cap);      // where class c starts in the list
    const int cb = A.cls[c], ce = A.cls[c + 1];
    if (cb == ce) continue;
    const double th = a_ref + A.p[cb].w + two_r + 1e-9;          // margin >> the rounding of d2 and of the sum
    const double th2 = th >= 0.0 ? th * th * (1.0 + 1e-12) : -1.0;
    for (int t0 = cb; t0 < ce; t0 += G) {
      const int t = t0 + A.gl;
      bool keep = false;
      if (t < ce) {
        const double4 q = A.p[t];
        const double dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
        keep = sq3(dx, dy, dz) <= th2;
      }
      const unsigned bits = (__ballot_sync(A.mask, keep) >> gbase_bit) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u));
      if (keep) {
        const int pos = n + __popc(bits & ((1u << A.gl) - 1u));
        if (pos < cap) lst[pos] = (unsigned short)t;
      }
      n += __popc(bits);
    }
  }
  if (A.gl == 0) lcls[A.ncls] = (unsigned short)min(n, cap);
  return n <= cap ? n : -1;
}

// group_min_clear over the list: the list is ascending, i.e. sorted by class like the atoms, and lcls[c] is where
// class c starts in it -- same class-wise reduce-before-sqrt scheme, one sqrt per (class, point).
template <int G, int K>
__device__ __forceinline__ void group_min_clear_list(const GroupAtoms& A, const unsigned short* lst,
                                                     const unsigned short* lcls, const double* px, const double* py,
                                                     const double* pz, double* out) {
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = PWB_INF;
  if constexpr (G < K) {
    for (int c = 0; c < A.ncls; c++) {
      const int lb = lcls[c], le = lcls[c + 1];
      if (lb == le) continue;
      double m[K];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = PWB_INF;
      for (int t = lb + A.gl; t < le; t += G) {
        const double4 q = A.p[lst[t]];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
          m[k] = dmin(m[k], sq3(dx, dy, dz));
        }
      }
      const double rc = A.p[A.cls[c]].w;
#pragma unroll
      for (int k = 0; k < K; k++) {
#pragma unroll
        for (int s = 1; s < G; s <<= 1) m[k] = dmin(m[k], __shfl_xor_sync(A.mask, m[k], s, G));
        out[k] = dmin(out[k], sqrt(m[k]) - rc);
      }
    }
  } else {
    constexpr int CPG = G / K;
    for (int c0 = 0; c0 < A.ncls; c0 += CPG) {
      double mine = PWB_INF;
      const int cend = min(A.ncls, c0 + CPG);
      for (int c = c0; c < cend; c++) {
        const int lb = lcls[c], le = lcls[c + 1];
        if (lb == le) continue;
        double m[K];
#pragma unroll
        for (int k = 0; k < K; k++) m[k] = PWB_INF;
        for (int t = lb + A.gl; t < le; t += G) {
          const double4 q = A.p[lst[t]];
#pragma unroll
          for (int k = 0; k < K; k++) {
            double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
            m[k] = dmin(m[k], sq3(dx, dy, dz));
          }
        }
        const int slot = (c - c0) * K;
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double red = group_min_nonneg<G>(A.mask, m[k]);
          if (A.gl == slot + k) mine = red;
        }
      }
      double v = PWB_INF;
      if (mine < PWB_INF) v = sqrt(mine) - A.p[A.cls[min(A.ncls - 1, c0 + A.gl / K)]].w;
#pragma unroll
      for (int s = K; s < G; s <<= 1) v = dmin(v, __shfl_xor_sync(A.mask, v, s, G));
#pragma unroll
      for (int k = 0; k < K; k++) out[k] = dmin(out[k], __shfl_sync(A.mask, v, k, G));
    }
  }
}

template <int G, int K>
__device__ __forceinline__ void group_min_list(const GroupAtoms& A, const unsigned short* lst, int nl, const double* px,
                                               const double* py, const double* pz, double* out) {
  double m[K];
#pragma unroll
  for (int k = 0; k < K; k++) m[k] = PWB_INF;
  // the list is sorted by class, so the atoms of one lane come in runs of equal radius: min d2 over a run, one sqrt
  // per run and point (sqrt is monotone: the same doubles as one sqrt per atom)
  double run[K], rr = -1.0;
#pragma unroll
  for (int k = 0; k < K; k++) run[k] = PWB_INF;
  for (int e = A.gl; e < nl; e += G) {
    const double4 q = A.p[lst[e]];
    if (q.w != rr) {
      if (rr >= 0.0) {
#pragma unroll
        for (int k = 0; k < K; k++) { m[k] = dmin(m[k], sqrt(run[k]) - rr); run[k] = PWB_INF; }
      }
      rr = q.w;
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
      run[k] = dmin(run[k], sq3(dx, dy, dz));
    }
  }
  if (rr >= 0.0) {
#pragma unroll
    for (int k = 0; k < K; k++) m[k] = dmin(m[k], sqrt(run[k]) - rr);
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int s = 1; s < G; s <<= 1) m[k] = dmin(m[k], __shfl_xor_sync(A.mask, m[k], s, G));
    out[k] = m[k];
  }
}

// pore_diameter with first-occurrence argmin over ORIGINAL atom indices (utilities.py:375-388), by one group
template <int G>
__device__ __forceinline__ ValIdx group_pore_argmin(const GroupAtoms& A, const int* orig, double cx, double cy, double cz) {
  ValIdx b; b.v = PWB_INF; b.i = 0x7fffffff;
#pragma unroll 1
  for (int i = A.gl; i < A.n; i += G) {
    const double4 q = A.p[i];
    double dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
    ValIdx c; c.v = sqrt(dx * dx + dy * dy + dz * dz) - q.w; c.i = orig[i];
    b = vi_min(b, c);
  }
  return group_vi_min<G>(A.mask, b);
}

// ---- direction grid: which atoms can a ray from the centroid hit at all? ---------------------------
// The pre-test of find_windows and the average-diameter sampling both ask, for ~1e3 rays from the
// centroid o, which of the N vdW spheres the ray's LINE intersects.  Sphere i subtends a cap of angular
// radius asin(r_i/|L_i|) around +-L_i/|L_i| (L_i = x_i - o); caps are rasterised (conservatively: bounding
// box in z = cos(theta) and phi, plus a 1e-7 rad margin) into a 16 x 32 grid of direction bins, so a ray only
// tests the handful of atoms listed in its bin.  The tests themselves are unchanged (exact, FP64).
constexpr int DG_NZ = 16, DG_NPHI = 32, DG_BINS = DG_NZ * DG_NPHI, DG_CAP = 24;
constexpr double DG_PI = 3.141592653589793;

struct DirGrid { unsigned short* list; int* cnt; int* overflow; };   // list[DG_BINS][DG_CAP], cnt[DG_BINS]

__device__ __forceinline__ int dg_bin(double nx, double ny, double nz) {
  int iz = (int)floor((nz + 1.0) * (DG_NZ * 0.5));
  iz = iz < 0 ? 0 : (iz >= DG_NZ ? DG_NZ - 1 : iz);
  // phi bin = floor((atan2(ny, nx) + pi) / (pi/16)) without the atan2: sixteenths of pi inside the first quadrant
  // by three comparisons against tan(k pi/16), then the quadrant from the signs.  A direction within rounding
  // of a bin boundary may land on either side; the caps are rasterised with a 1e-7 rad margin, so both bins
  // list every atom it can meet.
  static_assert(DG_NPHI == 32, "phi binning below assumes 32 bins");
  const double a = fabs(nx), c = fabs(ny);
  constexpr double T1 = 0.19891236737965800, T2 = 0.41421356237309503, T3 = 0.66817863791929890;
  const int q8 = c <= a ? (int)(c > a * T1) + (int)(c > a * T2) + (int)(c > a * T3)
                        : 7 - ((int)(a > c * T1) + (int)(a > c * T2) + (int)(a > c * T3));
  const int ip = ny >= 0.0 ? (nx >= 0.0 ? 16 + q8 : 31 - q8) : (nx >= 0.0 ? 15 - q8 : q8);
  return iz * DG_NPHI + ip;
}

__device__ __forceinline__ void dg_push(const DirGrid& G, int bin, int a) {
  const int pos = atomicAdd(&G.cnt[bin], 1);
  if (pos < DG_CAP) G.list[bin * DG_CAP + pos] = (unsigned short)a; else *G.overflow = 1;
}

// Cap of angular radius alpha = asin(sa) around the unit vector u, rasterised conservatively WITHOUT
// inverse trigonometry: with ct = u_z, st = sqrt(1 - u_z^2), ca = sqrt(1 - sa^2) the cap spans
// z in [ct ca - st sa, ct ca + st sa] (a pole inside the cap extends that to +-1 and to every phi), and its
// half-width in phi is asin(sa / st) <= x + (pi/2 - 1) x^3 (asin is convex on [0, 1]).  `sa` already carries
// the caller's safety margin; supersets only cost extra exact tests.
PWB_MATHFN void dg_insert_cap(const DirGrid& G, int a, double ux, double uy, double uz, double sa) {
  const double ct = fmin(1.0, fmax(-1.0, uz));
  const double st = sqrt(fmax(0.0, 1.0 - ct * ct)), ca = sqrt(fmax(0.0, 1.0 - sa * sa));
  const bool north = ct >= ca, south = -ct >= ca;          // theta_c - alpha <= 0, theta_c + alpha >= pi
  const double zhi = north ? 1.0 : fmin(1.0, ct * ca + st * sa + 1e-9);
  const double zlo = south ? -1.0 : fmax(-1.0, ct * ca - st * sa - 1e-9);
  int iz0 = (int)floor((zlo + 1.0) * (DG_NZ * 0.5) - 1e-9), iz1 = (int)floor((zhi + 1.0) * (DG_NZ * 0.5) + 1e-9);
  iz0 = iz0 < 0 ? 0 : iz0; iz1 = iz1 >= DG_NZ ? DG_NZ - 1 : iz1;
  bool full = north || south;
  int ip0 = 0, ip1 = DG_NPHI - 1;
  if (!full) {
    const double sr = sa / st;
    if (!(sr < 1.0)) full = true;
    else {
      const double dphi = sr + (0.5 * DG_PI - 1.0) * sr * sr * sr + 1e-7, pc = atan2(uy, ux) + DG_PI;
      ip0 = (int)floor((pc - dphi) * (DG_NPHI / (2.0 * DG_PI)));
      ip1 = (int)floor((pc + dphi) * (DG_NPHI / (2.0 * DG_PI)));
      if (ip1 - ip0 + 1 >= DG_NPHI) full = true;
    }
  }
  if (full) { ip0 = 0; ip1 = DG_NPHI - 1; }
  for (int iz = iz0; iz <= iz1; iz++)
    for (int ip = ip0; ip <= ip1; ip++) dg_push(G, iz * DG_NPHI + ((ip % DG_NPHI) + DG_NPHI) % DG_NPHI, a);
}

// block-wide build from centroid-relative positions L and radii r (one thread per atom); `onorm` = |o|,
// the distance of the centroid from the origin the forward test (x.n > 0) refers to
__device__ __forceinline__ void dg_build(const DirGrid& G, const double* Lx, const double* Ly, const double* Lz,
                                         const double* r, int n, double onorm) {
  for (int c = threadIdx.x; c < DG_BINS; c += blockDim.x) G.cnt[c] = 0;
  if (threadIdx.x == 0) *G.overflow = (n > 65535) ? 1 : 0;
  __syncthreads();
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const double x = Lx[a], y = Ly[a], z = Lz[a], ra = r[a];
    const double L = sqrt(x * x + y * y + z * z);
    if (L <= ra * (1.0 + 1e-9) + 1e-12) {               // the centroid lies inside this sphere: every ray hits it
      for (int c = 0; c < DG_BINS; c++) dg_push(G, c, a);
      continue;
    }
    const double sa = fmin(1.0, ra / L + 2e-7);          // sin of the cap radius, with a margin of > 1e-7 rad
    const double il = 1.0 / L;
    dg_insert_cap(G, a, x * il, y * il, z * il, sa);
    // the line also meets the sphere on the far side of o; that only counts if x.n > 0 can still hold,
    // i.e. for atoms closer to the centroid than the centroid is to the origin
    if (L - ra <= onorm + 1e-9) dg_insert_cap(G, a, -x * il, -y * il, -z * il, sa);
  }
  __syncthreads();
}

// ---- near lists: exact pruning for repeated evaluations around one point -------------------------
// a(p) = min_i(|x_i - p| - r_i) is 1-Lipschitz, so for |p - c| <= rho:  a(p) <= a(c) + rho, while atom i
// contributes at least |x_i - c| - rho - r_i.  Atoms with |x_i - c| - r_i > a(c) + 2 rho can therefore
// never attain the minimum inside the ball and are dropped; the minimum over the kept atoms is
// bit-identical to the minimum over all atoms.  The list holds atom indices, still grouped by class.
struct NearList {
  unsigned short* idx;   // [n] kept atom indices (shared memory)
  int* cs;               // [ncls+1] class starts inside idx
  double cx, cy, cz, rho2;
};

__device__ __forceinline__ void warp_build_near(const WarpAtoms& A, NearList& N, double cx, double cy, double cz,
                                                double rho) {
  double m[1];
  warp_min_clear<1>(A, &cx, &cy, &cz, m);
  const double thr = m[0] + 2.0 * rho + 1e-9;
  int pos = 0;
  for (int c = 0; c < A.ncls; c++) {
    if (A.lane == 0) N.cs[c] = pos;
    const int cb = A.cls[c], ce = A.cls[c + 1];
    for (int base = cb; base < ce; base += 32) {
      const int i = base + A.lane;
      bool keep = false;
      if (i < ce) {
        const double4 q = A.p[i];
        const double dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
        keep = sqrt(dx * dx + dy * dy + dz * dz) - A.r[i * A.rs] <= thr;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) N.idx[pos + __popc(mask & ((1u << A.lane) - 1u))] = (unsigned short)i;
      pos += __popc(mask);
    }
  }
  if (A.lane == 0) N.cs[A.ncls] = pos;
  N.cx = cx; N.cy = cy; N.cz = cz;
  const double re = rho - 1e-6;
  N.rho2 = re * re;
  __syncwarp();
}

__device__ __forceinline__ bool near_valid(const NearList& N, double x, double y, double z) {
  const double dx = x - N.cx, dy = y - N.cy, dz = z - N.cz;
  return dx * dx + dy * dy + dz * dz <= N.rho2;
}

// lanes stride over the kept atoms (optimiser objectives).  The list is short (the atoms lining one
// window, typically 20-50), so one flat pass with a square root per atom beats the class-wise loops of
// warp_min_clear: with one or two atoms per lane those were mostly loop overhead.  Same values, and min
// does not care about the order.
template <int K>
__device__ __forceinline__ void warp_min_clear_near(const WarpAtoms& A, const NearList& N, const double* px,
                                                    const double* py, const double* pz, double* out) {
  double v[K];
#pragma unroll
  for (int k = 0; k < K; k++) v[k] = PWB_INF;
  const int nn = N.cs[A.ncls];
  for (int t = A.lane; t < nn; t += 32) {
    const int i = N.idx[t];
    const double4 q = A.p[i];
    const double ri = A.r[i * A.rs];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double dx = q.x - px[k], dy = q.y - py[k], dz = q.z - pz[k];
      v[k] = dmin(v[k], sqrt(dx * dx + dy * dy + dz * dz) - ri);
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = warp_min(v[k]);
}

// every lane evaluates its own K points over the kept atoms (grid search)
template <int K>
__device__ __forceinline__ void min_clearance_k_near(const Bucketed& A, const NearList& N, const double* qx,
                                                     const double* qy, const double* qz, double* out) {
  double ax[K], ay[K], az[K], qq[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    qq[k] = qx[k] * qx[k] + qy[k] * qy[k] + qz[k] * qz[k];
    ax[k] = -2.0 * qx[k]; ay[k] = -2.0 * qy[k]; az[k] = -2.0 * qz[k];
    out[k] = PWB_INF;
  }
  for (int c = 0; c < A.ncls; c++) {
    const int b = N.cs[c], e = N.cs[c + 1];
    if (b == e) continue;
    double m[K];
#pragma unroll
    for (int k = 0; k < K; k++) m[k] = PWB_INF;
    for (int t = b; t < e; t++) {
      const int i = N.idx[t];
      const double4 q = A.p[i];
#pragma unroll
      for (int k = 0; k < K; k++) m[k] = dmin(m[k], fma(ax[k], q.x, fma(ay[k], q.y, fma(az[k], q.z, q.w))));
    }
    const double rc = A.cr[c];
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = dmin(out[k], sqrt(fmax(m[k] + qq[k], 0.0)) - rc);
  }
}

// pore_diameter with first-occurrence argmin over ORIGINAL atom indices (utilities.py:375-388)
__device__ __forceinline__ ValIdx warp_pore_argmin(const WarpAtoms& A, const int* orig, double cx, double cy, double cz) {
  ValIdx b; b.v = PWB_INF; b.i = 0x7fffffff;
  for (int i = A.lane; i < A.n; i += 32) {
    const double4 q = A.p[i];
    double dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
    ValIdx c; c.v = sqrt(dx * dx + dy * dy + dz * dz) - A.r[i * A.rs]; c.i = orig[i];
    b = vi_min(b, c);
  }
  return warp_vi_min(b);
}

}  // namespace pwb
