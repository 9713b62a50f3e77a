// pwb_kernels.cu -- sm_100a kernels for pywindow's per-frame pore-geometry path.
//
// One batch = F independent molecule-frames.  Stage kernels (all float64 in this revision):
//   k_prepare   CTA/item  : type-bucket atoms, centre of mass, pore diameter, max dimension,
//                           average diameter                    (utilities.py:127,355,375,1586)
//   k_opt_pore  warp/item : pore-centre search (L-BFGS-B)       (utilities.py:400)
//   k_rays      CTA/item  : golden-spiral ray pre-test + stepped path clearance
//                                                               (utilities.py:1132,1100,1398-1470)
//   k_knn_fill  CTA/P     : mean 10-NN distance of the unit P-point spiral (eps table, :1427-1434)
//   k_cluster   CTA/item  : DBSCAN of clear rays + arg-max ray per cluster (:1478, :1221)
//   k_windows   warp/(item,window): window_analysis            (utilities.py:1191)
#include <stdio.h>

#include "pwb_device.cuh"

namespace pwb {

struct Ws {   // resolved workspace pointers
  double* atoms;   // [total][4]
  int* atom_idx;   // [total]
  double* item_d;  // [F][ITEM_D]
  int* item_i;     // [F][ITEM_I]
  int* clear_k;    // [F][ray_cap]
  int* clear_pos;
  double* clear_a2;
  int* labels;
  int* win_k;      // [F][win_cap]
  double* big_pts; // [F][ray_cap][4] unit-sphere ray points (only if ray_cap > KNN_SMEM_CAP)
  int* big_i;      // [F][2][ray_cap] union-find / flag scratch of the large-ray-count path
  // Molecules whose atom tiles do not fit the 227 KB of shared memory (> ~2 900 atoms): the per-CTA / per-warp
  // atom arrays of k_prepare, k_rays and k_windows live here instead (read through L1 / L2); nullptr otherwise.
  double* atom_scratch;
  double* opt_hist;      // [F][LB_HISTORY(3)] (s, y) rings of k_opt_pore when they are not kept in shared memory
  long long scratch_item;   // doubles of scratch per item
  int opt_no_list;          // test switch (pwb_ctx_set_option): k_opt_pore evaluates every atom at every point
};

// shared-memory need of k_rays, the largest of the three, decides: host and layout code agree through these
constexpr size_t PWB_SMEM_LIMIT = 232448;              // 227 KB opt-in maximum per CTA on sm_100
constexpr int PWB_MAX_SMS = 160;                       // bound on the SM count for sizing (B200: 148)

constexpr int KNN_SMEM_CAP = 4096;   // ray counts served by the shared-memory eps table / DBSCAN paths
constexpr int BIG_SLICES = 64;      // CTAs per item in the large-ray-count k-NN kernel

struct Ctx {
  int device;
  int table_cap;
  double* knn_unit;  // [table_cap] mean 10-NN distance on the unit sphere, 0 = not computed
  int* knn_need;     // [table_cap]
  long long launches;
};

// =================================================================================================
// k_prepare
// =================================================================================================
#ifndef PREP_THREADS_N
#define PREP_THREADS_N 256
#endif
constexpr int PREP_THREADS = PREP_THREADS_N;
constexpr int AVG_TILE = 1024;       // average-diameter rays per tile (4 per thread)
// doubles of k_prepare's fixed shared-memory part (reductions, class counters, direction grid) and of its
// per-atom arrays (x, y, z, r, w + int type)
constexpr size_t PREP_FIXED_DOUBLES = (64 * 8 + (64 + PWB_MAX_TYPES + 1 + DG_BINS + 2) * 4 + DG_BINS * DG_CAP * 2 + 15) / 16 * 2;
#define PREP_ATOM_DOUBLES(NA) ((size_t)(NA) * 5 + ((size_t)(NA) + 1) / 2)

// Rarely taken paths kept out of line so the compiler does not if-convert a double-precision sqrt
// (~40 instructions) into the hot loops.
__device__ __noinline__ double rare_exit_param(double t, double under, double sb) { return fmax(sb, t + sqrt(under)); }
__device__ __noinline__ ValIdx rare_pair_max(ValIdx mb, double d2, double rr, int idx) {
  ValIdx c; c.v = sqrt(d2) + rr; c.i = idx;
  return vi_max(mb, c);
}

#ifndef PREP_MINB
#define PREP_MINB 4
#endif
// BIG: the per-atom arrays live in the global scratch region (molecules beyond the shared-memory budget); a
// separate instantiation, so that the common case keeps pure shared-memory addressing (LDS, not generic loads)
template <bool BIG>
__global__ void __launch_bounds__(PREP_THREADS, PREP_MINB)
k_prepare(pwb_batch b, Ws ws, int do_avg) {
  extern __shared__ double sm[];
  const int item = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  // blockIdx.y > 0: extra CTAs that only share the average-diameter rays of a huge item (see the tile loop)
  const bool lead = blockIdx.y == 0;
  const int off = b.item_off[item];
  const int n = b.item_off[item + 1] - off;
  double* rec = b.records + (size_t)item * b.rec_stride;
  double* idv = ws.item_d + (size_t)item * ITEM_D;
  int* iiv = ws.item_i + (size_t)item * ITEM_I;
  const int NA = b.max_atoms;
  // fixed-size part in shared memory, then the per-atom arrays -- in shared memory too, or (huge molecules) in
  // this CTA's slot of the global scratch
  double* red = sm;                 // 64 doubles
  int* redi = (int*)(red + 64);     // 64 ints
  int* cnt = redi + 64;             // [PWB_MAX_TYPES + 1]
  DirGrid DG;
  DG.cnt = cnt + PWB_MAX_TYPES + 1;                      // [DG_BINS]
  DG.overflow = DG.cnt + DG_BINS;                        // [1] (+1 pad)
  DG.list = (unsigned short*)(DG.overflow + 2);          // [DG_BINS][DG_CAP]
  double* abase;
  if (BIG) abase = ws.atom_scratch + (size_t)item * ws.scratch_item + (size_t)blockIdx.y * PREP_ATOM_DOUBLES(NA);
  else abase = sm + PREP_FIXED_DOUBLES;
  double* sx = abase;         // original order
  double* sy = sx + NA;
  double* sz = sy + NA;
  double* sr = sz + NA;
  double* sw = sr + NA;       // mass, later scratch
  int* stype = (int*)(sw + NA);     // [NA]
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int c = tid; c <= PWB_MAX_TYPES; c += nt) cnt[c] = 0;
  __syncthreads();
  if (n <= 0 || n > NA) {
    if (tid == 0 && lead) { rec[PWB_REC_STATUS] = PWB_ST_BAD_ITEM; iiv[I_STATUS] = PWB_ST_BAD_ITEM; iiv[I_N] = 0; }
    return;
  }
  const int* tp = b.atom_type + (b.shared_types ? 0 : off);
  for (int a = tid; a < n; a += nt) {
    const double* p = b.coords + (size_t)(off + a) * 3;
    sx[a] = p[0]; sy[a] = p[1]; sz[a] = p[2];
    int t = tp[a];
    if (t < 0 || t >= b.n_types) { s_bad = 1; t = 0; }
    stype[a] = t;
    sr[a] = b.type_vdw[t];
    sw[a] = b.type_mass[t];
    atomicAdd(&cnt[t + 1], 1);
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0 && lead) { rec[PWB_REC_STATUS] = PWB_ST_BAD_ITEM; iiv[I_STATUS] = PWB_ST_BAD_ITEM; iiv[I_N] = 0; }
    return;
  }
  if (tid == 0) {
    for (int c = 1; c <= PWB_MAX_TYPES; c++) cnt[c] += cnt[c - 1];   // cnt[c] = start of class c
    if (lead) {
      for (int c = 0; c <= PWB_MAX_TYPES; c++) iiv[I_CLS0 + c] = cnt[c < b.n_types ? c : b.n_types];
      iiv[I_N] = n;
      iiv[I_STATUS] = 0;
    }
  }
  __syncthreads();
  // stable bucket scatter to the workspace (x, y, z, r) + original index.  Position of atom a = start of its class
  // + the atoms of that class before it: counted per chunk of `nt` atoms with match_any (equal types inside the
  // warp) and a small per-warp table (lower warps of the chunk), plus a running count over earlier chunks.
  if (lead) {
    __shared__ int wcnt[PREP_THREADS / 32][PWB_MAX_TYPES];
    __shared__ int run[PWB_MAX_TYPES];
    const int lane = tid & 31, wd = tid >> 5, nw = nt >> 5;
    if (tid < PWB_MAX_TYPES) run[tid] = 0;
    for (int base = 0; base < n; base += nt) {
      const int a = base + tid;
      const bool valid = a < n;
      const int t = valid ? stype[a] : -1;
      for (int k = tid; k < nw * PWB_MAX_TYPES; k += nt) (&wcnt[0][0])[k] = 0;
      __syncthreads();
      const unsigned peers = __match_any_sync(0xffffffffu, t);
      const int inwarp = __popc(peers & ((1u << lane) - 1u));
      if (valid && inwarp == 0) wcnt[wd][t] = __popc(peers);
      __syncthreads();
      if (valid) {
        int below = run[t];
        for (int k = 0; k < wd; k++) below += wcnt[k][t];
        const int pos = cnt[t] + below + inwarp;
        double* q = ws.atoms + (size_t)(off + pos) * 4;
        q[0] = sx[a]; q[1] = sy[a]; q[2] = sz[a]; q[3] = sr[a];
        ws.atom_idx[off + pos] = a;
      }
      __syncthreads();
      if (tid < PWB_MAX_TYPES) { int tot = 0; for (int k = 0; k < nw; k++) tot += wcnt[k][tid]; run[tid] += tot; }
      __syncthreads();               // the table is cleared at the top of the next chunk
    }
  }
  // ---- centre of mass (utilities.py:127-148)
  double mx = 0, my = 0, mz = 0, mt = 0;
  for (int a = tid; a < n; a += nt) { double m = sw[a]; mx += sx[a] * m; my += sy[a] * m; mz += sz[a] * m; mt += m; }
  mx = block_sum(mx, red); my = block_sum(my, red); mz = block_sum(mz, red); mt = block_sum(mt, red);
  const double cx = mx / mt, cy = my / mt, cz = mz / mt;
  // ---- pore diameter at the centre of mass (utilities.py:375-388)
  // rho_a = |x_a - com| + r_a (kept in sw: the masses are done with) bounds every pair: d_ij + r_i + r_j
  // <= rho_i + rho_j.
  ValIdx pb; pb.v = PWB_INF; pb.i = 0x7fffffff;
  ValIdx fa; fa.v = -PWB_INF; fa.i = 0x7fffffff;
  __syncthreads();                      // everyone has read the masses in sw
  for (int a = tid; a < n; a += nt) {
    double dx = sx[a] - cx, dy = sy[a] - cy, dz = sz[a] - cz;
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    ValIdx c; c.v = d - sr[a]; c.i = a;
    pb = vi_min(pb, c);
    c.v = d + sr[a]; sw[a] = c.v;
    fa = vi_max(fa, c);
  }
  pb = block_vi(pb, false, red, redi);
  fa = block_vi(fa, true, red, redi);   // outermost atom
  // ---- maximum dimension (utilities.py:355-372): upper triangle incl. diagonal, first arg-max.
  // Exact pruning: the outermost atom and its farthest partner give a lower bound LB; an atom with
  // rho_i + rho_max < LB cannot be part of a maximal pair.  For a cage that leaves the outer shell
  // (~1/4 of the atoms, 1/16 of the pairs).  Among the kept pairs the sqrt is only taken when the pair
  // can beat the thread's running maximum: d + ri + rj > best <=> d^2 > (best - ri - rj)^2.
  ValIdx mb; mb.v = -PWB_INF; mb.i = 0x7fffffff;
  {
    const double xi = sx[fa.i], yi = sy[fa.i], zi = sz[fa.i], ri = sr[fa.i];
    for (int j = tid; j < n; j += nt) {
      double dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
      ValIdx c; c.v = sqrt(dx * dx + dy * dy + dz * dz) + (ri + sr[j]);
      c.i = fa.i <= j ? fa.i * n + j : j * n + fa.i;
      mb = vi_max(mb, c);
    }
  }
  mb = block_vi(mb, true, red, redi);
  int* cand = stype;                    // the types are in the workspace by now; reuse as candidate list
  __shared__ int s_ncand;
  if (tid == 0) s_ncand = 0;
  __syncthreads();
  {
    const double need = mb.v - fa.v - 1e-9;
    for (int a = tid; a < n; a += nt) if (sw[a] >= need) cand[atomicAdd(&s_ncand, 1)] = a;
  }
  __syncthreads();
  {
    const int m = s_ncand;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    for (int u = wid; u < m; u += nw) {
      const int i = cand[u];
      const double xi = sx[i], yi = sy[i], zi = sz[i], ri = sr[i];
      for (int v = u + lane; v < m; v += 32) {
        const int j = cand[v];
        double dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
        double d2 = dx * dx + dy * dy + dz * dz;
        double rr = ri + sr[j];
        double thr = (mb.v - rr) * (1.0 - 1e-12);
        if (thr < 0 || d2 >= thr * thr) mb = rare_pair_max(mb, d2, rr, i <= j ? i * n + j : j * n + i);
      }
    }
  }
  mb = block_vi(mb, true, red, redi);
  const double D = mb.v;
  if (tid == 0 && lead) {
    rec[PWB_REC_STATUS] = 0;
    rec[PWB_REC_NATOMS] = n;
    rec[PWB_REC_COM] = cx; rec[PWB_REC_COM + 1] = cy; rec[PWB_REC_COM + 2] = cz;
    rec[PWB_REC_MAXD] = D; rec[PWB_REC_MAXD_A1] = mb.i / n; rec[PWB_REC_MAXD_A2] = mb.i % n;
    rec[PWB_REC_PORE_D] = 2.0 * pb.v; rec[PWB_REC_PORE_ATOM] = pb.i;
    idv[D_COM] = cx; idv[D_COM + 1] = cy; idv[D_COM + 2] = cz; idv[D_MAXD] = D;
  }
  if (!do_avg) return;                  // (launched with gridDim.y = 1 then)
  // ---- average diameter (utilities.py:1586-1650, :1556-1583): sphere radius = D (not D/2).
  // Atom i counts for ray n iff r_i^2 - d_i^2 > 0 and |o+(t-h)n| < |o+(t+h)n| <=> t + o.n > 0; the exit
  // point is o + s n with s = t + sqrt(r_i^2 - d_i^2) and |o + s n| grows with s, so the farthest exit
  // is the largest s.  s <= t + r_i prunes the sqrt.
  __syncthreads();
  double ox = 0, oy = 0, oz = 0;
  for (int a = tid; a < n; a += nt) {
    sx[a] -= cx; sy[a] -= cy; sz[a] -= cz;          // shift_com with com_adjust = 0
    ox += sx[a]; oy += sy[a]; oz += sz[a];
  }
  ox = block_sum(ox, red) / n; oy = block_sum(oy, red) / n; oz = block_sum(oz, red) / n;
  // centroid-relative positions in place, sw = r^2 - |L|^2.  (Round 1 also rank-sorted the atoms by |L| + r so
  // that the outermost ones set the running maximum first; with the direction-bin lists deciding the visiting
  // order that O(n^2) sort only cost 10 % of the kernel's instructions -- max() does not care about the order.)
  for (int a = tid; a < n; a += nt) {
    const double lx = sx[a] - ox, ly = sy[a] - oy, lz = sz[a] - oz, ra = sr[a];
    sx[a] = lx; sy[a] = ly; sz[a] = lz; sw[a] = ra * ra - (lx * lx + ly * ly + lz * lz);
  }
  __syncthreads();
  const int P = n_sampling_points(D, b.adjust);
  const double oo = ox * ox + oy * oy + oz * oz;
  dg_build(DG, sx, sy, sz, sr, n, sqrt(oo));
  const bool use_grid = *DG.overflow == 0;
  // Rays are processed in tiles of AVG_TILE; the per-tile sums are added in tile order, so the result does
  // not depend on how many CTAs (gridDim.y) share the tiles of one item.
  const int ntiles = (P + AVG_TILE - 1) / AVG_TILE;
  const bool solo = gridDim.y == 1 && ntiles == 1;
  double* part = ws.clear_a2 + (size_t)item * b.ray_cap;     // [ntiles][2] partial sums (free until k_rays)
  double acc_all = 0, hits_all = 0;
  for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
  const int base = tile * AVG_TILE, tend = min(P, base + AVG_TILE);
  double acc = 0, hits = 0;
  if (use_grid) {
    for (int k = base + tid; k < tend; k += nt) {
      double px, py, pz;
      spiral_point(k, P, D, px, py, pz);
      const double nr = sqrt(px * px + py * py + pz * pz);
      const double nx = px / nr, ny = py / nr, nz = pz / nr;
      const double on = -(ox * nx + oy * ny + oz * nz);
      const int bin = dg_bin(nx, ny, nz), m = DG.cnt[bin];
      const unsigned short* lst = DG.list + bin * DG_CAP;
      double sb = -PWB_INF;
      for (int e = 0; e < m; e++) {
        const int a = lst[e];
        const double t = fma(sx[a], nx, fma(sy[a], ny, sz[a] * nz));
        const double under = fma(t, t, sw[a]);
        if (under > 0.0 && t > on && t + sr[a] > sb) sb = rare_exit_param(t, under, sb);
      }
      if (sb > -PWB_INF) { acc += sqrt(fma(sb, sb, fma(-2.0 * sb, on, oo))); hits += 1.0; }   // |o + s n|, o.n = -on
    }
  } else {
  constexpr int RPT = 4;
  for (int k0 = base + tid * RPT; k0 < tend; k0 += nt * RPT) {
    double nx[RPT], ny[RPT], nz[RPT], on[RPT], sb[RPT];
#pragma unroll
    for (int u = 0; u < RPT; u++) {
      double px, py, pz;
      spiral_point(min(k0 + u, P - 1), P, D, px, py, pz);
      double nr = sqrt(px * px + py * py + pz * pz);
      nx[u] = px / nr; ny[u] = py / nr; nz[u] = pz / nr;
      on[u] = -(ox * nx[u] + oy * ny[u] + oz * nz[u]);
      sb[u] = -PWB_INF;
    }
    for (int a = 0; a < n; a++) {
      const double Lx = sx[a], Ly = sy[a], Lz = sz[a], W = sw[a], ra = sr[a];
#pragma unroll
      for (int u = 0; u < RPT; u++) {
        double t = fma(Lx, nx[u], fma(Ly, ny[u], Lz * nz[u]));
        double under = fma(t, t, W);
        if (under > 0.0 && t > on[u] && t + ra > sb[u]) sb[u] = rare_exit_param(t, under, sb[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < RPT; u++) {
      if (k0 + u < tend && sb[u] > -PWB_INF) {
        double s = sb[u];
        acc += sqrt(fma(s, s, fma(-2.0 * s, on[u], oo)));     // |o + s n|, o.n = -on
        hits += 1.0;
      }
    }
  }
  }
  acc = block_sum(acc, red); hits = block_sum(hits, red);
  if (solo) { acc_all = acc; hits_all = hits; }
  else if (tid == 0) { part[2 * tile] = acc; part[2 * tile + 1] = hits; }
  }
  if (!solo) {
    // the last CTA of the item to arrive adds the tile sums in tile order
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      s_last = 1;
      if (gridDim.y > 1) {
        const int prev = atomicAdd(&iiv[I_TILES_PREP], 1);
        s_last = prev == (int)gridDim.y - 1;
        if (s_last) iiv[I_TILES_PREP] = 0;
      }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) for (int t = 0; t < ntiles; t++) { acc_all += __ldcg(&part[2 * t]); hits_all += __ldcg(&part[2 * t + 1]); }
  }
  if (tid == 0) rec[PWB_REC_AVGD] = 2.0 * (acc_all / hits_all);
}

// =================================================================================================
// k_opt_pore: opt_pore_diameter (utilities.py:400-426) -- 32/G concurrent L-BFGS-B searches per warp
// =================================================================================================
// One pore-centre search is a serial optimiser: ~100 objective evaluations separated by scalar bookkeeping
// (Cauchy point, subspace step, dense-B update, More-Thuente step).  A group of G lanes owns one problem: the
// lanes share the atom loop of the objective (group_min_clear) and run the bookkeeping redundantly, so a warp
// spends its bookkeeping issue slots on 32/G problems at once instead of one.  The solver is the resumable
// LbSolver: every trip of the loop below evaluates the objective of all groups of the warp at the SAME call site
// (converged code), then each group advances its own state machine (groups in different branches serialise) and
// pulls the next item from the work queue when it finishes -- iteration counts vary 10x between frames, so a
// static item -> group map would leave lanes waiting for the slowest problem.
// Measured on B200, 100 000 CC3 frames, ms per 10 000 frames (profiles/r2_opt_pore_variants.md): the kernel is
// latency bound (one instruction per ~10 cycles per warp), so what counts is problems in flight per SM times
// work shared per instruction.  G = 32 with atom tiles in shared memory (round 1: 32 problems per SM) 2.21;
// G = 16 / 8 with shared tiles 2.56 / 3.28 (same 32 problems, fewer warps to hide latency); G = 8 with the tiles
// read through L1 / L2 (1.3 KB of state per problem in shared memory, 96 problems per SM at 80 registers) 1.96;
// G = 4 with the objective evaluated over an exact survivor list (OPT_LIST below: the <= 64 atoms that can attain the
// minimum within 0.5 A of a reference point, group_build_list / group_min_list) 1.52 -- what the list shortens is the
// latency of an evaluation (a dozen loads instead of 42 per lane); 1.39 in the limit of long launches, the rest is
// the tail of the slowest searches (5 x the mean number of evaluations).
#ifndef OPT_G
#define OPT_G 4                      // lanes per problem (1 ... 32)
#endif
#ifndef OPT_ATOMS_SMEM
#define OPT_ATOMS_SMEM 0             // atom tile of each problem in shared memory (else read through L1 / L2)
#endif
#ifndef OPT_WARPS_N
#define OPT_WARPS_N 4
#endif
constexpr int OPT_WARPS = OPT_WARPS_N;
#ifndef OPT_MINB
#define OPT_MINB 4                   // CTAs per SM the register budget must allow (112 registers per thread)
#endif
// doubles of shared memory per problem: double4[NA] atoms (optional), the solver state (incl. bounds), class
// starts (17 ints, padded to 20); even, so that the double4 tiles stay 16-byte aligned
// First pruning experiment, superseded by OPT_LIST: per-atom flags (group_min_clear_pruned) -- bit-identical records,
// ~95 % of the atoms skipped in late line-search steps and NO change in kernel time (1.956 vs 1.964 ms per 10 000
// frames, profiles/r2_opt_pore_variants.md): the loop still visits every atom.  Off by default (0.7 KB of shared
// memory per problem).
#ifndef OPT_PRUNE
#define OPT_PRUNE 0
#endif
#ifndef OPT_PRUNE_RBIG
#define OPT_PRUNE_RBIG 0.5           // re-reference when two consecutive evaluations lie farther than this (A)
#endif
// (+ with OPT_PRUNE: reference point [3], its clearance, and float[NA] reference squared distances)
#ifndef OPT_HIST_GLOBAL
#define OPT_HIST_GLOBAL 0            // 1: (s, y) ring of every search in global memory, 0.9 instead of 1.4 KB of shared
                                     // memory per problem (pays only with one or two lanes per problem)
#endif
#define OPT_SOLVER_DOUBLES (OPT_HIST_GLOBAL ? LB_STATE(3) : LB_STORAGE(3))
#ifndef OPT_LIST
#define OPT_LIST 1                   // survivor list of the atoms that can attain the minimum near a reference point
#endif
#ifndef OPT_LIST_CAP
#define OPT_LIST_CAP 64              // entries (16-bit atom indices)
#endif
#ifndef OPT_LIST_CLASSWISE
#define OPT_LIST_CLASSWISE 0
#endif
#ifndef OPT_LIST_EAGER
#define OPT_LIST_EAGER 0             // 1: re-reference at the first evaluation outside the ball
#endif
#ifndef OPT_LIST_R
#define OPT_LIST_R 0.5               // radius (A) of the ball around the reference point the list is valid for
#endif
#define OPT_PROB_RAW(NA, SMEM_ATOMS) (((SMEM_ATOMS) ? (NA) * 4 : 0) + OPT_SOLVER_DOUBLES + 10 + (OPT_PRUNE ? 4 + ((NA) + 1) / 2 : 0) + (OPT_LIST ? 4 + OPT_LIST_CAP / 4 + 5 : 0))
// The lanes of a warp address the SAME offset in 32 / G different problems; OPT_STRIDE_REM pads the per-problem stride
// to that remainder (mod 16 doubles) to spread those addresses over the shared-memory banks
#ifdef OPT_STRIDE_REM
#define OPT_PROB_DOUBLES(NA, SMEM_ATOMS) (OPT_PROB_RAW(NA, SMEM_ATOMS) + ((OPT_STRIDE_REM + 16 - OPT_PROB_RAW(NA, SMEM_ATOMS) % 16) % 16))
#else      // (measured: no effect -- the accesses are not bank-conflict bound; even, so that double4 tiles stay aligned)
#define OPT_PROB_DOUBLES(NA, SMEM_ATOMS) ((OPT_PROB_RAW(NA, SMEM_ATOMS) + 1) & ~(size_t)1)
#endif

template <int G, bool GENERAL, bool ATOMS_SMEM>
__global__ void __launch_bounds__(OPT_WARPS * 32, OPT_MINB)
k_opt_pore(pwb_batch b, Ws ws) {
  extern __shared__ __align__(16) double sm[];
  constexpr int PPW = 32 / G;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
  const unsigned mask = group_mask<G>(lane);
  double* smp = sm + (size_t)(w * PPW + lane / G) * OPT_PROB_DOUBLES(b.max_atoms, ATOMS_SMEM);
  double4* p4 = (double4*)smp;                                        // x, y, z, radius (if ATOMS_SMEM)
  double* hist = smp + (ATOMS_SMEM ? (size_t)b.max_atoms * 4 : 0);    // [OPT_SOLVER_DOUBLES] solver state
  int* cls = (int*)(hist + OPT_SOLVER_DOUBLES);                       // [PWB_MAX_TYPES+1] class starts
#if OPT_PRUNE
  double* pref = (double*)(cls + 20);                                 // [4] reference point and its clearance
  float* d2ref = (float*)(pref + 4);                                  // [NA] squared distances to the reference point
  bool have_ref = false, last_far = false;
#endif
#if OPT_LIST
  double* lref = (double*)(cls + 20);                                 // [4] reference point (the clearance there is in a register)
  unsigned short* lst = (unsigned short*)(lref + 4);                  // [OPT_LIST_CAP] atoms that can attain the minimum
  unsigned short* lcls = lst + OPT_LIST_CAP;                          // [PWB_MAX_TYPES+1] class starts inside the list
  int nl = -2;                                                        // -2 no list yet, -1 too many survivors
  bool last_far = false;
#endif
  int* queue = ws.item_i + I_QUEUE;                                   // zeroed by run_opt before the launch
  LbSolver<3, GENERAL ? -1 : 2> S;
  S.st = hist;
  S.hs = OPT_HIST_GLOBAL ? nullptr : hist + LB_STATE(3);
  GroupAtoms A;
  A.gl = gl; A.mask = mask; A.cls = cls; A.ncls = b.n_types; A.n = 0; A.p = p4;
  const int* orig = nullptr;
  int item = -1, nfev = 0, fstop = 0;
  bool active = false, exhausted = false, fin = false;
  for (;;) {
#ifdef OPT_CTA_SYNC
    // experiment: keep the warps of a CTA in step so that they share instruction-cache lines
    if (!__syncthreads_or(active || fin || !exhausted)) break;
#endif
    if (!active && !exhausted) {
      // ---- pull the next item; set its search up (or finish it at once when there is nothing to optimise)
      int it = 0;
      if (gl == 0) it = atomicAdd(queue, 1);
      it = __shfl_sync(mask, it, gbase);
      if (it >= b.n_items) exhausted = true;
      else {
        item = it;
        int* iiv = ws.item_i + (size_t)item * ITEM_I;
        double* idv = ws.item_d + (size_t)item * ITEM_D;
        double* rec = b.records + (size_t)item * b.rec_stride;
        if (!(iiv[I_STATUS] & PWB_ST_BAD_ITEM)) {
          const int n = iiv[I_N], off = b.item_off[item];
          __syncwarp(mask);                                           // the previous item's tile is no longer read
          for (int c = gl; c <= PWB_MAX_TYPES; c += G) cls[c] = iiv[I_CLS0 + c];
          const double4* src = (const double4*)(ws.atoms + (size_t)off * 4);
          if (ATOMS_SMEM) { for (int a = gl; a < n; a += G) p4[a] = src[a]; A.p = p4; }
          else A.p = src;
          A.n = n;
          orig = ws.atom_idx + off;
          __syncwarp(mask);
          double c[3] = {idv[D_COM], idv[D_COM + 1], idv[D_COM + 2]};
          double lo3[3] = {0, 0, 0}, hi3[3] = {0, 0, 0};
          int nb3[3] = {2, 2, 2};
          bool search = true, bad = false;
          nfev = 0;
#if OPT_LIST
          nl = -2; last_far = false;
#endif
          if (b.centre_override) {
            const double* o = b.centre_override + (size_t)item * 3;
            c[0] = o[0]; c[1] = o[1]; c[2] = o[2];
            search = false;
          } else {
            const bool has_start = b.opt_start != nullptr;
            if (has_start) {   // opt_pore_diameter(com=...): start there; the default box is built around it (:411-420)
              const double* o = b.opt_start + (size_t)item * 3;
              c[0] = o[0]; c[1] = o[1]; c[2] = o[2];
            }
            if (b.opt_bounds) {
              const double* ob = b.opt_bounds + (size_t)item * 6;
#pragma unroll
              for (int i = 0; i < 3; i++) {
                const double l = ob[2 * i], h = ob[2 * i + 1];
                const bool fl = l > -PWB_INF, fh = h < PWB_INF;       // +-inf (or beyond 1e300) = no bound
                bad |= fl && fh && l > h;
                lo3[i] = l; hi3[i] = h; nb3[i] = fl ? (fh ? 2 : 1) : (fh ? 3 : 0);
              }
            } else {
              const double pore_r = has_start ? group_pore_argmin<G>(A, orig, c[0], c[1], c[2]).v : rec[PWB_REC_PORE_D] / 2.0;
              bad = pore_r < 0;
#pragma unroll
              for (int i = 0; i < 3; i++) { lo3[i] = c[i] - pore_r; hi3[i] = c[i] + pore_r; }
            }
          }
          if (bad) {   // scipy raises on inverted bounds; the reference has no handler
            if (gl == 0) { iiv[I_STATUS] |= PWB_ST_OPT_BOUNDS; rec[PWB_REC_STATUS] = iiv[I_STATUS]; }
          } else if (search) {
            S.init(hist, OPT_HIST_GLOBAL ? ws.opt_hist + (size_t)item * LB_HISTORY(3) : hist + LB_STATE(3), c, lo3, hi3, nb3);
#if OPT_PRUNE
            have_ref = false; last_far = false;
#endif
            __syncwarp(mask);
            active = true;
          } else {     // forced centre: parked where a finished search leaves its result
            if (gl < 3) S.xs()[gl] = c[gl];
            if (gl == 0) S.iter() = 0;
            fin = true; fstop = -1;
          }
        }
      }
    }
#ifndef OPT_CTA_SYNC
    if (!__any_sync(0xffffffffu, active || fin)) {
      if (__all_sync(0xffffffffu, exhausted)) break;
      continue;
    }
#endif
    if (active) {
      // ---- the objective: -pore_diameter and its bound-aware forward differences (utilities.py:391-397 +
      // scipy's 2-point scheme), 4 points per evaluation
      const double* x = S.xs();
      double px[4], py[4], pz[4], h[3], m[4];
#pragma unroll
      for (int k = 0; k < 4; k++) { px[k] = x[0]; py[k] = x[1]; pz[k] = x[2]; }
#pragma unroll
      for (int i = 0; i < 3; i++) h[i] = lb_fd_step(x[i], S.nbd(i), S.lo()[i], S.hi()[i]);
      px[1] = x[0] + h[0]; py[2] = x[1] + h[1]; pz[3] = x[2] + h[2];
#if OPT_PRUNE
      {
        // distance of this evaluation's stencil from the reference point (+ the forward-difference steps)
        const double ex = x[0] - pref[0], ey = x[1] - pref[1], ez = x[2] - pref[2];
        const double hmax = fmax(fabs(h[0]), fmax(fabs(h[1]), fabs(h[2])));
        const double delta = (sqrt(ex * ex + ey * ey + ez * ez) + hmax) * (1.0 + 1e-12) + 1e-12;
        const bool far = !(delta <= OPT_PRUNE_RBIG);
        const bool store = !have_ref || (far && last_far);     // two far evaluations in a row: the search has moved
        last_far = far && !store;
        __syncwarp(mask);
        group_min_clear_pruned<G, 4>(A, px, py, pz, m, d2ref, store ? 0.0 : pref[3] + 2.0 * delta, store);
#ifdef OPT_PRUNE_CHECK
        {
          double m2[4];
          group_min_clear<G, 4>(A, px, py, pz, m2);
          if (m2[0] != m[0] || m2[1] != m[1] || m2[2] != m[2] || m2[3] != m[3]) {
            if (gl == 0) printf("prune mismatch item %d store %d delta %.3e aref %.6f m %.17g %.17g m2 %.17g %.17g\n", item, (int)store, delta, pref[3], m[0], m[1], m2[0], m2[1]);
          }
        }
#endif
        if (store) {
          __syncwarp(mask);
          if (gl < 3) pref[gl] = x[gl];
          if (gl == 3) pref[3] = m[0];
          have_ref = true;
        }
        __syncwarp(mask);
      }
#elif OPT_LIST
      {
        double delta = PWB_INF;
        if (nl != -2) {
          const double ex = x[0] - lref[0], ey = x[1] - lref[1], ez = x[2] - lref[2];
          // every point of the stencil is within |x - ref| + max |h| of the reference point
          const double hmax = fmax(fabs(h[0]), fmax(fabs(h[1]), fabs(h[2])));
          delta = (sqrt(ex * ex + ey * ey + ez * ez) + hmax) * (1.0 + 1e-12) + 1e-12;
        }
        const bool far = !(delta <= OPT_LIST_R);
        // (per atom for >= 4 lanes per problem: more square roots but no reduction per class in the dependent chain
        // -- 1.50 against 1.78 ms; class-wise below that)
        if (nl >= 0 && !far && !ws.opt_no_list) {
          if (G >= 4 && !OPT_LIST_CLASSWISE) group_min_list<G, 4>(A, lst, nl, px, py, pz, m);
          else group_min_clear_list<G, 4>(A, lst, lcls, px, py, pz, m);
          last_far = false;
#ifdef OPT_LIST_STATS
          if (gl == 0) { atomicAdd(&queue[1], 1); atomicAdd(&queue[4], nl); }
#endif
        }
        else {
          group_min_clear<G, 4>(A, px, py, pz, m);
#ifdef OPT_LIST_STATS
          if (gl == 0) atomicAdd(&queue[2], 1);
#endif
          if (nl == -2 || (far && (last_far || OPT_LIST_EAGER))) {   // first evaluation, or two far ones in a row: the search moved
            __syncwarp(mask);
            for (int i = gl; i < 3; i += G) lref[i] = x[i];
            nl = group_build_list<G>(A, x[0], x[1], x[2], m[0], 2.0 * OPT_LIST_R, lst, lcls, OPT_LIST_CAP);
#ifdef OPT_LIST_STATS
            if (gl == 0) atomicAdd(&queue[3], 1);
#endif
            last_far = false;
            __syncwarp(mask);
          } else last_far = far;
        }
#ifdef OPT_LIST_CHECK
        {
          double m2[4];
          group_min_clear<G, 4>(A, px, py, pz, m2);
          if (m2[0] != m[0] || m2[1] != m[1] || m2[2] != m[2] || m2[3] != m[3])
            if (gl == 0) printf("list mismatch item %d nl %d delta %.3e m %.17g %.17g m2 %.17g %.17g\n", item, nl, delta, m[0], m[1], m2[0], m2[1]);
        }
#endif
      }
#else
      group_min_clear<G, 4>(A, px, py, pz, m);
#endif
      nfev += 4;
      const double f = -(m[0] * 2.0);
      double* g = S.gs();
      __syncwarp(mask);
#pragma unroll
      for (int i = 0; i < 3; i++) g[i] = lb_div(-(m[i + 1] * 2.0) - f, (x[i] + h[i]) - x[i]);
      __syncwarp(mask);
      const int st = S.advance(f, 2.220446049250313e-09, 1e-5, 20, 15000);
      __syncwarp(mask);
      if (st >= 0) { fin = true; fstop = st; active = false; }
    }
    if (fin) {   // one epilogue for searched and forced centres: pore_diameter at the final centre (:424-425)
      fin = false;
      __syncwarp(mask);
      const double fc0 = S.xs()[0], fc1 = S.xs()[1], fc2 = S.xs()[2];
      const int fnit = S.iter();
      const ValIdx pa = group_pore_argmin<G>(A, orig, fc0, fc1, fc2);
      if (gl == 0) {
        double* idv = ws.item_d + (size_t)item * ITEM_D;
        double* rec = b.records + (size_t)item * b.rec_stride;
        idv[D_CENTRE] = fc0; idv[D_CENTRE + 1] = fc1; idv[D_CENTRE + 2] = fc2;
        rec[PWB_REC_OPT_D] = 2.0 * pa.v; rec[PWB_REC_OPT_ATOM] = pa.i;
        rec[PWB_REC_OPT_C] = fc0; rec[PWB_REC_OPT_C + 1] = fc1; rec[PWB_REC_OPT_C + 2] = fc2;
        rec[PWB_REC_OPT_NIT] = fnit; rec[PWB_REC_OPT_STOP] = fstop; rec[PWB_REC_OPT_NFEV] = fstop < 0 ? 0 : nfev;
      }
      __syncwarp(mask);
    }
  }
}

// =================================================================================================
// k_rays
// =================================================================================================
#ifndef RAY_THREADS_N
#define RAY_THREADS_N 256
#endif
constexpr int RAY_THREADS = RAY_THREADS_N;
constexpr int RAY_TILE = PWB_RAY_TILE;   // rays per pass (4 per thread, register tiled)
constexpr int RAY_RPT = RAY_TILE / RAY_THREADS;
constexpr int RAY_ACAP = RAY_TILE;   // 2*min clearance of the survivors of one tile
// doubles of k_rays' fixed shared-memory part (a multiple of 4: the double4 atom tile behind it stays 32-byte
// aligned) and of its per-atom arrays
constexpr size_t RAY_FIXED_DOUBLES = ((64 + PWB_MAX_TYPES + RAY_ACAP) * 8 + (PWB_MAX_TYPES + 1 + 32 + 2 * RAY_TILE + DG_BINS + 2) * 4 +
                                      DG_BINS * DG_CAP * 2 + 31) / 32 * 4;
#define RAY_ATOM_DOUBLES(NA) ((size_t)(NA) * 8)

// Exclusive block scan of one int per thread; returns this thread's offset, `total` for all.
__device__ __forceinline__ int block_excl_scan(int v, int* wsum, int& total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
  for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
  __syncthreads();
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  for (int k = 0; k < nw; k++) { int s = wsum[k]; if (k < w) base += s; tot += s; }
  total = tot;
  return base + inc - v;
}

// Direction n = p/|p| of ray k, its length |p| and the point p itself (utilities.py:1139, :1108)
__device__ __forceinline__ void ray_dir(int k, int P, double R, double& px, double& py, double& pz, double& nr) {
  spiral_point(k, P, R, px, py, pz);
  nr = sqrt(px * px + py * py + pz * pz);
}

#ifndef RAY_MINB
#define RAY_MINB 4
#endif
template <bool BIG>
__global__ void __launch_bounds__(RAY_THREADS, RAY_MINB)
k_rays(pwb_batch b, Ws ws, Ctx ctx) {
  extern __shared__ __align__(16) double sm[];
  const int item = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  int* iiv = ws.item_i + (size_t)item * ITEM_I;
  double* idv = ws.item_d + (size_t)item * ITEM_D;
  double* rec = b.records + (size_t)item * b.rec_stride;
  const int status0 = iiv[I_STATUS];
  if (status0 & (PWB_ST_BAD_ITEM | PWB_ST_OPT_BOUNDS)) return;
  // blockIdx.y > 0: extra CTAs sharing the ray tiles of a huge item (config 5: 1e5 rays on one molecule)
  const bool lead = blockIdx.y == 0;
  const int n = iiv[I_N], off = b.item_off[item], NA = b.max_atoms;
  // shared memory: [per-atom arrays | fixed part]; BIG: the per-atom arrays are in the global scratch instead
  double* red = sm + (BIG ? 0 : RAY_ATOM_DOUBLES(b.max_atoms));   // 64
  double* cr = red + 64;                    // [PWB_MAX_TYPES]
  double* sa = cr + PWB_MAX_TYPES;          // [RAY_ACAP] path clearances of the current sub-pass
  int* cls = (int*)(sa + RAY_ACAP);         // [PWB_MAX_TYPES+1]
  int* wsum = cls + PWB_MAX_TYPES + 1;      // [32]
  int* surv = wsum + 32;                    // [RAY_TILE] ray index of survivors in this tile
  int* spos = surv + RAY_TILE;              // [RAY_TILE] arg-min path step (or -1)
  DirGrid DG;
  DG.cnt = spos + RAY_TILE;                 // [DG_BINS]
  DG.overflow = DG.cnt + DG_BINS;           // [1] (+1 pad)
  DG.list = (unsigned short*)(DG.overflow + 2);   // [DG_BINS][DG_CAP]
  // per-atom arrays: shared memory behind the fixed part, or this CTA's slot of the global scratch (huge molecules)
  double* abase;
  if (BIG) abase = ws.atom_scratch + (size_t)item * ws.scratch_item + (size_t)blockIdx.y * RAY_ATOM_DOUBLES(NA);
  else abase = sm;
  double4* ap = (double4*)abase;              // [NA] shifted, bucketed atoms: x, y, z, |x|^2
  double* lx = (double*)(ap + NA); double* ly = lx + NA; double* lz = ly + NA; double* lw = lz + NA;  // about the centroid; r^2-|L|^2
#ifdef RAYS_FP32_FILTER
  // experiment: FP32 copy (Lx, Ly, Lz, r^2 - |L|^2) for a filter in front of the exact FP64 pre-test
  float4* lf = (float4*)(red + RAY_FIXED_DOUBLES);
  __shared__ float s_eps_u, s_eps_t;
#endif

  // shift exactly as find_windows does (utilities.py:1379-1393)
  const double cmx = idv[D_COM], cmy = idv[D_COM + 1], cmz = idv[D_COM + 2];
  double shx, shy, shz, icx, icy, icz;
  if (b.pore_opt) {
    double adx = cmx - idv[D_CENTRE], ady = cmy - idv[D_CENTRE + 1], adz = cmz - idv[D_CENTRE + 2];
    icx = cmx - adx; icy = cmy - ady; icz = cmz - adz;      // initial_com - com_adjust
    shx = cmx - adx; shy = cmy - ady; shz = cmz - adz;      // COM - com_adjust
  } else { icx = cmx; icy = cmy; icz = cmz; shx = cmx; shy = cmy; shz = cmz; }
  for (int c = tid; c <= PWB_MAX_TYPES; c += nt) cls[c] = iiv[I_CLS0 + c];
  for (int c = tid; c < PWB_MAX_TYPES; c += nt) cr[c] = c < b.n_types ? b.type_vdw[c] : 0.0;
  double ox = 0, oy = 0, oz = 0;
  for (int a = tid; a < n; a += nt) {
    const double* q = ws.atoms + (size_t)(off + a) * 4;
    double x = q[0] - shx, y = q[1] - shy, z = q[2] - shz;
    ap[a] = make_double4(x, y, z, x * x + y * y + z * z);
    lw[a] = q[3] * q[3];
    ox += x; oy += y; oz += z;
  }
  ox = block_sum(ox, red) / n; oy = block_sum(oy, red) / n; oz = block_sum(oz, red) / n;
  for (int a = tid; a < n; a += nt) {
    double x = ap[a].x - ox, y = ap[a].y - oy, z = ap[a].z - oz;
    lx[a] = x; ly[a] = y; lz[a] = z;
  }
  __syncthreads();
  // direction grid over the centroid-relative atoms (radii are still in lw as r^2 here)
  for (int c = tid; c < DG_BINS; c += nt) DG.cnt[c] = 0;
  if (tid == 0) *DG.overflow = (n > 65535) ? 1 : 0;
  __syncthreads();
  {
    const double onorm = sqrt(ox * ox + oy * oy + oz * oz);
    for (int a = tid; a < n; a += nt) {
      const double x = lx[a], y = ly[a], z = lz[a], ra = sqrt(lw[a]);
      const double L2 = x * x + y * y + z * z, L = sqrt(L2);
      lw[a] = lw[a] - L2;                                                       // r^2 - |L|^2
      if (L <= ra * (1.0 + 1e-9) + 1e-12) { for (int c = 0; c < DG_BINS; c++) dg_push(DG, c, a); continue; }
      const double sa = fmin(1.0, ra / L + 2e-7), il = 1.0 / L;   // sin of the cap radius + margin (see dg_build)
      dg_insert_cap(DG, a, x * il, y * il, z * il, sa);
      if (L - ra <= onorm + 1e-9) dg_insert_cap(DG, a, -x * il, -y * il, -z * il, sa);
    }
#ifdef RAYS_FP32_FILTER
    // Rigorous FP32 error bounds (u = 2^-24): |t32 - t| <= 5.1 u |L| and |under32 - under| <= 7.3e-7 (|L|^2 + r^2)
    // (conversion of L, n and W, three rounded products / FMAs, one squared term); the margins below are larger.
    double m2 = 0;
    for (int a = tid; a < n; a += nt) {
      const double x = lx[a], y = ly[a], z = lz[a], L2 = x * x + y * y + z * z;
      lf[a] = make_float4((float)x, (float)y, (float)z, (float)lw[a]);
      m2 = fmax(m2, 2.0 * L2 + lw[a]);                    // |L|^2 + r^2
    }
    m2 = warp_max(m2);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = m2;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int k = 0; k < (nt >> 5); k++) t = fmax(t, red[k]);
      s_eps_u = (float)(1e-6 * t + 1e-30);
      s_eps_t = (float)(5e-7 * (sqrt(t) + onorm) + 1e-30);
    }
#endif
  }
  const double R = idv[D_MAXD] / 2.0;     // max_dim is translation invariant: reuse (utilities.py:1398)
  const int P = n_sampling_points(R, b.adjust);
  __syncthreads();
  const bool use_grid = *DG.overflow == 0;
  if (tid == 0 && lead) {
    idv[D_RADIUS] = R; idv[D_ICOM] = icx; idv[D_ICOM + 1] = icy; idv[D_ICOM + 2] = icz;
    idv[D_SHIFT] = shx; idv[D_SHIFT + 1] = shy; idv[D_SHIFT + 2] = shz;
    iiv[I_P] = P;
    rec[PWB_REC_NRAYS] = P; rec[PWB_REC_RADIUS] = R;
    if (P >= 2 && P < ctx.table_cap && P <= KNN_SMEM_CAP) ctx.knn_need[P] = 1;
  }
  if (P > b.ray_cap || (P <= KNN_SMEM_CAP && P >= ctx.table_cap) || P < 2) {
    if (tid == 0 && lead) { iiv[I_STATUS] |= PWB_ST_RAY_CAP; rec[PWB_REC_STATUS] = iiv[I_STATUS]; iiv[I_NCLEAR] = 0; iiv[I_NSURV] = 0; }
    return;
  }
  int* out_k = ws.clear_k + (size_t)item * b.ray_cap;
  int* out_pos = ws.clear_pos + (size_t)item * b.ray_cap;
  double* out_a2 = ws.clear_a2 + (size_t)item * b.ray_cap;
  // Tiles of RAY_TILE rays; the clear rays of a tile are written at the tile's own offset (they fit: at most
  // RAY_TILE of them) and, when an item has more than one tile, packed in tile order afterwards -- by
  // whichever CTA of the item finishes last -- so the lists stay sorted by ray index whatever gridDim.y is.
  const int ntiles = (P + RAY_TILE - 1) / RAY_TILE;
  // ray sharding over GPUs (pwb_batch.ray_tile_begin/end): only the tiles [tb, te) of this item are analysed here
  const bool sharded = b.ray_tile_end > 0;
  const int tb = sharded ? min(b.ray_tile_begin, ntiles) : 0, te = sharded ? min(b.ray_tile_end, ntiles) : ntiles;
  const int Pend = min(P, te * RAY_TILE);
  const bool solo = gridDim.y == 1 && ntiles == 1 && !sharded;
  int* tile_cnt = ws.labels + (size_t)item * b.ray_cap;      // [ntiles] clear rays, [ntiles] survivors (free until k_cluster)
  int n_clear = 0, n_surv_total = 0;
  for (int base = (tb + blockIdx.y) * RAY_TILE; base < Pend; base += gridDim.y * RAY_TILE) {
    n_clear = 0;
    // ---- pre-test (utilities.py:1132-1161), RAY_RPT consecutive rays per thread, branch free:
    // atom i counts iff r_i^2 - d_i^2 > 0 (d_i about the line through the centroid o) and
    // |o+(t-h)n| < |o+(t+h)n|  <=>  t + o.n > 0   (both norms expand to |o|^2+2s(o.n)+s^2, h > 0)
    double nx[RAY_RPT], ny[RAY_RPT], nz[RAY_RPT], on[RAY_RPT];
    int blk[RAY_RPT];
#pragma unroll
    for (int u = 0; u < RAY_RPT; u++) {
      int k = base + tid * RAY_RPT + u;
      double px, py, pz, nr;
      ray_dir(k < P ? k : 0, P, R, px, py, pz, nr);
      nx[u] = px / nr; ny[u] = py / nr; nz[u] = pz / nr;
      on[u] = -(ox * nx[u] + oy * ny[u] + oz * nz[u]);
      blk[u] = k >= P;
    }
    if (use_grid) {
#pragma unroll
      for (int u = 0; u < RAY_RPT; u++) {
        if (blk[u]) continue;
        const int bin = dg_bin(nx[u], ny[u], nz[u]), m = DG.cnt[bin];
        const unsigned short* lst = DG.list + bin * DG_CAP;
        int hit = 0;
#ifdef RAYS_FP32_FILTER
        const float nxf = (float)nx[u], nyf = (float)ny[u], nzf = (float)nz[u], onf = (float)on[u];
        const float eu = s_eps_u, et = s_eps_t;
        for (int e = 0; e < m; e++) {
          const int a = lst[e];
          const float4 Lf = lf[a];
          const float tf = fmaf(Lf.x, nxf, fmaf(Lf.y, nyf, Lf.z * nzf));
          const float uf = fmaf(tf, tf, Lf.w);
          const bool sure = uf > eu && tf > onf + et;           // a hit whatever the rounding
          const bool maybe = uf > -eu && tf > onf - et;         // cannot be excluded in FP32
          hit |= (int)sure;
          if (maybe && !sure) {                                 // rare: decide exactly
            const double t = fma(lx[a], nx[u], fma(ly[a], ny[u], lz[a] * nz[u]));
            const double under = fma(t, t, lw[a]);
            hit |= (int)(under > 0.0) & (int)(t > on[u]);
          }
        }
#else
        for (int e = 0; e < m; e++) {
          const int a = lst[e];
          const double t = fma(lx[a], nx[u], fma(ly[a], ny[u], lz[a] * nz[u]));
          const double under = fma(t, t, lw[a]);
          hit |= (int)(under > 0.0) & (int)(t > on[u]);
        }
#endif
        blk[u] = hit;
      }
    } else {
#pragma unroll 2
    for (int a = 0; a < n; a++) {
      const double Lx = lx[a], Ly = ly[a], Lz = lz[a], W = lw[a];
#pragma unroll
      for (int u = 0; u < RAY_RPT; u++) {
        double t = fma(Lx, nx[u], fma(Ly, ny[u], Lz * nz[u]));
        double under = fma(t, t, W);
        blk[u] |= (int)(under > 0.0) & (int)(t > on[u]);
      }
    }
    }
    int cntm = 0;
#pragma unroll
    for (int u = 0; u < RAY_RPT; u++) cntm += blk[u] ? 0 : 1;
    int nsurv;
    int o = block_excl_scan(cntm, wsum, nsurv);
#pragma unroll
    for (int u = 0; u < RAY_RPT; u++) if (!blk[u]) surv[o++] = base + tid * RAY_RPT + u;
    __syncthreads();
    n_surv_total += nsurv;
    // ---- stepped path clearance (utilities.py:1100-1129).  The reference evaluates
    // a_j = min_i(|x_i - q_j| - r_i) at every step q_j = j*chunk and keeps min_j a_j, its first arg-min
    // and "all a_j > 0".  For a fixed atom |x_i - q_j|^2 = |x|^2 + j*(j*|chunk|^2 - 2 x.chunk) is a
    // parabola in j, so min_j is attained at the step nearest to the atom's projection (clamped to the
    // path; an exact half-way projection takes the smaller step, the reference's first arg-min):
    // min_j min_i = min_i d_i(j_i) -- ONE instead of chunks+1 evaluations per atom.
    // (Splitting each surviving ray over 4 lanes to even out the last, nearly empty round of this loop was
    // measured 70 % SLOWER: the per-ray set-up -- spiral point, floor division, three divisions -- is paid per
    // warp round whatever the number of active lanes, so fewer rays per round cost more than the idle tail.)
    for (int s = tid; s < nsurv; s += nt) {
      double px, py, pz, nr;
      ray_dir(surv[s], P, R, px, py, pz, nr);
      const int chunks = (int)py_floordiv(nr, b.increment);
      const double cxs = px / chunks, cys = py / chunks, czs = pz / chunks;
      const double cc = cxs * cxs + cys * cys + czs * czs, icc = 1.0 / cc;
      const int jmax = chunks > 0 ? chunks : 0;     // steps 0..chunks inclusive (range(chunks + 1), :1112)
      double bestv = PWB_INF;
      int bestj = 0;
      for (int c = 0; c < b.n_types; c++) {
        const int cb = cls[c], ce = cls[c + 1];
        if (cb == ce) continue;
        double m = PWB_INF;
        int mj = 0;
        for (int a = cb; a < ce; a++) {
          const double4 q = ap[a];
          const double g = fma(q.x, cxs, fma(q.y, cys, q.z * czs));
          // nearest step to the projection (ties -> smaller), clamped to the path: round up in the conversion,
          // clamp as integers (one min-max instruction instead of two double selects)
          const int ji = min(jmax, max(0, __double2int_ru(fma(g, icc, -0.5))));
          const double jm = (double)ji;
          const double dm = fma(jm, fma(jm, cc, -2.0 * g), q.w);
          if (dm < m) { m = dm; mj = ji; }
        }
        const double v = sqrt(fmax(m, 0.0)) - cr[c];
        if (v < bestv || (v == bestv && mj < bestj)) { bestv = v; bestj = mj; }
      }
      const bool clear = chunks > 0 && bestv > 0.0;
      spos[s] = clear ? bestj : -1;
      sa[s] = bestv * 2.0;
    }
    __syncthreads();
    // ---- ordered compaction of the clear rays into the global lists
    for (int sb = 0; sb < nsurv; sb += nt) {
      int s = sb + tid;
      int ok = (s < nsurv && spos[s] >= 0) ? 1 : 0;
      int tot;
      int oo = block_excl_scan(ok, wsum, tot);
      if (ok) { out_k[base + n_clear + oo] = surv[s]; out_pos[base + n_clear + oo] = spos[s]; out_a2[base + n_clear + oo] = sa[s]; }
      n_clear += tot;
    }
    if (!solo && tid == 0) { tile_cnt[base / RAY_TILE] = n_clear; tile_cnt[ntiles + base / RAY_TILE] = nsurv; }
    __syncthreads();
  }
  if (!solo) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      s_last = 1;
      if (gridDim.y > 1) {
        const int prev = atomicAdd(&iiv[I_TILES_RAYS], 1);
        s_last = prev == (int)gridDim.y - 1;
        if (s_last) iiv[I_TILES_RAYS] = 0;
      }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    n_clear = 0; n_surv_total = 0;
    for (int t = tb; t < te; t++) {
      const int c = __ldcg(&tile_cnt[t]), src = t * RAY_TILE;
      n_surv_total += __ldcg(&tile_cnt[ntiles + t]);
      if (src != n_clear) {
        for (int e0 = 0; e0 < c; e0 += nt) {               // dest <= src: read, barrier, write
          const int e = e0 + tid;
          int k = 0, ps = 0; double a2 = 0;
          if (e < c) { k = __ldcg(&out_k[src + e]); ps = __ldcg(&out_pos[src + e]); a2 = __ldcg(&out_a2[src + e]); }
          __syncthreads();
          if (e < c) { out_k[n_clear + e] = k; out_pos[n_clear + e] = ps; out_a2[n_clear + e] = a2; }
          __syncthreads();
        }
      }
      n_clear += c;
    }
  }
  if (tid == 0) {
    iiv[I_NCLEAR] = n_clear; iiv[I_NSURV] = n_surv_total;
    rec[PWB_REC_NCLEAR] = n_clear; rec[PWB_REC_NSURV] = n_surv_total;
    if (n_clear == 0 && !sharded) { iiv[I_STATUS] |= PWB_ST_NO_CLEAR_RAY; rec[PWB_REC_STATUS] = iiv[I_STATUS]; }
  }
}

// =================================================================================================
// k_knn_fill: m_unit(P) = mean over the P unit-sphere spiral points of their 10 smallest distances
// (self included), i.e. KDTree.query(k=10) of utilities.py:1427-1433 divided by the radius.
// =================================================================================================
constexpr int KNN_THREADS = 256;
__global__ void __launch_bounds__(KNN_THREADS)
k_knn_fill(Ctx ctx, int kcap) {
  extern __shared__ double sm[];
  const int P = blockIdx.x + 2;
  if (P >= kcap) return;
  if (!ctx.knn_need[P] || ctx.knn_unit[P] != 0.0) return;
  double* ux = sm; double* uy = ux + kcap; double* uz = uy + kcap;
  __shared__ double red[64];
  for (int i = threadIdx.x; i < P; i += blockDim.x) spiral_point(i, P, 1.0, ux[i], uy[i], uz[i]);
  __syncthreads();
  double acc = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    double xi = ux[i], yi = uy[i], zi = uz[i];
    double best[10];
#pragma unroll
    for (int q = 0; q < 10; q++) best[q] = PWB_INF;
    for (int j = 0; j < P; j++) {
      double dx = xi - ux[j], dy = yi - uy[j], dz = zi - uz[j];
      double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best[9]) {
        best[9] = d2;
#pragma unroll
        for (int q = 9; q > 0; q--)
          if (best[q] < best[q - 1]) { double t = best[q]; best[q] = best[q - 1]; best[q - 1] = t; }
      }
    }
#pragma unroll
    for (int q = 0; q < 10; q++) acc += sqrt(best[q]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) ctx.knn_unit[P] = acc / (10.0 * P);
}

// =================================================================================================
// k_cluster: DBSCAN label semantics of sklearn (min_samples = 5, dist <= eps), order-free rule:
// core = #neighbours (self included) >= 5; clusters = components of the core-core graph numbered
// by their smallest core index; border point -> smallest cluster among adjacent cores; else noise.
// Then per cluster the first arg-max of column 1 (2*min clearance), utilities.py:1221.
// =================================================================================================
constexpr int CL_THREADS = 256;
constexpr unsigned CL_SMALL_SMEM = 26 * 1024;   // pass-0 budget: S <= ~330 clear rays, 8 CTAs per SM

__device__ __forceinline__ int uf_find(volatile int* parent, int i) {
  int p = parent[i];
  while (p != parent[p]) p = parent[p];
  return p;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {   // smaller index becomes the root
  while (true) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { int t = a; a = b; b = t; }
    int old = atomicCAS(&parent[b], b, a);
    if (old == b) return;
  }
}

// Two launches: pass 0 with a small shared-memory budget (high occupancy) handles every item whose
// S clear rays fit, pass 1 with the full 200 KB handles the rest (large molecules / dense ray sets).
__global__ void __launch_bounds__(CL_THREADS)
k_cluster(pwb_batch b, Ws ws, Ctx ctx, int pass, unsigned budget) {
  extern __shared__ double sm[];
  const int item = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  int* iiv = ws.item_i + (size_t)item * ITEM_I;
  double* idv = ws.item_d + (size_t)item * ITEM_D;
  double* rec = b.records + (size_t)item * b.rec_stride;
  const int st = iiv[I_STATUS];
  if (st & (PWB_ST_BAD_ITEM | PWB_ST_OPT_BOUNDS | PWB_ST_RAY_CAP)) { if (tid == 0) iiv[I_NWIN] = 0; return; }
  if (pass == 1 && iiv[I_DEFER] != 1) return;
  const int S = iiv[I_NCLEAR], P = iiv[I_P];
  if (P > KNN_SMEM_CAP) {                        // eps table not available: large-ray-count path
    if (pass == 0 && tid == 0) { iiv[I_DEFER] = 2; iiv[I_NWIN] = 0; rec[PWB_REC_NWIN] = 0; }
    return;
  }
  const double R = idv[D_RADIUS];
  const double m = R * ctx.knn_unit[P];
  const double eps = m + sqrt(m);
  if (pass == 0 && tid == 0) { idv[D_EPS] = eps; rec[PWB_REC_EPS] = eps; rec[PWB_REC_NWIN] = 0; iiv[I_NWIN] = 0; iiv[I_DEFER] = 0; }
  if (S == 0) return;
  const int cap = b.ray_cap;
  const int Wd = (S + 31) >> 5;
  // shared memory carved by the ACTUAL number of clear rays
  const size_t need_base = 32 * (size_t)S + 12 * (size_t)b.win_cap + 8 * (size_t)S + 4 * (32 + 128) + ((S + 7) & ~7) + 64;
  const bool use_bits = need_base + 4 * (size_t)S * Wd <= budget;
  if (pass == 0 ? !use_bits : need_base > budget) {
    if (tid == 0) iiv[I_DEFER] = (pass == 0) ? 1 : 2;   // 2: does not fit any shared-memory path
    return;
  }
  double* px = sm;                                               // [S][4] points
  unsigned long long* cbest = (unsigned long long*)(px + 4 * (size_t)S);   // [win_cap] best a2 bits per cluster
  int* parent = (int*)(cbest + b.win_cap);                       // [S]
  int* label = parent + S;                                       // [S]
  int* cidx = label + S;                                         // [win_cap] arg-max member per cluster
  int* wsum = cidx + b.win_cap;                                  // [32]
  unsigned* cm = (unsigned*)(wsum + 32);                         // [128] core mask words (S <= 4096)
  unsigned char* core = (unsigned char*)(cm + 128);              // [S]
  unsigned* adj = (unsigned*)(core + ((S + 7) & ~7));            // [S*Wd] adjacency bit rows (if use_bits)
  const int* ck = ws.clear_k + (size_t)item * cap;
  const double* ca2 = ws.clear_a2 + (size_t)item * cap;
  // points as (x, y, z, |p|^2/2): |pi-pj|^2 <= eps^2  <=>  pi.pj - |pj|^2/2 >= (|pi|^2 - eps^2)/2
  double4* pt = (double4*)px;                      // [cap] overlays px/py/pz (+ the slack behind them)
  for (int i = tid; i < S; i += nt) {
    double x, y, z;
    spiral_point(ck[i], P, R, x, y, z);
    pt[i] = make_double4(x, y, z, 0.5 * (x * x + y * y + z * z));
    parent[i] = i;
    label[i] = 0;                                  // neighbour counter in pass 1
  }
  __syncthreads();
  const double eps2 = eps * eps, epsz = eps * (1.0 + 1e-12) + 1e-9;
  // ---- adjacency as a bit matrix in shared memory (one FP64 test per ordered pair, once); core
  // flags, components (min-label propagation with pointer jumping) and border labels are then pure
  // integer work on the bit rows.  Falls back to recomputing distances if S is too large for the tile.
  const int Q = max(1, (4 * nt + S - 1) / S), jl = (S + Q - 1) / Q;   // fallback path: (i, slice of j) items
  if (use_bits) {
    // item = (word w, point i) with i fastest: a warp shares the 32 partner points (broadcast loads)
    // and the bit rows are stored transposed, adj[w*S + i], so row scans are conflict free
    for (int it = tid; it < S * Wd; it += nt) {
      const int w = it / S, i = it - w * S;
      const double4 a = pt[i];
      const double hi = a.w - 0.5 * eps2;
      const int jb = w << 5, jn = min(32, S - jb);
      unsigned bits = 0;
      // the clear rays are sorted by spiral index, i.e. by descending z: a word whose z range is more than
      // eps away from z_i cannot hold a neighbour (consecutive i share the verdict, so whole warps skip)
      if (a.z - epsz <= pt[jb].z && a.z + epsz >= pt[jb + jn - 1].z) {
        for (int t = 0; t < jn; t++) {
          const double4 o = pt[jb + t];
          bits |= (unsigned)(fma(a.x, o.x, fma(a.y, o.y, fma(a.z, o.z, -o.w))) >= hi) << t;
        }
      }
      adj[it] = bits;
    }
    __syncthreads();
    for (int i = tid; i < S; i += nt) {
      int c = 0;
      for (int w = 0; w < Wd; w++) c += __popc(adj[w * S + i]);
      core[i] = c >= 5;
      parent[i] = (c >= 5) ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int w = tid; w < Wd; w += nt) {
      unsigned mbits = 0;
      for (int t = 0; t < 32 && (w << 5) + t < S; t++) mbits |= (unsigned)core[(w << 5) + t] << t;
      cm[w] = mbits;
    }
    __syncthreads();
    int changed;
    do {
      changed = 0;
      for (int i = tid; i < S; i += nt) {
        if (!core[i]) continue;
        int m0 = ((volatile int*)parent)[i], m = m0;
        for (int w = 0; w < Wd; w++) {
          unsigned bits = adj[w * S + i] & cm[w];
          while (bits) {
            const int j = (w << 5) + __ffs(bits) - 1;
            bits &= bits - 1;
            m = min(m, ((volatile int*)parent)[j]);
          }
        }
        m = min(m, ((volatile int*)parent)[m]);
        if (m < m0) { atomicMin(&parent[i], m); changed = 1; }
      }
      changed = __syncthreads_or(changed);
    } while (changed);
  } else {
  for (int it = tid; it < S * Q; it += nt) {
    const int i = it / Q, q = it - i * Q;
    const double4 a = pt[i];
    const double hi = a.w - 0.5 * eps2;
    const int j1 = min(S, (q + 1) * jl);
    int c = 0;
    for (int j = q * jl; j < j1; j++) {
      const double4 o = pt[j];
      c += (fma(a.x, o.x, fma(a.y, o.y, fma(a.z, o.z, -o.w))) >= hi);
    }
    if (c) atomicAdd(&label[i], c);
  }
  __syncthreads();
  for (int i = tid; i < S; i += nt) core[i] = label[i] >= 5;
  __syncthreads();
  // union pass: warp-uniform trip counts and an explicit reconvergence point per step, otherwise the
  // divergent union-find calls serialise the lanes for the rest of the loop
  for (int it0 = 0; it0 < S * Q; it0 += nt) {
    const int it = it0 + tid;
    const bool valid = it < S * Q;
    const int i = valid ? it / Q : 0, q = valid ? it - i * Q : 0;
    const bool act = valid && core[i];
    const double4 a = pt[i];
    const double hi = a.w - 0.5 * eps2;
    const int jb = q * jl, j1 = act ? min(i, (q + 1) * jl) : 0;
    for (int s = 0; s < jl; s++) {
      const int j = jb + s;
      bool hit = false;
      if (j < j1 && core[j]) {
        const double4 o = pt[j];
        hit = fma(a.x, o.x, fma(a.y, o.y, fma(a.z, o.z, -o.w))) >= hi;
      }
      if (hit) uf_union(parent, i, j);
      __syncwarp();
    }
  }
  __syncthreads();
  }
  // roots in ascending index order -> cluster numbers
  int nclus = 0;
  for (int s0 = 0; s0 < S; s0 += nt) {
    int i = s0 + tid;
    int isroot = (i < S && core[i] && parent[i] == i) ? 1 : 0;
    int tot;
    int o = block_excl_scan(isroot, wsum, tot);
    if (isroot) label[i] = nclus + o;
    nclus += tot;
  }
  __syncthreads();
  for (int i = tid; i < S; i += nt)
    if (core[i] && parent[i] != i) label[i] = label[uf_find(parent, i)];
  __syncthreads();
  for (int i = tid; i < S; i += nt) if (!core[i]) parent[i] = 0x7fffffff;   // reuse parent[] as the border label
  __syncthreads();
  if (use_bits) {
    for (int i = tid; i < S; i += nt) {
      if (core[i]) continue;
      int best = 0x7fffffff;
      for (int w = 0; w < Wd; w++) {
        unsigned bits = adj[w * S + i] & cm[w];
        while (bits) { const int j = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; best = min(best, label[j]); }
      }
      parent[i] = best;
    }
    __syncthreads();
  } else {
  for (int it0 = 0; it0 < S * Q; it0 += nt) {
    const int it = it0 + tid;
    const bool valid = it < S * Q;
    const int i = valid ? it / Q : 0, q = valid ? it - i * Q : 0;
    const bool act = valid && !core[i];
    if (__any_sync(0xffffffffu, act)) {                 // border / noise points are rare
      const double4 a = pt[i];
      const double hi = a.w - 0.5 * eps2;
      const int jb = q * jl, j1 = act ? min(S, (q + 1) * jl) : 0;
      int best = 0x7fffffff;
      for (int s = 0; s < jl; s++) {
        const int j = jb + s;
        if (j < j1 && core[j]) {
          const double4 o = pt[j];
          if (fma(a.x, o.x, fma(a.y, o.y, fma(a.z, o.z, -o.w))) >= hi) best = min(best, label[j]);
        }
      }
      if (best != 0x7fffffff) atomicMin(&parent[i], best);
    }
  }
  __syncthreads();
  }
  for (int i = tid; i < S; i += nt) if (!core[i]) label[i] = (parent[i] == 0x7fffffff) ? -1 : parent[i];
  const int W = min(nclus, b.win_cap);
  for (int c = tid; c < b.win_cap; c += nt) { cbest[c] = 0ull; cidx[c] = 0x7fffffff; }
  __syncthreads();
  int* glabel = ws.labels + (size_t)item * cap;
  for (int i = tid; i < S; i += nt) {
    int l = label[i];
    glabel[i] = l;
    if (l >= 0 && l < W) atomicMax(&cbest[l], (unsigned long long)__double_as_longlong(ca2[i]));  // a2 > 0
  }
  __syncthreads();
  for (int i = tid; i < S; i += nt) {
    int l = label[i];
    if (l >= 0 && l < W && (unsigned long long)__double_as_longlong(ca2[i]) == cbest[l]) atomicMin(&cidx[l], i);
  }
  __syncthreads();
  int* wk = ws.win_k + (size_t)item * b.win_cap;
  for (int c = tid; c < W; c += nt) wk[c] = ck[cidx[c]];
  if (tid == 0) {
    iiv[I_NWIN] = W; rec[PWB_REC_NWIN] = W;
    if (nclus > b.win_cap) { iiv[I_STATUS] |= PWB_ST_WIN_CAP; rec[PWB_REC_STATUS] = iiv[I_STATUS]; }
  }
}

// =================================================================================================
// Large ray counts (adjust >> 1, P up to ~1e6; BASELINE config 5).  The golden spiral is monotone in z
// (z_k falls by 2/P per point), so two points closer than d differ by at most d*P/2 in index: both the
// 10-NN mean (eps) and the DBSCAN neighbourhoods only need an index window instead of all pairs.
//   k_big_points  unit-sphere points of the item's ray set -> big_pts
//   k_knn_big     BIG_SLICES CTAs per item: partial sums of the 10 smallest distances (stored in .w)
//   k_cluster_big one CTA per item: windowed DBSCAN with union-find in global memory
// =================================================================================================
__global__ void __launch_bounds__(256)
k_big_points(pwb_batch b, Ws ws) {
  const int item = blockIdx.y;
  const int* iiv = ws.item_i + (size_t)item * ITEM_I;
  if (iiv[I_DEFER] != 2 || (iiv[I_STATUS] & (PWB_ST_BAD_ITEM | PWB_ST_OPT_BOUNDS | PWB_ST_RAY_CAP))) return;
  const int P = iiv[I_P];
  double4* U = (double4*)(ws.big_pts + (size_t)item * b.ray_cap * 4);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P; k += gridDim.x * blockDim.x) {
    double x, y, z;
    spiral_point(k, P, 1.0, x, y, z);
    U[k] = make_double4(x, y, z, 0.0);
  }
}

__global__ void __launch_bounds__(256)
k_knn_big(pwb_batch b, Ws ws) {
  const int item = blockIdx.y, slice = blockIdx.x;
  const int* iiv = ws.item_i + (size_t)item * ITEM_I;
  if (iiv[I_DEFER] != 2 || (iiv[I_STATUS] & (PWB_ST_BAD_ITEM | PWB_ST_OPT_BOUNDS | PWB_ST_RAY_CAP))) return;
  const int P = iiv[I_P];
  double4* U = (double4*)(ws.big_pts + (size_t)item * b.ray_cap * 4);
  __shared__ double red[64];
  // a disc holding 10 of P uniformly spread points has radius sqrt(40/P); 1.5x that bounds the 10th
  // neighbour everywhere on the spiral (checked per point, full scan otherwise)
  const double dmax = 1.5 * sqrt(40.0 / (double)P), dmax2 = dmax * dmax;
  const int Wk = (int)ceil(dmax * P * 0.5) + 2;
  const int per = (P + BIG_SLICES - 1) / BIG_SLICES;
  const int i0 = slice * per, i1 = min(P, i0 + per);
  double acc = 0;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const double4 a = U[i];
    double best[10];
    int lo = max(0, i - Wk), hi = min(P, i + Wk + 1);
    for (int attempt = 0; attempt < 2; attempt++) {
#pragma unroll
      for (int q = 0; q < 10; q++) best[q] = PWB_INF;
      for (int j = lo; j < hi; j++) {
        const double4 o = U[j];
        const double dx = a.x - o.x, dy = a.y - o.y, dz = a.z - o.z;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best[9]) {
          best[9] = d2;
#pragma unroll
          for (int q = 9; q > 0; q--)
            if (best[q] < best[q - 1]) { double t = best[q]; best[q] = best[q - 1]; best[q - 1] = t; }
        }
      }
      if (best[9] <= dmax2 || (lo == 0 && hi == P)) break;   // window provably contained the 10 nearest
      lo = 0; hi = P;
    }
#pragma unroll
    for (int q = 0; q < 10; q++) acc += sqrt(best[q]);
  }
  acc = block_sum(acc, red);
  __syncthreads();
  if (threadIdx.x == 0) ((double*)&U[slice])[3] = acc;      // partial sum parked in the unused .w
}

__device__ __forceinline__ int uf_find_halving(int* parent, int i) {
  while (true) {
    int p = ((volatile int*)parent)[i];
    if (p == i) return i;
    int gp = ((volatile int*)parent)[p];
    if (gp != p) parent[i] = gp;           // path halving (benign race: gp is still an ancestor)
    i = p;
  }
}
__device__ __forceinline__ void uf_union_g(int* parent, int a, int b) {
  while (true) {
    a = uf_find_halving(parent, a); b = uf_find_halving(parent, b);
    if (a == b) return;
    if (a > b) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[b], b, a) == b) return;
  }
}

// Windowed DBSCAN for one big item, split into phases so that the O(S * window) passes run on many
// CTAs (grid = blocks x items); phases 3 and 6 are single-CTA ordered scans.
//   0 count neighbours -> core flags   1 union core pairs   2 flatten   3 number the roots (ordered)
//   4 labels of the other core points  5 border labels      6 arg-max ray per cluster
constexpr int CLB_THREADS = 256;
__global__ void __launch_bounds__(CLB_THREADS)
k_cluster_big(pwb_batch b, Ws ws, int phase) {
  const int item = blockIdx.y, nt = blockDim.x * gridDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int tid = threadIdx.x;
  int* iiv = ws.item_i + (size_t)item * ITEM_I;
  if (iiv[I_DEFER] != 2 || (iiv[I_STATUS] & (PWB_ST_BAD_ITEM | PWB_ST_OPT_BOUNDS | PWB_ST_RAY_CAP))) return;
  double* idv = ws.item_d + (size_t)item * ITEM_D;
  double* rec = b.records + (size_t)item * b.rec_stride;
  const int S = iiv[I_NCLEAR], P = iiv[I_P], cap = b.ray_cap;
  const double R = idv[D_RADIUS];
  const double4* U = (const double4*)(ws.big_pts + (size_t)item * cap * 4);
  __shared__ int wsum[32];
  __shared__ unsigned long long cbest[64];
  __shared__ int cidx[64];
  double t = 0;
  for (int sl = 0; sl < BIG_SLICES; sl++) t += U[sl].w;     // fixed order: deterministic
  const double m = R * (t / (10.0 * (double)P));
  const double eps = m + sqrt(m), eps2 = eps * eps;
  if (phase == 0 && gtid == 0) { idv[D_EPS] = eps; rec[PWB_REC_EPS] = eps; rec[PWB_REC_NWIN] = 0; iiv[I_NWIN] = 0; }
  if (S == 0) return;
  const int Wk = (int)ceil((eps / R) * P * 0.5) + 2;
  const int* ck = ws.clear_k + (size_t)item * cap;
  const double* ca2 = ws.clear_a2 + (size_t)item * cap;
  int* parent = ws.big_i + (size_t)item * 2 * cap;
  int* aux = parent + cap;                    // core flag
  int* label = ws.labels + (size_t)item * cap;
  // window [jlo, jhi) of clear rays whose ray index is within Wk of ray i's (ck is ascending)
  auto window = [&](int i, int& jlo, int& jhi) {
    const int klo = ck[i] - Wk, khi = ck[i] + Wk;
    int a = 0, c = i;
    while (a < c) { int mid = (a + c) >> 1; if (ck[mid] < klo) a = mid + 1; else c = mid; }
    jlo = a;
    a = i; c = S;
    while (a < c) { int mid = (a + c) >> 1; if (ck[mid] <= khi) a = mid + 1; else c = mid; }
    jhi = a;
  };
  auto adjacent = [&](const double4& ui, int j) {
    const double4 uj = U[ck[j]];
    const double dx = R * ui.x - R * uj.x, dy = R * ui.y - R * uj.y, dz = R * ui.z - R * uj.z;
    return dx * dx + dy * dy + dz * dz <= eps2;
  };
  if (phase == 0) {
    for (int i = gtid; i < S; i += nt) {
      int jlo, jhi, c = 0;
      window(i, jlo, jhi);
      const double4 ui = U[ck[i]];
      for (int j = jlo; j < jhi; j++) c += adjacent(ui, j);
      aux[i] = c >= 5;
      parent[i] = i;
    }
  } else if (phase == 1) {
    for (int i = gtid; i < S; i += nt) {
      if (!aux[i]) continue;
      int jlo, jhi;
      window(i, jlo, jhi);
      const double4 ui = U[ck[i]];
      for (int j = jlo; j < i; j++)
        if (aux[j] && adjacent(ui, j)) uf_union_g(parent, i, j);
    }
  } else if (phase == 2) {
    for (int i = gtid; i < S; i += nt) if (aux[i]) parent[i] = uf_find_halving(parent, i);
  } else if (phase == 3) {
    if (blockIdx.x != 0) return;
    int nclus = 0;
    for (int s0 = 0; s0 < S; s0 += blockDim.x) {
      const int i = s0 + tid;
      const int isroot = (i < S && aux[i] && parent[i] == i) ? 1 : 0;
      int tot;
      const int o = block_excl_scan(isroot, wsum, tot);
      if (isroot) label[i] = nclus + o;
      nclus += tot;
    }
    if (tid == 0) iiv[I_NWIN] = nclus;        // number of clusters (clamped to the window slots in phase 6)
  } else if (phase == 4) {
    // roots were final before phase 3 ran, so parent[i] is a root index here
    for (int i = gtid; i < S; i += nt) if (aux[i] && parent[i] != i) label[i] = label[parent[i]];
  } else if (phase == 5) {
    for (int i = gtid; i < S; i += nt) {
      if (aux[i]) continue;
      int jlo, jhi, best = 0x7fffffff;
      window(i, jlo, jhi);
      const double4 ui = U[ck[i]];
      for (int j = jlo; j < jhi; j++)
        if (aux[j] && adjacent(ui, j)) best = min(best, label[j]);
      label[i] = (best == 0x7fffffff) ? -1 : best;
    }
  } else {
    if (blockIdx.x != 0) return;
    const int nclus = iiv[I_NWIN];
    const int W = min(nclus, min(b.win_cap, 64));
    for (int c = tid; c < 64; c += blockDim.x) { cbest[c] = 0ull; cidx[c] = 0x7fffffff; }
    __syncthreads();
    for (int i = tid; i < S; i += blockDim.x) {
      const int l = label[i];
      if (l >= 0 && l < W) atomicMax(&cbest[l], (unsigned long long)__double_as_longlong(ca2[i]));
    }
    __syncthreads();
    for (int i = tid; i < S; i += blockDim.x) {
      const int l = label[i];
      if (l >= 0 && l < W && (unsigned long long)__double_as_longlong(ca2[i]) == cbest[l]) atomicMin(&cidx[l], i);
    }
    __syncthreads();
    int* wk = ws.win_k + (size_t)item * b.win_cap;
    for (int c = tid; c < W; c += blockDim.x) wk[c] = ck[cidx[c]];
    if (tid == 0) {
      iiv[I_NWIN] = W; rec[PWB_REC_NWIN] = W;
      if (nclus > W) { iiv[I_STATUS] |= PWB_ST_WIN_CAP; rec[PWB_REC_STATUS] = iiv[I_STATUS]; }
    }
  }
}

// =================================================================================================
// k_windows: window_analysis (utilities.py:1191-1361), one warp per (item, window)
// =================================================================================================
constexpr int WIN_WARPS = 4;
#ifndef WIN_TIGHT_RHO
#define WIN_TIGHT_RHO 0.5
#endif
// doubles of shared memory per warp: double4[NA] + r[NA] + class radii/starts + near list + z-stage state +
// coarse grid values and candidate list (100 + 76), even
// per-atom part: double4[NA] + r[NA] + near-list indices (ushort[NA]); fixed part: class radii / starts, z-stage
// solver state, coarse grid + candidate list.  Both even, the atom part a multiple of 4 (double4 alignment).
#define WIN_ATOM_DOUBLES(NA) ((((size_t)(NA) * 5 + (((size_t)(NA) + 3) & ~(size_t)3) / 4) + 3) & ~(size_t)3)
#define WIN_FIXED_DOUBLES ((size_t)((PWB_MAX_TYPES + 32 + LB_STORAGE(1) + 176 + 3) & ~3))
#define WIN_WARP_DOUBLES(NA) (WIN_ATOM_DOUBLES(NA) + WIN_FIXED_DOUBLES)

struct DevZObj {   // optimise_z, utilities.py:1174-1188
  WarpAtoms A; double lo;
  __device__ void fg(const double* x, double& f, double* g) {
    double h = lb_fd_step(x[0], 1, lo, 0.0);
    double px[2] = {0.0, 0.0}, py[2] = {0.0, 0.0}, pz[2] = {x[0], x[0] + h}, m[2];
    warp_min_clear<2>(A, px, py, pz, m);
    f = m[0] * 2.0;
    g[0] = (m[1] * 2.0 - f) / ((x[0] + h) - x[0]);
  }
};
struct DevXYObj {  // optimise_xy, utilities.py:1164-1171
  WarpAtoms A; NearList N; double z;
  // tight list: the <= 32 atoms that can attain the minimum within sqrt(trho2) of (tx, ty, z), one per
  // lane, in registers.  Nelder-Mead rarely leaves that disc, so its objective is load-free.
  double tx, ty, trho2, ax, ay, az, ar; bool have;
  __device__ double f2(double x, double y) {
    const double ex = x - tx, ey = y - ty;
    if (ex * ex + ey * ey <= trho2) {
      double v = PWB_INF;
      if (have) {
        const double dx = ax - x, dy = ay - y, dz = az - z;
        v = sqrt(dx * dx + dy * dy + dz * dz) - ar;
      }
      return -(warp_min(v) * 2.0);
    }
    double m[1];
    if (near_valid(N, x, y, z)) warp_min_clear_near<1>(A, N, &x, &y, &z, m);
    else warp_min_clear<1>(A, &x, &y, &z, m);
    return -(m[0] * 2.0);
  }
};

PWB_MATHFN double angle_between(double x0, double x1, double x2, double y0, double y1, double y2) {
  double first = fabs(x0 * y0 + x1 * y1 + x2 * y2) /
                 (sqrt(x0 * x0 + x1 * x1 + x2 * x2) * sqrt(y0 * y0 + y1 * y1 + y2 * y2));
  return acos(first);
}

#ifndef WIN_MINB
#define WIN_MINB 6
#endif
#ifndef WIN_SPLIT
#define WIN_SPLIT 2                  // 2: window search in two kernels, cut behind the grid search (4: behind the z
                                     // stage; 3: three kernels; 0: one)
#endif
// one (item, window) on one warp; `smw` is the warp's private shared memory (WIN_WARP_DOUBLES).
// PARTS is a bit mask of what this instantiation runs: 1 = rescan + rotation + z stage, 2 = 20 x 20 grid search,
// 4 = tight list + Nelder-Mead + output; 7 = the whole window in one kernel.  With WIN_SPLIT the function is cut into
// two or three kernels so that the warps resident on an SM walk the same part of this large function (instruction
// cache): a kernel that does not finish the window leaves (x0, y0, wz, new_z) in the window's record slot; the next
// one loads and rotates the atoms again (same expressions, same values), rebuilds d0 and the near list and goes on.
template <int PARTS>
__device__ __forceinline__ void window_slot(const pwb_batch& b, const Ws& ws, const int item, const int win,
                                            const int lane, double* smw, double* atoms_w) {
  int* iiv = ws.item_i + (size_t)item * ITEM_I;
  double* idv = ws.item_d + (size_t)item * ITEM_D;
  double* rec = b.records + (size_t)item * b.rec_stride;
  double* out = rec + PWB_REC_WIN0 + 4 * win;
  const int n = iiv[I_N], off = b.item_off[item], NA = b.max_atoms;
  double4* p4 = (double4*)atoms_w;           // x, y, z, |x|^2 (shifted, later rotated)
  double* r = (double*)(p4 + NA);
  unsigned short* nidx = (unsigned short*)(r + NA);                     // [NA] near-list atom indices
  double* cr = smw;                          // [PWB_MAX_TYPES]
  int* cls = (int*)(cr + PWB_MAX_TYPES);     // [PWB_MAX_TYPES+1]
  int* ncs = cls + PWB_MAX_TYPES + 1;        // [PWB_MAX_TYPES+1] near-list class starts
  double* hist = cr + PWB_MAX_TYPES + 32;                               // [LB_STORAGE(1)] z-stage L-BFGS state
  double h_x0 = 0, h_y0 = 0, h_wz = 0, h_newz = 0;
  if (!(PARTS & 1)) {
    h_x0 = out[0]; h_y0 = out[1]; h_wz = out[2]; h_newz = out[3];
    if (h_x0 != h_x0) return;                                           // dropped in phase 1 (NaNs stay)
    __syncwarp();
  }
  const double shx = idv[D_SHIFT], shy = idv[D_SHIFT + 1], shz = idv[D_SHIFT + 2];
  for (int a = lane; a < n; a += 32) {
    const double* q = ws.atoms + (size_t)(off + a) * 4;
    const double X = q[0] - shx, Y = q[1] - shy, Z = q[2] - shz;
    p4[a] = make_double4(X, Y, Z, X * X + Y * Y + Z * Z); r[a] = q[3];
  }
  if (lane <= PWB_MAX_TYPES) cls[lane] = iiv[I_CLS0 + lane];
  if (lane < PWB_MAX_TYPES) cr[lane] = lane < b.n_types ? b.type_vdw[lane] : 0.0;
  __syncwarp();
  Bucketed B{p4, cls, cr, b.n_types};
  const int P = iiv[I_P];
  const double R = idv[D_RADIUS];
  // ---- rescan of the arg-max ray with increment2 (utilities.py:1221-1227)
  double vx, vy, vz;
  spiral_point(ws.win_k[(size_t)item * b.win_cap + win], P, R, vx, vy, vz);
  const double nr = sqrt(vx * vx + vy * vy + vz * vz);
  const int chunks = (int)py_floordiv(nr, b.increment2);
  const double cxs = vx / chunks, cys = vy / chunks, czs = vz / chunks;
  // same bracketing argument as in k_rays: two candidate steps per atom, lanes stride over atoms
  ValIdx bj; bj.v = PWB_INF; bj.i = 0x7fffffff;
  double new_z = h_newz;
  if (PARTS & 1) {
  {
    const double cc = cxs * cxs + cys * cys + czs * czs, icc = 1.0 / cc;
    const double jmax = (double)(chunks > 0 ? chunks - 1 : 0);
    for (int a = lane; a < n; a += 32) {
      const double4 q = p4[a];
      const double g = fma(q.x, cxs, fma(q.y, cys, q.z * czs));
      const double j0 = floor(fmin(fmax(g * icc, 0.0), jmax)), j1 = j0 + 1.0;
      const double g2 = -2.0 * g, wv = q.w;
      const double d0 = fma(j0, fma(j0, cc, g2), wv), d1 = fma(j1, fma(j1, cc, g2), wv);
      ValIdx c0; c0.v = sqrt(fmax(d0, 0.0)) - r[a]; c0.i = (int)j0;
      ValIdx c1; c1.v = sqrt(fmax(d1, 0.0)) - r[a]; c1.i = (int)j1;
      bj = vi_min(bj, vi_min(c0, c1));
    }
  }
  bj = warp_vi_min(bj);
  const int bad = !(chunks > 0 && bj.v > 0.0);
  if (bad) {
    if (lane == 0) {
      out[0] = nan(""); out[1] = out[2] = out[3] = nan("");
      atomicOr(&iiv[I_STATUS], PWB_ST_WIN_DROPPED);
    }
    return;
  }
#if defined(WIN_STOP) && WIN_STOP == 1      // (phase timing experiment, profiles/r2_opt_pore_variants.md)
  if (lane == 0) out[0] = bj.v;
  return;
#endif
  const double qx = cxs * bj.i, qy = cys * bj.i, qz = czs * bj.i;
  new_z = sqrt(qx * qx + qy * qy + qz * qz);
  }
  // ---- rotation angles, octant table (utilities.py:1235-1259)
  double a1 = angle_between(vx, vy, 0.0, 1.0, 0.0, 0.0);
  double a2 = angle_between(vx, vy, vz, 0.0, 0.0, 1.0);
  const double PI = 3.141592653589793;
  const bool gx = vx >= 0, gy = vy >= 0, gz = vz >= 0;
  if (gx && gy && gz) { a1 = -a1; a2 = -a2; }
  if (!gx && gy && gz) { a1 = PI * 2 + a1; }
  if (gx && !gy && gz) { a2 = -a2; }
  if (!gx && !gy && gz) { a1 = PI * 2 - a1; }
  if (gx && gy && !gz) { a1 = -a1; a2 = PI + a2; }
  if (!gx && gy && !gz) { a2 = PI - a2; }
  if (gx && !gy && !gz) { a2 = a2 + PI; }
  if (!gx && !gy && !gz) { a1 = -a1; a2 = PI - a2; }
  double s1, c1, s2, c2;
  sincos(a1, &s1, &c1);
  sincos(a2, &s2, &c2);
  for (int a = lane; a < n; a += 32) {
    double X0 = p4[a].x, Y0 = p4[a].y, Z0 = p4[a].z;
    double X1 = c1 * X0 + (-s1) * Y0, Y1 = s1 * X0 + c1 * Y0, Z1 = Z0;         // Rz(a1)
    double X2 = c2 * X1 + s2 * Z1, Y2 = Y1, Z2 = (-s2) * X1 + c2 * Z1;          // Ry(a2)
    const double Z3 = Z2 - new_z;
    p4[a] = make_double4(X2, Y2, Z3, X2 * X2 + Y2 * Y2 + Z3 * Z3);
  }
  __syncwarp();
  WarpAtoms A{p4, r, 1, n, lane, cls, b.n_types};
  double m0[1], zero = 0.0;
  warp_min_clear<1>(A, &zero, &zero, &zero, m0);
  const double d0 = m0[0] * 2.0;
  // ---- z stage: L-BFGS-B from 0 with lower bound -new_z (utilities.py:1296-1305)
  double wz = h_wz;
  if (PARTS & 1) {
    double zc[1] = {0.0}, zlo[1] = {-new_z}, zhi[1] = {0.0}, fz; int nitz;
    DevZObj zo{A, -new_z};
    lbfgsb_minimize<1, DevZObj, 1>(zo, zc, zlo, zhi, nullptr, 2.220446049250313e-09, 1e-5, 20, 15000, &fz, &nitz, hist);
    wz = zc[0];
  }
#if defined(WIN_STOP) && WIN_STOP == 2
  if (lane == 0) out[0] = wz;
  return;
#endif
  // ---- xy stage: 20x20 brute grid (first arg-min, row-major) + Nelder-Mead (utilities.py:1307-1316)
  const double glo = -d0 / 2, ghi = d0 / 2;
  const double gstep = (ghi - glo) / 19.0;
  // only atoms that can touch a sphere centred within the grid square (+ slack for the simplex) matter
  NearList NL; NL.idx = nidx; NL.cs = ncs;
  warp_build_near(A, NL, 0.0, 0.0, wz, fabs(d0) * 0.7071067811865476 + 0.25);
  // Exact branch and bound over the 400 grid points.  The clearance is 1-Lipschitz, so after the coarse
  // 10 x 10 sub-lattice (even ix, iy) has been evaluated, a remaining point can only reach the best
  // coarse value c* if its even-even neighbour (at most gstep*sqrt(2) away) has clearance >= c* minus
  // that distance.  Only those candidates (the cap of the clearance cone, typically 20-40 points) are
  // evaluated; every evaluated value is computed exactly as before and the arg-min rule (value, then
  // row-major index) does not depend on the evaluation order.
  double x0 = h_x0, y0 = h_y0;
  if (PARTS & 2) {
  double* cgrid = hist + LB_STORAGE(1);                        // [100] coarse clearances
  unsigned short* clist = (unsigned short*)(cgrid + 100);      // [300] candidate grid indices
  ValIdx gb; gb.v = PWB_INF; gb.i = 0x7fffffff;
  int nwork = 128;                                             // slots 0..127: coarse points (100 used)
  bool built = false;
  for (int base = 0; base < nwork || !built; base += 64) {
    if (base >= 128 && !built) {
      built = true;
      __syncwarp();
      const double cstar = -(warp_vi_min(gb).v) * 0.5;
      int pos = 0;
      for (int g0 = 0; g0 < 400; g0 += 32) {
        const int g = g0 + lane;
        bool cand = false;
        if (g < 400) {
          const int ix = g / 20, iy = g - ix * 20;
          const int odd = (ix & 1) + (iy & 1);
          if (odd) {
            const double reach = fabs(gstep) * (odd == 2 ? 1.4142135623730951 : 1.0);
            cand = cgrid[(ix >> 1) * 10 + (iy >> 1)] + reach * (1.0 + 1e-12) + 1e-9 >= cstar;
          }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, cand);
        if (cand) clist[pos + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)g;
        pos += __popc(mask);
      }
      __syncwarp();
      nwork = 128 + pos;
      if (base >= nwork) break;
    }
    double gxp[2], gyp[2], gzp[2] = {wz, wz}, gv[2];
    int gi[2];
#pragma unroll
    for (int v = 0; v < 2; v++) {
      const int t = base + lane + 32 * v;
      if (t < 128) { const int c = min(t, 99), cx = c / 10; gi[v] = cx * 40 + (c - cx * 10) * 2; }
      else gi[v] = clist[min(t, nwork - 1) - 128];
      const int ix = gi[v] / 20, iy = gi[v] - ix * 20;
      gxp[v] = __dadd_rn(__dmul_rn((double)ix, gstep), glo); gyp[v] = __dadd_rn(__dmul_rn((double)iy, gstep), glo);
    }
    min_clearance_k_near<2>(B, NL, gxp, gyp, gzp, gv);
#pragma unroll
    for (int v = 0; v < 2; v++) {
      const int t = base + lane + 32 * v;
      if (t < 100) cgrid[t] = gv[v];
      ValIdx c; c.v = -(gv[v] * 2.0); c.i = gi[v];
      gb = vi_min(gb, c);
    }
  }
  gb = warp_vi_min(gb);
  const int ix = gb.i / 20, iy = gb.i - ix * 20;
  x0 = __dadd_rn(__dmul_rn((double)ix, gstep), glo); y0 = __dadd_rn(__dmul_rn((double)iy, gstep), glo);
  }
  if (!(PARTS & 4)) {                                          // hand over to the next kernel through the record slot
    if (lane == 0) { out[0] = x0; out[1] = y0; out[2] = wz; out[3] = new_z; }
    return;
  }
#if defined(WIN_STOP) && WIN_STOP == 3
  if (lane == 0) out[0] = x0 + y0;
  return;
#endif
  DevXYObj xo{A, NL, wz, x0, y0, -1.0, 0.0, 0.0, 0.0, 0.0, false};
  {
    // same 1-Lipschitz argument as warp_build_near, radius WIN_TIGHT_RHO around the grid optimum
    unsigned short* tidx = (unsigned short*)(ncs + PWB_MAX_TYPES + 2);   // [32] scratch
    double mc[1];
    warp_min_clear_near<1>(A, NL, &x0, &y0, &wz, mc);
    const double thr = mc[0] + 2.0 * WIN_TIGHT_RHO + 1e-9;
    const int nn = NL.cs[A.ncls];
    int pos = 0;
    for (int base = 0; base < nn && pos <= 32; base += 32) {
      const int t = base + lane;
      bool keep = false; int i = 0;
      if (t < nn) {
        i = NL.idx[t];
        const double4 q = p4[i];
        const double dx = q.x - x0, dy = q.y - y0, dz = q.z - wz;
        keep = sqrt(dx * dx + dy * dy + dz * dz) - r[i] <= thr;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      const int at = pos + __popc(mask & ((1u << lane) - 1u));
      if (keep && at < 32) tidx[at] = (unsigned short)i;
      pos += __popc(mask);
    }
    __syncwarp();
    if (pos <= 32) {
      const double re = WIN_TIGHT_RHO - 1e-6;
      xo.trho2 = re * re;
      if (lane < pos) {
        const int i = tidx[lane];
        const double4 q = p4[i];
        xo.ax = q.x; xo.ay = q.y; xo.az = q.z; xo.ar = r[i]; xo.have = true;
      }
    }
  }
  double xy[2], fxy; int nfev;
  nelder_mead2(xo, x0, y0, 1e-4, 1e-4, 400, 400, xy, &fxy, &nfev);
  double mf[1];
  warp_min_clear<1>(A, &xy[0], &xy[1], &wz, mf);
  const double diam = mf[0] * 2.0;
  // ---- back to the molecule frame (utilities.py:1337-1360) + initial_com (:1529-1535)
  double wx = xy[0], wy = xy[1], wzz = wz + new_z;
  double s2m, c2m, s1m, c1m;
  sincos(-a2, &s2m, &c2m);
  sincos(-a1, &s1m, &c1m);
  double X1 = c2m * wx + s2m * wzz, Y1 = wy, Z1 = (-s2m) * wx + c2m * wzz;
  double X2 = c1m * X1 + (-s1m) * Y1, Y2 = s1m * X1 + c1m * Y1, Z2 = Z1;
  if (lane == 0) {
    out[0] = diam;
    out[1] = X2 + idv[D_ICOM]; out[2] = Y2 + idv[D_ICOM + 1]; out[3] = Z2 + idv[D_ICOM + 2];
    if (diam < 0) atomicOr(&iiv[I_STATUS], PWB_ST_WIN_NEGATIVE);
  }
}

// Static (item, window) -> warp map: a work queue like k_opt_pore's was measured 45 % SLOWER here (warps
// of a CTA drift into different phases of this 200 KB kernel and miss in the instruction cache).
// (Measured and rejected: running the Nelder-Mead searches of a CTA's four windows together on one warp, eight lanes
// each, the other warps waiting at a barrier -- bit-identical records, 2.25 instead of 1.63 ms per 10 000 frames.
// The kernel is latency bound with idle issue slots; sharing the scalar simplex bookkeeping saves instructions nobody
// was waiting for and lengthens every evaluation: four atoms per lane in series, group shuffles, divergent groups.)
template <bool BIG, int PARTS>
__global__ void __launch_bounds__(WIN_WARPS * 32, WIN_MINB)
k_windows(pwb_batch b, Ws ws) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * (blockDim.x >> 5) + w;        // warps per CTA chosen at launch (<= WIN_WARPS)
  const int item = gw / b.win_cap, win = gw - item * b.win_cap;
  if (item >= b.n_items) return;
  if (win >= ws.item_i[(size_t)item * ITEM_I + I_NWIN]) return;
  // per-warp shared memory: [atoms | fixed]; huge molecules (BIG) keep the atom part in the warp's global scratch slot
  if (BIG) window_slot<PARTS>(b, ws, item, win, lane, sm + (size_t)w * WIN_FIXED_DOUBLES,
                       ws.atom_scratch + (size_t)item * ws.scratch_item + (size_t)win * WIN_ATOM_DOUBLES(b.max_atoms));
  else {
    double* smw = sm + (size_t)w * WIN_WARP_DOUBLES(b.max_atoms);
    window_slot<PARTS>(b, ws, item, win, lane, smw + WIN_ATOM_DOUBLES(b.max_atoms), smw);
  }
}

// final status word (window kernels OR flags into item_i after the record status was written)
__global__ void k_finish(pwb_batch b, Ws ws) {
  int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= b.n_items) return;
  b.records[(size_t)item * b.rec_stride + PWB_REC_STATUS] = ws.item_i[(size_t)item * ITEM_I + I_STATUS];
}

// bounding-box bound on the ray count (sizes ray_cap); one block, result in out[0]
__global__ void k_ray_capacity(pwb_batch b, int* out) {
  __shared__ int smax;
  if (threadIdx.x == 0) smax = 0;
  __syncthreads();
  double rmax = 0;
  for (int t = 0; t < b.n_types; t++) rmax = fmax(rmax, b.type_vdw[t]);
  int local = 0;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < b.n_items; item += gridDim.x * blockDim.x) {
    int off = b.item_off[item], n = b.item_off[item + 1] - off;
    double lo[3] = {PWB_INF, PWB_INF, PWB_INF}, hi[3] = {-PWB_INF, -PWB_INF, -PWB_INF};
    for (int a = 0; a < n; a++)
      for (int d = 0; d < 3; d++) { double v = b.coords[(size_t)(off + a) * 3 + d]; lo[d] = fmin(lo[d], v); hi[d] = fmax(hi[d], v); }
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    double D = sqrt(dx * dx + dy * dy + dz * dz) + 2 * rmax;
    local = max(local, max(n_sampling_points(D / 2, b.adjust), n_sampling_points(D, b.adjust)));
  }
  atomicMax(&smax, local);
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(out, smax);
}

// FMA-pipe microbenchmark (roofline denominator for the FP32/FP64 pipes)
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b) {
  T v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v0 = v0 * a + b; v1 = v1 * a + b; v2 = v2 * a + b; v3 = v3 * a + b;
      v4 = v4 * a + b; v5 = v5 * a + b; v6 = v6 * a + b; v7 = v7 * a + b;
    }
  }
  T s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
  if (s == (T)123456789) out[0] = s;
}

}  // namespace pwb

// =================================================================================================
// C ABI
// =================================================================================================
using namespace pwb;

constexpr int PWB_MAX_CHUNKS = 8;
struct pwb_ctx {
  Ctx c; int* d_tmp; int max_smem, num_sms;
  int chunks;                                   // sub-batches pipelined over internal streams
  int opt_no_list;                              // pwb_ctx_set_option(PWB_OPTION_OPT_NO_LIST)
  cudaStream_t streams[PWB_MAX_CHUNKS];
  cudaEvent_t fork, join[PWB_MAX_CHUNKS];
  cudaStream_t copy_stream;                    // host -> device copies of the *_host entry points
  cudaEvent_t h2d[PWB_MAX_CHUNKS];
};

static thread_local char g_err[512] = "";
static int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, e_); } while (0)

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

template <class K>
static int set_smem(K kern, size_t bytes, int max_smem) {
  if ((int)bytes > max_smem) { snprintf(g_err, sizeof g_err, "shared memory %zu > device limit %d (too many atoms)", bytes, max_smem); return 1; }
  if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

extern "C" {

void pwb_set_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); }
int pwb_abi_version(void) { return PWB_ABI_VERSION; }
const char* pwb_last_error(void) { return g_err; }

// CTAs per item the ray-tiled kernels may use (see tile_ctas): bound used for sizing
static int tile_ctas_bound(int n_items, int ray_cap, int tile, int num_sms) {
  const int tiles = (ray_cap + tile - 1) / tile;
  int want = (2 * num_sms) / (n_items > 0 ? n_items : 1);
  if (want > tiles) want = tiles;
  if (want > 128) want = 128;
  return want < 1 ? 1 : want;
}
static size_t rays_smem(int max_atoms) {
  size_t sm = sizeof(double) * (RAY_FIXED_DOUBLES + RAY_ATOM_DOUBLES((size_t)max_atoms));
#ifdef RAYS_FP32_FILTER
  sm += sizeof(float) * 4 * (size_t)max_atoms + 16;
#endif
  return sm;
}
// huge molecules: the per-atom arrays of k_prepare / k_rays / k_windows go to global scratch
static bool big_atoms(int max_atoms) { return rays_smem(max_atoms) > PWB_SMEM_LIMIT; }
static size_t scratch_item_doubles(int n_items, int max_atoms, int ray_cap, int win_cap) {
  const size_t y = (size_t)tile_ctas_bound(n_items, ray_cap, RAY_TILE, PWB_MAX_SMS);
  size_t a = y * RAY_ATOM_DOUBLES((size_t)max_atoms), p = y * PREP_ATOM_DOUBLES((size_t)max_atoms);
  size_t w = (size_t)win_cap * WIN_ATOM_DOUBLES((size_t)max_atoms);
  size_t m = a > p ? a : p;
  m = m > w ? m : w;
  return (m + 3) & ~(size_t)3;
}

int pwb_workspace_layout(int F, int total_atoms, int max_atoms, int ray_cap, int win_cap, pwb_ws_layout* L) {
  size_t o = 0;
  L->atoms = o; o = align_up(o + (size_t)total_atoms * 4 * sizeof(double));
  L->atom_idx = o; o = align_up(o + (size_t)total_atoms * sizeof(int));
  L->item_d = o; o = align_up(o + (size_t)F * ITEM_D * sizeof(double));
  L->item_i = o; o = align_up(o + (size_t)F * ITEM_I * sizeof(int));
  L->clear_k = o; o = align_up(o + (size_t)F * ray_cap * sizeof(int));
  L->clear_pos = o; o = align_up(o + (size_t)F * ray_cap * sizeof(int));
  L->clear_a2 = o; o = align_up(o + (size_t)F * ray_cap * sizeof(double));
  L->labels = o; o = align_up(o + (size_t)F * ray_cap * sizeof(int));
  L->win_k = o; o = align_up(o + (size_t)F * win_cap * sizeof(int));
  const size_t big = ray_cap > KNN_SMEM_CAP ? (size_t)F * ray_cap : 0;   // large-ray-count scratch
  L->big_pts = o; o = align_up(o + big * 4 * sizeof(double));
  L->big_i = o; o = align_up(o + big * 2 * sizeof(int));
  L->atom_scratch = o;
  if (big_atoms(max_atoms)) o = align_up(o + (size_t)F * scratch_item_doubles(F, max_atoms, ray_cap, win_cap) * sizeof(double));
  L->opt_hist = o; o = align_up(o + (size_t)F * LB_HISTORY(3) * sizeof(double));
  L->total = o;
  return 0;
}
size_t pwb_workspace_bytes(int F, int total_atoms, int max_atoms, int ray_cap, int win_cap) {
  pwb_ws_layout L;
  pwb_workspace_layout(F, total_atoms, max_atoms, ray_cap, win_cap, &L);
  return L.total;
}

int pwb_ctx_create(pwb_ctx** out, int device, int ray_table_cap) {
  CK(cudaSetDevice(device));
  pwb_ctx* c = new pwb_ctx();
  c->c.device = device;
  c->c.table_cap = ray_table_cap;
  c->c.launches = 0;
  CK(cudaMalloc(&c->c.knn_unit, sizeof(double) * ray_table_cap));
  CK(cudaMalloc(&c->c.knn_need, sizeof(int) * ray_table_cap));
  CK(cudaMalloc(&c->d_tmp, sizeof(int) * 4));
  CK(cudaMemset(c->c.knn_unit, 0, sizeof(double) * ray_table_cap));
  CK(cudaMemset(c->c.knn_need, 0, sizeof(int) * ray_table_cap));
  CK(cudaDeviceGetAttribute(&c->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  c->chunks = 1;
  c->opt_no_list = 0;
  for (int g = 0; g < PWB_MAX_CHUNKS; g++) {
    CK(cudaStreamCreateWithFlags(&c->streams[g], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->join[g], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->h2d[g], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  *out = c;
  return 0;
}
int pwb_ctx_set_option(pwb_ctx* c, int option, int value) {
  if (!c) { snprintf(g_err, sizeof g_err, "null context"); return 1; }
  if (option == PWB_OPTION_OPT_NO_LIST) { c->opt_no_list = value != 0; return 0; }
  snprintf(g_err, sizeof g_err, "unknown option %d", option);
  return 1;
}
int pwb_ctx_set_chunks(pwb_ctx* c, int chunks) {
  c->chunks = chunks < 1 ? 1 : (chunks > PWB_MAX_CHUNKS ? PWB_MAX_CHUNKS : chunks);
  return 0;
}
int pwb_ctx_destroy(pwb_ctx* c) {
  if (!c) return 0;
  cudaFree(c->c.knn_unit); cudaFree(c->c.knn_need); cudaFree(c->d_tmp);
  for (int g = 0; g < PWB_MAX_CHUNKS; g++) { cudaStreamDestroy(c->streams[g]); cudaEventDestroy(c->join[g]); cudaEventDestroy(c->h2d[g]); }
  cudaEventDestroy(c->fork);
  cudaStreamDestroy(c->copy_stream);
  delete c;
  return 0;
}
long long pwb_launch_count(pwb_ctx* c) { return c->c.launches; }

static int resolve(const pwb_batch* b, Ws* w) {
  if (!b->workspace) { snprintf(g_err, sizeof g_err, "workspace is null"); return 1; }
  pwb_ws_layout L;
  pwb_workspace_layout(b->n_items, b->total_atoms, b->max_atoms, b->ray_cap, b->win_cap, &L);
  if (b->workspace_bytes < L.total) { snprintf(g_err, sizeof g_err, "workspace too small: %zu < %zu", b->workspace_bytes, L.total); return 1; }
  char* p = (char*)b->workspace;
  w->atoms = (double*)(p + L.atoms); w->atom_idx = (int*)(p + L.atom_idx);
  w->item_d = (double*)(p + L.item_d); w->item_i = (int*)(p + L.item_i);
  w->clear_k = (int*)(p + L.clear_k); w->clear_pos = (int*)(p + L.clear_pos);
  w->clear_a2 = (double*)(p + L.clear_a2); w->labels = (int*)(p + L.labels); w->win_k = (int*)(p + L.win_k);
  w->big_pts = (double*)(p + L.big_pts); w->big_i = (int*)(p + L.big_i);
  const bool big = big_atoms(b->max_atoms);
  w->atom_scratch = big ? (double*)(p + L.atom_scratch) : nullptr;
  w->scratch_item = big ? (long long)scratch_item_doubles(b->n_items, b->max_atoms, b->ray_cap, b->win_cap) : 0;
  w->opt_hist = (double*)(p + L.opt_hist);
  return 0;
}

static int check(const pwb_batch* b) {
  if (b->n_items <= 0) { snprintf(g_err, sizeof g_err, "n_items must be positive"); return 1; }
  if (b->ray_tile_end < 0 || b->ray_tile_begin < 0 || (b->ray_tile_end > 0 && b->ray_tile_begin > b->ray_tile_end)) {
    snprintf(g_err, sizeof g_err, "bad ray tile range"); return 1;
  }
  if (b->n_types < 1 || b->n_types > PWB_MAX_TYPES) { snprintf(g_err, sizeof g_err, "n_types out of range"); return 1; }
  if (b->rec_stride < PWB_REC_WIN0 + 4 * b->win_cap) { snprintf(g_err, sizeof g_err, "rec_stride too small"); return 1; }
  if (b->ray_cap > (1 << 22)) { snprintf(g_err, sizeof g_err, "ray_cap > 4M rays not supported"); return 1; }
  if (b->max_atoms < 1 || b->max_atoms > 65535) {   // 16-bit atom indices: near lists (k_windows), survivor lists (k_opt_pore)
    snprintf(g_err, sizeof g_err, "max_atoms %d outside 1..65535 atoms per molecule", b->max_atoms); return 1;
  }
  return 0;
}

// CTAs per item for the ray-tiled kernels: one for a batch that already fills the GPU, up to the number of ray
// tiles when a few huge items (config 5: ~1e5 rays on one molecule) would otherwise run on a handful of SMs
static int tile_ctas(const pwb_ctx* c, const pwb_batch* b, int tile) {
  return tile_ctas_bound(b->n_items, b->ray_cap, tile, c->num_sms < PWB_MAX_SMS ? c->num_sms : PWB_MAX_SMS);
}
static int run_prepare(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s, int do_avg) {
  size_t sm = sizeof(double) * (PREP_FIXED_DOUBLES + (w.atom_scratch ? 0 : PREP_ATOM_DOUBLES((size_t)b->max_atoms)));
  auto kern = w.atom_scratch ? k_prepare<true> : k_prepare<false>;
  if (set_smem(kern, sm, c->max_smem)) return 1;
  const dim3 grid((unsigned)b->n_items, do_avg ? (unsigned)tile_ctas(c, b, AVG_TILE) : 1u);
  // per-item "tiles done" counters of the multi-CTA paths (k_prepare average diameter, k_rays); self-resetting
  // afterwards, only ever touched when an item is shared by several CTAs
  if (tile_ctas(c, b, RAY_TILE) > 1)
    CK(cudaMemset2DAsync(w.item_i + I_TILES_PREP, sizeof(int) * ITEM_I, 0, 2 * sizeof(int), (size_t)b->n_items, s));
  kern<<<grid, PREP_THREADS, sm, s>>>(*b, w, do_avg);
  c->c.launches++;
  CK(cudaGetLastError());
  return 0;
}
static int run_opt(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s) {
  constexpr bool SA = OPT_ATOMS_SMEM != 0;
  constexpr int PPW = 32 / OPT_G;
  // fewer warps per CTA for very large molecules so the per-problem atom tiles still fit shared memory
  const size_t per_warp = sizeof(double) * OPT_PROB_DOUBLES((size_t)b->max_atoms, SA) * PPW;
  int wpb = OPT_WARPS;
  while (wpb > 1 && per_warp * wpb > (size_t)c->max_smem) wpb >>= 1;
  size_t sm = per_warp * wpb;
  auto kern = b->opt_bounds ? k_opt_pore<OPT_G, true, SA> : k_opt_pore<OPT_G, false, SA>;
  if (set_smem(kern, sm, c->max_smem)) return 1;
#ifdef OPT_LIST_STATS
  CK(cudaMemsetAsync(w.item_i + I_QUEUE, 0, 5 * sizeof(int), s));      // (slots 27..30 of the first item are free)
#else
  CK(cudaMemsetAsync(w.item_i + I_QUEUE, 0, sizeof(int), s));
#endif
  int ctas = (b->n_items + wpb * PPW - 1) / (wpb * PPW);
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, sm));
  const int resident = c->num_sms * (per_sm > 0 ? per_sm : 1);
  if (ctas > resident) ctas = resident;
  kern<<<ctas, wpb * 32, sm, s>>>(*b, w);
  c->c.launches++;
  CK(cudaGetLastError());
#ifdef OPT_LIST_STATS
  {
    int st[5];
    CK(cudaMemcpyAsync(st, w.item_i + I_QUEUE, sizeof st, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    fprintf(stderr, "k_opt_pore stencils: %d over the list (mean length %.1f), %d over all atoms, %d list builds\n",
            st[1], st[1] ? (double)st[4] / st[1] : 0.0, st[2], st[3]);
  }
#endif
  return 0;
}
static int run_rays(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s) {
  size_t sm = w.atom_scratch ? sizeof(double) * RAY_FIXED_DOUBLES : rays_smem(b->max_atoms);
#ifdef RAYS_FP32_FILTER
  if (w.atom_scratch) sm += sizeof(float) * 4 * (size_t)b->max_atoms + 16;
#endif
  auto kern = w.atom_scratch ? k_rays<true> : k_rays<false>;
  if (set_smem(kern, sm, c->max_smem)) return 1;
  const dim3 grid((unsigned)b->n_items, (unsigned)tile_ctas(c, b, RAY_TILE));
  kern<<<grid, RAY_THREADS, sm, s>>>(*b, w, c->c);
  c->c.launches++;
  CK(cudaGetLastError());
  return 0;
}
static int run_cluster(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s) {
  const int kcap = c->c.table_cap < KNN_SMEM_CAP + 1 ? c->c.table_cap : KNN_SMEM_CAP + 1;
  size_t smk = sizeof(double) * 3 * (size_t)kcap;
  if (set_smem(k_knn_fill, smk, c->max_smem)) return 1;
  k_knn_fill<<<kcap - 2, KNN_THREADS, smk, s>>>(c->c, kcap);
  c->c.launches++;
  // pass 0: small budget, high occupancy; pass 1: whatever the device allows, deferred items only
  unsigned big = (unsigned)c->max_smem;
  if (set_smem(k_cluster, big, c->max_smem)) return 1;
  k_cluster<<<b->n_items, CL_THREADS, CL_SMALL_SMEM, s>>>(*b, w, c->c, 0, CL_SMALL_SMEM);
  k_cluster<<<b->n_items, CL_THREADS, big, s>>>(*b, w, c->c, 1, big);
  c->c.launches += 2;
  if (b->ray_cap > KNN_SMEM_CAP) {                 // items whose ray count exceeds the shared-memory paths
    k_big_points<<<dim3(64, b->n_items), 256, 0, s>>>(*b, w);
    k_knn_big<<<dim3(BIG_SLICES, b->n_items), 256, 0, s>>>(*b, w);
    const int cb_blocks = (b->ray_cap + CLB_THREADS - 1) / CLB_THREADS;
    for (int ph = 0; ph < 7; ph++) {
      const int gx = (ph == 3 || ph == 6) ? 1 : cb_blocks;
      k_cluster_big<<<dim3(gx, b->n_items), CLB_THREADS, 0, s>>>(*b, w, ph);
    }
    c->c.launches += 9;
  }
  CK(cudaGetLastError());
  return 0;
}
static int run_windows(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s) {
  const size_t per_warp = sizeof(double) * (w.atom_scratch ? WIN_FIXED_DOUBLES : WIN_WARP_DOUBLES((size_t)b->max_atoms));
  int wpb = WIN_WARPS;
  while (wpb > 1 && per_warp * wpb + 64 > (size_t)c->max_smem) wpb >>= 1;
  size_t sm = per_warp * wpb + 64;
  long long warps = (long long)b->n_items * b->win_cap;
  const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
#if WIN_SPLIT == 2 || WIN_SPLIT == 4
  constexpr int P1 = WIN_SPLIT == 2 ? 3 : 1, P2 = 7 - P1;       // cut behind the grid search / behind the z stage
  auto k1 = w.atom_scratch ? k_windows<true, P1> : k_windows<false, P1>;
  auto k2 = w.atom_scratch ? k_windows<true, P2> : k_windows<false, P2>;
  if (set_smem(k1, sm, c->max_smem) || set_smem(k2, sm, c->max_smem)) return 1;
  k1<<<grid, wpb * 32, sm, s>>>(*b, w);
  k2<<<grid, wpb * 32, sm, s>>>(*b, w);
  c->c.launches++;
#elif WIN_SPLIT == 3
  auto k1 = w.atom_scratch ? k_windows<true, 1> : k_windows<false, 1>;
  auto k2 = w.atom_scratch ? k_windows<true, 2> : k_windows<false, 2>;
  auto k3 = w.atom_scratch ? k_windows<true, 4> : k_windows<false, 4>;
  if (set_smem(k1, sm, c->max_smem) || set_smem(k2, sm, c->max_smem) || set_smem(k3, sm, c->max_smem)) return 1;
  k1<<<grid, wpb * 32, sm, s>>>(*b, w);
  k2<<<grid, wpb * 32, sm, s>>>(*b, w);
  k3<<<grid, wpb * 32, sm, s>>>(*b, w);
  c->c.launches += 2;
#else
  auto kern = w.atom_scratch ? k_windows<true, 7> : k_windows<false, 7>;
  if (set_smem(kern, sm, c->max_smem)) return 1;
  kern<<<grid, wpb * 32, sm, s>>>(*b, w);
#endif
  k_finish<<<(b->n_items + 255) / 256, 256, 0, s>>>(*b, w);
  c->c.launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int pwb_ray_capacity(pwb_ctx* c, const pwb_batch* b, void* stream, int* out_cap) {
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemsetAsync(c->d_tmp, 0, sizeof(int), s));
  int blocks = (b->n_items + 127) / 128;
  if (blocks > 1184) blocks = 1184;
  k_ray_capacity<<<blocks, 128, 0, s>>>(*b, c->d_tmp);
  c->c.launches++;
  int v = 0;
  CK(cudaMemcpyAsync(&v, c->d_tmp, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *out_cap = v;
  return 0;
}

static int run_stages(pwb_ctx* c, const pwb_batch* b, const Ws& w, cudaStream_t s) {
  if (b->stages & (PWB_STAGE_SCALARS | PWB_STAGE_AVGD)) {
    CK(cudaMemsetAsync(b->records, 0, sizeof(double) * (size_t)b->n_items * b->rec_stride, s));
    if (run_prepare(c, b, w, s, (b->stages & PWB_STAGE_AVGD) ? 1 : 0)) return 1;
  }
  if (b->stages & PWB_STAGE_OPT) if (run_opt(c, b, w, s)) return 1;
  if (b->stages & PWB_STAGE_RAYS) if (run_rays(c, b, w, s)) return 1;
  if (b->stages & PWB_STAGE_CLUSTER) if (run_cluster(c, b, w, s)) return 1;
  if (b->stages & PWB_STAGE_WINDOWS) if (run_windows(c, b, w, s)) return 1;
  return 0;
}

// Items are independent, so a large batch is cut into `chunks` contiguous sub-batches that run on
// internal streams: the latency-bound optimiser kernels of one sub-batch overlap the FP-pipe-bound
// ray / grid kernels of another, and the long tail of one kernel no longer idles the GPU.
static int batch_impl(pwb_ctx* c, const pwb_batch* b, cudaStream_t s, const double* h_coords, double* h_records,
                      bool staging_idle = false) {
  if (check(b)) return 1;
  Ws w;
  if (resolve(b, &w)) return 1;
  w.opt_no_list = c->opt_no_list;
  // (huge molecules share one global scratch region sized for the whole batch: not cut into sub-batches)
  const int G = (c->chunks > 1 && b->n_items >= 64 * c->chunks && !w.atom_scratch) ? c->chunks : 1;
  // item_off lives on the device; the host variant needs atom offsets to cut the copies, which it
  // can only do cheaply for equally sized items (trajectory of one molecule): total_atoms / n_items
  const bool uniform = b->n_items > 0 && b->total_atoms % b->n_items == 0 && b->shared_types;
  const size_t apf = uniform ? (size_t)(b->total_atoms / b->n_items) : 0;
  const bool chunk_copies = h_coords && uniform;
  if (h_coords && !chunk_copies)
    CK(cudaMemcpyAsync((void*)b->coords, h_coords, sizeof(double) * 3 * (size_t)b->total_atoms, cudaMemcpyHostToDevice, s));
  if (G == 1) {
    if (chunk_copies) {
      if (staging_idle) {   // pipelined reader: the input copy runs under the kernels of the batch queued before
        CK(cudaMemcpyAsync((void*)b->coords, h_coords, sizeof(double) * 3 * (size_t)b->total_atoms, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->h2d[0], c->copy_stream));
        CK(cudaStreamWaitEvent(s, c->h2d[0], 0));
      } else
        CK(cudaMemcpyAsync((void*)b->coords, h_coords, sizeof(double) * 3 * (size_t)b->total_atoms, cudaMemcpyHostToDevice, s));
    }
    if (run_stages(c, b, w, s)) return 1;
  } else {
    CK(cudaEventRecord(c->fork, s));
    const int per = (b->n_items + G - 1) / G;
    for (int g = 0; g < G; g++) {
      const int first = g * per, count = (first + per <= b->n_items) ? per : b->n_items - first;
      if (count <= 0) break;
      pwb_batch t = *b;
      t.n_items = count;
      t.item_off = b->item_off + first;
      t.records = b->records + (size_t)first * b->rec_stride;
      if (b->centre_override) t.centre_override = b->centre_override + (size_t)first * 3;
      if (b->opt_start) t.opt_start = b->opt_start + (size_t)first * 3;
      if (b->opt_bounds) t.opt_bounds = b->opt_bounds + (size_t)first * 6;
      Ws u = w;
      u.item_d += (size_t)first * ITEM_D; u.item_i += (size_t)first * ITEM_I;
      u.clear_k += (size_t)first * b->ray_cap; u.clear_pos += (size_t)first * b->ray_cap;
      u.clear_a2 += (size_t)first * b->ray_cap; u.labels += (size_t)first * b->ray_cap;
      u.win_k += (size_t)first * b->win_cap;
      u.opt_hist += (size_t)first * LB_HISTORY(3);
      if (b->ray_cap > KNN_SMEM_CAP) { u.big_pts += (size_t)first * b->ray_cap * 4; u.big_i += (size_t)first * b->ray_cap * 2; }
      cudaStream_t sg = c->streams[g];
      // (`staging_idle`: the batch depends on nothing queued on `s` -- its input comes from the host and its
      // buffers are idle -- so sub-batch g simply follows sub-batch g of the previous call on its own stream and the
      // tail of one batch overlaps the head of the next instead of ending in a barrier)
      if (!staging_idle) CK(cudaStreamWaitEvent(sg, c->fork, 0));
      if (chunk_copies) {
        // The copies go to their own stream so that, with two calls in flight, the next batch's input
        // arrives while this batch's kernels still occupy the chunk streams.
        // (`staging_idle`: the caller guarantees that nothing queued earlier still reads the staging buffer, so
        // the copies need not wait for the work already on `s` -- the pipelined trajectory reader's case.)
        if (g == 0 && !staging_idle) CK(cudaStreamWaitEvent(c->copy_stream, c->fork, 0));
        CK(cudaMemcpyAsync((double*)b->coords + 3 * apf * first, h_coords + 3 * apf * first,
                           sizeof(double) * 3 * apf * count, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->h2d[g], c->copy_stream));
        CK(cudaStreamWaitEvent(sg, c->h2d[g], 0));
      }
      if (run_stages(c, &t, u, sg)) return 1;
      if (h_records)
        CK(cudaMemcpyAsync(h_records + (size_t)first * b->rec_stride, t.records,
                           sizeof(double) * (size_t)count * b->rec_stride, cudaMemcpyDeviceToHost, sg));
      CK(cudaEventRecord(c->join[g], sg));
      CK(cudaStreamWaitEvent(s, c->join[g], 0));
    }
    return 0;
  }
  if (h_records)
    CK(cudaMemcpyAsync(h_records, b->records, sizeof(double) * (size_t)b->n_items * b->rec_stride, cudaMemcpyDeviceToHost, s));
  return 0;
}

// Items are independent, so a large batch is cut into `chunks` contiguous sub-batches that run on
// internal streams: the latency-bound optimiser kernels of one sub-batch overlap the FP-pipe-bound
// ray / grid kernels (and, in the host variant, the PCIe copies) of another, and the long tail of one
// kernel no longer idles the GPU.
int pwb_full_analysis_batch(pwb_ctx* c, const pwb_batch* b, void* stream) {
  return batch_impl(c, b, (cudaStream_t)stream, nullptr, nullptr);
}

static int run_stage(pwb_ctx* c, const pwb_batch* b, void* stream, int stages) {
  pwb_batch t = *b;
  t.stages = stages;
  return pwb_full_analysis_batch(c, &t, stream);
}
int pwb_frame_scalars(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_SCALARS); }
int pwb_average_diameter(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_SCALARS | PWB_STAGE_AVGD); }
int pwb_opt_pore(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_OPT); }
int pwb_ray_search(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_RAYS); }
int pwb_cluster(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_CLUSTER); }
int pwb_window_refine(pwb_ctx* c, const pwb_batch* b, void* s) { return run_stage(c, b, s, PWB_STAGE_WINDOWS); }

int pwb_fma_peak(pwb_ctx* c, int fp64, void* stream, double* tflops) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int blocks = 148 * 8, iters = fp64 ? 4096 : 8192;
  void* out = c->d_tmp;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0, s));
    if (fp64) k_fma_peak<double><<<blocks, 256, 0, s>>>((double*)c->c.knn_unit + 1, iters, 1.0000001, 1e-9);
    else k_fma_peak<float><<<blocks, 256, 0, s>>>((float*)out, iters, 1.0000001f, 1e-9f);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  c->c.launches += 4;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
  *tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}

int pwb_full_analysis_host(pwb_ctx* c, const pwb_batch* b, const double* h_coords, double* h_records,
                           double* d_coords, double* d_records, int total_atoms, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  pwb_batch t = *b;
  t.coords = d_coords;
  t.records = d_records;
  t.total_atoms = total_atoms;
  if (batch_impl(c, &t, s, h_coords, h_records)) return 1;
  CK(cudaStreamSynchronize(s));
  return 0;
}

// Same without the final synchronisation: the records are valid once `stream` has been synchronised.  Two
// calls in flight on two streams (each with its own staging buffers and workspace) overlap one batch's copies
// with the other's kernels.
int pwb_full_analysis_host_async(pwb_ctx* c, const pwb_batch* b, const double* h_coords, double* h_records,
                                 double* d_coords, double* d_records, int total_atoms, void* stream) {
  pwb_batch t = *b;
  t.coords = d_coords;
  t.records = d_records;
  t.total_atoms = total_atoms;
  return batch_impl(c, &t, (cudaStream_t)stream, h_coords, h_records, true);
}

}  // extern "C"
