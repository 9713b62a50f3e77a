// pwb_periodic.cu -- periodic decomposition on the GPU: create_supercell (utilities.py:768-810) and
// discrete_molecules (utilities.py:820-1085) for batches of cell-frames.
//
// The reference grows molecules with value-equality list operations over a 27-image supercell; the
// same result is obtained here from (1) a periodic bond table of the cell atoms -- for every
// non-terminal atom a, every bonded (atom b, integer image offset m), sorted by (m, b) -- and (2) an
// ordered, level-synchronous flood fill over nodes (atom, image) whose visiting order reproduces the
// reference's atom order (see oracle/pw_periodic_oracle.py for the rule and its derivation).
//   k_dm_prepare  CTA/frame      : inverse lattice, fractional + rounded coordinates, COM, boundary
//   k_dm_bonds    thread/(frame,atom): bond table, 27 candidate images around the minimum image
//   k_dm_fill     warp/frame     : seeds, flood fill, centre-of-mass filter, output
#include <stdio.h>

#include "pwb_device.cuh"

namespace pwb {

constexpr int DM_ADJ_CAP = 12;
constexpr int DM_FRAME_D = 32;   // per-frame doubles: lattice[9], inv[9], pseudo[3], com[3], boundary[2]

__device__ __forceinline__ double round8(double v) { return rint(v * 1e8) / 1e8; }

struct DmArgs {
  int F, N, rebuild, periodic, n_types, out_cap, mol_cap;
  double tol;
  const double* coords;    // [F][N][3]
  const double* lattice;   // [F][9] row-major (or [1][9] if lattice_shared)
  int lattice_shared;
  const int* atom_type;    // [N]
  const double* type_cov;  // [n_types]
  const double* type_mass;
  const int* type_terminal;
  // workspace
  double* frame_d;   // [F][DM_FRAME_D]
  double* frac;      // [F][N][3]
  double* xr;        // [F][N][3] coordinates rounded to 8 decimals
  unsigned* adj;     // [F][N][DM_ADJ_CAP] packed (image key << 20 | atom)
  int* deg;          // [F][N]
  unsigned* vis;     // [F][ceil(27N/32)]
  unsigned char* rem;  // [F][N]
  int* front;        // [F][2][out_cap]
  double* seed_d;    // [F][N] distance of every atom to the pseudo origin (k_dm_fill)
  // outputs
  int* n_mol;        // [F]
  int* mol_off;      // [F][mol_cap+1]
  int* out_src;      // [F][out_cap] source atom index
  double* out_xyz;   // [F][out_cap][3]
  int* status;       // [F] bit 1: adjacency overflow, 2: output overflow, 4: molecule overflow
};

__global__ void __launch_bounds__(128)
k_dm_prepare(DmArgs A) {
  const int f = blockIdx.x, tid = threadIdx.x;
  __shared__ double red[64];
  __shared__ double inv[9], lat[9];
  double* fd = A.frame_d + (size_t)f * DM_FRAME_D;
  const double* X = A.coords + (size_t)f * A.N * 3;
  if (tid == 0) {
    A.status[f] = 0;
    if (A.periodic) {
      const double* L = A.lattice + (A.lattice_shared ? 0 : (size_t)f * 9);
      for (int i = 0; i < 9; i++) lat[i] = L[i];
      double a = L[0], b = L[1], c = L[2], d = L[3], e = L[4], g = L[5], h = L[6], i2 = L[7], j = L[8];
      double det = a * (e * j - g * i2) - b * (d * j - g * h) + c * (d * i2 - e * h);
      inv[0] = (e * j - g * i2) / det; inv[1] = (c * i2 - b * j) / det; inv[2] = (b * g - c * e) / det;
      inv[3] = (g * h - d * j) / det; inv[4] = (a * j - c * h) / det; inv[5] = (c * d - a * g) / det;
      inv[6] = (d * i2 - e * h) / det; inv[7] = (b * h - a * i2) / det; inv[8] = (a * e - b * d) / det;
      for (int i = 0; i < 9; i++) { fd[i] = lat[i]; fd[9 + i] = inv[i]; }
    }
  }
  __syncthreads();
  double mx = 0, my = 0, mz = 0, mt = 0;
  for (int a = tid; a < A.N; a += blockDim.x) {
    double x = X[3 * a], y = X[3 * a + 1], z = X[3 * a + 2];
    double m = A.type_mass[A.atom_type[a]];
    mx += x * m; my += y * m; mz += z * m; mt += m;
    double* xr = A.xr + ((size_t)f * A.N + a) * 3;
    xr[0] = round8(x); xr[1] = round8(y); xr[2] = round8(z);
    if (A.periodic) {
      double* fr = A.frac + ((size_t)f * A.N + a) * 3;
      fr[0] = inv[0] * x + inv[1] * y + inv[2] * z;
      fr[1] = inv[3] * x + inv[4] * y + inv[5] * z;
      fr[2] = inv[6] * x + inv[7] * y + inv[8] * z;
    }
    A.rem[(size_t)f * A.N + a] = 1;
  }
  mx = block_sum(mx, red); my = block_sum(my, red); mz = block_sum(mz, red); mt = block_sum(mt, red);
  if (tid == 0) {
    double cx = mx / mt, cy = my / mt, cz = mz / mt;
    fd[21] = cx; fd[22] = cy; fd[23] = cz;
    if (A.periodic) {
      // pseudo origin = lattice * (0.26, 0.25, 0.25)   (utilities.py:905-906)
      fd[18] = lat[0] * 0.26 + lat[1] * 0.25 + lat[2] * 0.25;
      fd[19] = lat[3] * 0.26 + lat[4] * 0.25 + lat[5] * 0.25;
      fd[20] = lat[6] * 0.26 + lat[7] * 0.25 + lat[8] * 0.25;
      // np.allclose(com, [0.01, 0, 0], atol=1): |a-b| <= atol + rtol*|b|  (utilities.py:919-923)
      bool near = fabs(cx - 0.01) <= 1.0 + 1e-5 * 0.01 && fabs(cy) <= 1.0 && fabs(cz) <= 1.0;
      fd[24] = near ? -0.5 : 0.0; fd[25] = near ? 0.5 : 1.0;
    } else {
      fd[18] = cx + 0.01; fd[19] = cy; fd[20] = cz;       // utilities.py:925-927
    }
  }
}

// Periodic images of atom b bonded to atom a: the minimum image offset m0, then the offsets around it that can
// be within the cut-off at all -- |delta f_c| <= |delta x| * |row c of inv(lattice)|, so an image needs
// |df_c - m_c| <= tau_c on every axis; for ordinary cells that leaves at most the minimum image and rejects
// > 99 % of the pairs after three subtractions.  Calls emit(image key << 20 | b) for every bond found.
template <class Emit>
__device__ __forceinline__ void dm_pair_images(const DmArgs& A, const double (&L)[9], double taux, double tauy, double tauz,
                                               double max_dist, double ax, double ay, double az, double afx, double afy,
                                               double afz, double rs, double bx, double by, double bz, double bfx,
                                               double bfy, double bfz, int bb, Emit emit) {
  const double dfx = afx - bfx, dfy = afy - bfy, dfz = afz - bfz;
  const double m0x = rint(dfx), m0y = rint(dfy), m0z = rint(dfz);
  const double ex = dfx - m0x, ey = dfy - m0y, ez = dfz - m0z;
  if (fabs(ex) - 1.0 > taux || fabs(ey) - 1.0 > tauy || fabs(ez) - 1.0 > tauz) return;
  if (taux < 0.5 && tauy < 0.5 && tauz < 0.5 && (fabs(ex) > taux || fabs(ey) > tauy || fabs(ez) > tauz)) return;
  for (int ia = -1; ia <= 1; ia++) {
    if (fabs(ex - ia) > taux) continue;
    for (int ib = -1; ib <= 1; ib++) {
      if (fabs(ey - ib) > tauy) continue;
      for (int ic = -1; ic <= 1; ic++) {
        if (fabs(ez - ic) > tauz) continue;
        const double ma = m0x + ia, mb = m0y + ib, mc = m0z + ic;
        double X, Y, Z;
        if (ma == 0 && mb == 0 && mc == 0) { X = bx; Y = by; Z = bz; }
        else {
          double fx = bfx + ma, fy = bfy + mb, fz = bfz + mc;
          X = round8(L[0] * fx + L[1] * fy + L[2] * fz);
          Y = round8(L[3] * fx + L[4] * fy + L[5] * fz);
          Z = round8(L[6] * fx + L[7] * fy + L[8] * fz);
        }
        double dx = ax - X, dy = ay - Y, dz = az - Z;
        double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 >= max_dist * max_dist) continue;
        double d = sqrt(d2);
        if (d > 0.1 && d < max_dist && rs - A.tol < d && d < rs + A.tol) {
          int ka = (int)ma + 4, kb = (int)mb + 4, kc = (int)mc + 4;
          if (ka < 0 || ka > 8 || kb < 0 || kb > 8 || kc < 0 || kc > 8) continue;   // > 4 cells away: unreachable
          emit(((unsigned)((ka << 8) | (kb << 4) | kc) << 20) | (unsigned)bb);   // sorts like (ka, kb, kc)
        }
      }
    }
  }
}

__device__ __forceinline__ void dm_store_bonds(const DmArgs& A, int f, int a, unsigned* list, int n, int overflow) {
  // sort by (image key, atom): ascending supercell index after translation
  for (int i = 1; i < n; i++) {
    unsigned v = list[i];
    int j = i - 1;
    while (j >= 0 && list[j] > v) { list[j + 1] = list[j]; j--; }
    list[j + 1] = v;
  }
  unsigned* out = A.adj + ((size_t)f * A.N + a) * DM_ADJ_CAP;
  for (int i = 0; i < n; i++) out[i] = list[i];
  A.deg[(size_t)f * A.N + a] = n;
  if (overflow) atomicOr(&A.status[f], 1);
}

// Bond table with a cell list (periodic + rebuild): one CTA per frame.  Atoms are binned by wrapped fractional
// coordinates into G_c = floor(1 / tau_c) <= 16 bins per axis (tau_c = the largest fractional distance a bond
// can span on axis c), so every bonded image of b lies in the 27 bins around a's; an axis too short for three
// bins is not binned.  The pair test is dm_pair_images, unchanged, on ~20 candidates instead of N.
constexpr int DM_CL_THREADS = 256, DM_CL_MAXG = 12;
constexpr int DM_CL_NB = DM_CL_MAXG * DM_CL_MAXG * DM_CL_MAXG;
// shared memory: bin starts + cursors, then the atoms SORTED BY BIN (7 doubles + index each), so that the
// candidate loop streams contiguous shared memory instead of chasing indices through global memory
#define DM_CL_SMEM(N) (sizeof(int) * (2 * DM_CL_NB + 2) + (size_t)(N) * (7 * sizeof(double) + 2 * sizeof(int)) + 16)
__global__ void __launch_bounds__(DM_CL_THREADS, 2)
k_dm_bonds_cells(DmArgs A, double max_dist) {
  extern __shared__ __align__(16) double smd[];
  const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, N = A.N;
  double* qx = smd; double* qy = qx + N; double* qz = qy + N;           // rounded Cartesian
  double* qfx = qz + N; double* qfy = qfx + N; double* qfz = qfy + N;    // fractional
  double* qc = qfz + N;                                                  // covalent radius
  int* qi = (int*)(qc + N);                                              // atom index
  int* bstart = qi + N;                    // [NB + 1]
  int* bcur = bstart + DM_CL_NB + 1;       // [NB]
  int* nts = bcur + DM_CL_NB;              // [N] sorted slots of the non-terminal atoms
  __shared__ int nnt;
  __shared__ int G[3];
  __shared__ double tau[3];
  const double* fd = A.frame_d + (size_t)f * DM_FRAME_D;
  const double* frac = A.frac + (size_t)f * N * 3;
  const double* xr = A.xr + (size_t)f * N * 3;
  double L[9];
#pragma unroll
  for (int i = 0; i < 9; i++) L[i] = fd[i];
  if (tid < 3) {
    const double md = max_dist + 1e-6;     // slack for the 8-decimal rounding of the coordinates
    const double* r = fd + 9 + 3 * tid;
    const double t = md * sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * (1.0 + 1e-9);
    tau[tid] = t;
    int g = (int)floor(1.0 / (t * (1.0 + 1e-9) + 1e-12));
    G[tid] = g < 3 ? 1 : (g > DM_CL_MAXG ? DM_CL_MAXG : g);
  }
  __syncthreads();
  const int Gx = G[0], Gy = G[1], Gz = G[2], nb = Gx * Gy * Gz;
  const double taux = tau[0], tauy = tau[1], tauz = tau[2];
  for (int c = tid; c <= nb; c += nt) bstart[c] = 0;
  __syncthreads();
  auto bin_of = [&](double fx, double fy, double fz) {
    const int bx = min(Gx - 1, max(0, (int)((fx - floor(fx)) * Gx)));
    const int by = min(Gy - 1, max(0, (int)((fy - floor(fy)) * Gy)));
    const int bz = min(Gz - 1, max(0, (int)((fz - floor(fz)) * Gz)));
    return (bx * Gy + by) * Gz + bz;
  };
  for (int a = tid; a < N; a += nt) atomicAdd(&bstart[bin_of(frac[3 * a], frac[3 * a + 1], frac[3 * a + 2]) + 1], 1);
  __syncthreads();
  // inclusive scan of the counts: each thread owns a contiguous chunk, chunk totals scanned by one warp
  {
    __shared__ int csum[DM_CL_THREADS];
    const int per = (nb + nt - 1) / nt, c0 = 1 + tid * per, c1 = min(nb + 1, c0 + per);
    int t = 0;
    for (int c = c0; c < c1; c++) t += bstart[c];
    csum[tid] = t;
    __syncthreads();
    if (tid < 32) {
      int carry = 0;
      for (int base = 0; base < nt; base += 32) {
        int v = csum[base + tid], x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= o) x += y; }
        csum[base + tid] = carry + x - v;              // exclusive prefix of the chunk
        carry += __shfl_sync(0xffffffffu, x, 31);
      }
    }
    __syncthreads();
    int run = csum[tid];
    for (int c = c0; c < c1; c++) { run += bstart[c]; bstart[c] = run; }
  }
  __syncthreads();
  for (int c = tid; c < nb; c += nt) bcur[c] = bstart[c];
  __syncthreads();
  for (int a = tid; a < N; a += nt) {
    const double fx = frac[3 * a], fy = frac[3 * a + 1], fz = frac[3 * a + 2];
    const int pos = atomicAdd(&bcur[bin_of(fx, fy, fz)], 1);
    qfx[pos] = fx; qfy[pos] = fy; qfz[pos] = fz;
    qx[pos] = xr[3 * a]; qy[pos] = xr[3 * a + 1]; qz[pos] = xr[3 * a + 2];
    qc[pos] = A.type_cov[A.atom_type[a]];
    qi[pos] = a;
  }
  __syncthreads();
  // Pair phase: one thread per NON-TERMINAL atom (their sorted slots are compacted first, so the lanes of a warp all
  // have work), walking the <= 9 (x, y) bin columns around it; the <= 3 z bins of a column are adjacent in the sorted
  // order, i.e. one contiguous range of candidates (two if the column wraps around the cell).
  if (tid == 0) nnt = 0;
  __syncthreads();
  for (int sa = tid; sa < N; sa += nt) {
    if (!A.type_terminal[A.atom_type[qi[sa]]]) nts[atomicAdd(&nnt, 1)] = sa;
    else A.deg[(size_t)f * N + qi[sa]] = 0;
  }
  __syncthreads();
  for (int it = tid; it < nnt; it += nt) {
    const int sa = nts[it];
    const int a = qi[sa];
    unsigned list[DM_ADJ_CAP];
    int n = 0, overflow = 0;
    const double ax = qx[sa], ay = qy[sa], az = qz[sa], afx = qfx[sa], afy = qfy[sa], afz = qfz[sa], acov = qc[sa];
    const int bin = bin_of(afx, afy, afz);
    const int bz = bin % Gz, by = (bin / Gz) % Gy, bx = bin / (Gz * Gy);
    // z bins bz-1 .. bz+1 (mod Gz) as one or two contiguous bin ranges [lo, hi]
    int lo1 = 0, hi1 = 0, lo2 = 1, hi2 = 0;
    if (Gz > 1) {
      lo1 = bz - 1; hi1 = bz + 1;
      if (lo1 < 0) { lo1 = 0; lo2 = hi2 = Gz - 1; }
      else if (hi1 >= Gz) { hi1 = Gz - 1; lo2 = hi2 = 0; }
    }
    auto emit = [&](unsigned key) { if (n < DM_ADJ_CAP) list[n++] = key; else overflow = 1; };
    // (no unrolling: with the 27-image test inlined, unrolled bin loops made a 1.5 MB kernel that spent
    // 92 % of its time waiting for instructions)
#pragma unroll 1
    for (int ox = (Gx > 1 ? -1 : 0); ox <= (Gx > 1 ? 1 : 0); ox++)
#pragma unroll 1
      for (int oy = (Gy > 1 ? -1 : 0); oy <= (Gy > 1 ? 1 : 0); oy++) {
        const int c0 = (((bx + ox + Gx) % Gx) * Gy + (by + oy + Gy) % Gy) * Gz;
#pragma unroll 1
        for (int r = 0; r < 2; r++) {
          const int lo = r ? lo2 : lo1, hi = r ? hi2 : hi1;
          if (lo > hi) continue;
#pragma unroll 1
          for (int t = bstart[c0 + lo]; t < bstart[c0 + hi + 1]; t++)
            dm_pair_images(A, L, taux, tauy, tauz, max_dist, ax, ay, az, afx, afy, afz, acov + qc[t],
                           qx[t], qy[t], qz[t], qfx[t], qfy[t], qfz[t], qi[t], emit);
        }
      }
    dm_store_bonds(A, f, a, list, n, overflow);
  }
}

// bond table: one thread per (frame, parent atom); tiles of atoms staged in shared memory
constexpr int DM_TILE = 128;
__global__ void __launch_bounds__(DM_TILE)
k_dm_bonds(DmArgs A, double max_dist) {
  const int f = blockIdx.y, a = blockIdx.x * DM_TILE + threadIdx.x;
  __shared__ double sfx[DM_TILE], sfy[DM_TILE], sfz[DM_TILE], sx[DM_TILE], sy[DM_TILE], sz[DM_TILE], sc[DM_TILE];
  const double* fd = A.frame_d + (size_t)f * DM_FRAME_D;
  const double* frac = A.frac + (size_t)f * A.N * 3;
  const double* xr = A.xr + (size_t)f * A.N * 3;
  double L[9];
#pragma unroll
  for (int i = 0; i < 9; i++) L[i] = A.periodic ? fd[i] : 0.0;
  const bool active = a < A.N && !A.type_terminal[A.atom_type[a < A.N ? a : 0]];
  double ax = 0, ay = 0, az = 0, afx = 0, afy = 0, afz = 0, acov = 0;
  if (a < A.N) {
    ax = xr[3 * a]; ay = xr[3 * a + 1]; az = xr[3 * a + 2];
    acov = A.type_cov[A.atom_type[a]];
    if (A.periodic) { afx = frac[3 * a]; afy = frac[3 * a + 1]; afz = frac[3 * a + 2]; }
  }
  unsigned list[DM_ADJ_CAP];
  int n = 0, overflow = 0;
  const bool images = A.rebuild != 0;
  double taux = 0, tauy = 0, tauz = 0;
  if (A.periodic) {
    const double md = max_dist + 1e-6;     // slack for the 8-decimal rounding of the coordinates
    taux = md * sqrt(fd[9] * fd[9] + fd[10] * fd[10] + fd[11] * fd[11]) * (1.0 + 1e-9);
    tauy = md * sqrt(fd[12] * fd[12] + fd[13] * fd[13] + fd[14] * fd[14]) * (1.0 + 1e-9);
    tauz = md * sqrt(fd[15] * fd[15] + fd[16] * fd[16] + fd[17] * fd[17]) * (1.0 + 1e-9);
  }
  for (int t0 = 0; t0 < A.N; t0 += DM_TILE) {
    int b = t0 + threadIdx.x;
    __syncthreads();
    if (b < A.N) {
      sx[threadIdx.x] = xr[3 * b]; sy[threadIdx.x] = xr[3 * b + 1]; sz[threadIdx.x] = xr[3 * b + 2];
      sc[threadIdx.x] = A.type_cov[A.atom_type[b]];
      if (A.periodic) { sfx[threadIdx.x] = frac[3 * b]; sfy[threadIdx.x] = frac[3 * b + 1]; sfz[threadIdx.x] = frac[3 * b + 2]; }
    }
    __syncthreads();
    if (!active) continue;
    const int lim = min(DM_TILE, A.N - t0);
    for (int j = 0; j < lim; j++) {
      const int bb = t0 + j;
      const double rs = acov + sc[j];
      if (!images) {
        double dx = ax - sx[j], dy = ay - sy[j], dz = az - sz[j];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (d > 0.1 && d < max_dist && rs - A.tol < d && d < rs + A.tol) {
          if (n < DM_ADJ_CAP) list[n++] = ((unsigned)((4 << 8) | (4 << 4) | 4) << 20) | (unsigned)bb; else overflow = 1;
        }
        continue;
      }
      dm_pair_images(A, L, taux, tauy, tauz, max_dist, ax, ay, az, afx, afy, afz, rs,
                     sx[j], sy[j], sz[j], sfx[j], sfy[j], sfz[j], bb,
                     [&](unsigned key) { if (n < DM_ADJ_CAP) list[n++] = key; else overflow = 1; });
    }
  }
  if (a >= A.N) return;
  dm_store_bonds(A, f, a, list, n, overflow);
}

// Flood fill: one warp per frame, FRONTIER-PARALLEL and order preserving.  The reference grows a molecule level by
// level; inside a level the parents are taken in frontier order and each parent appends, first, its bonded cell
// atoms that are still unassigned (ascending index) and then its other bonded images (ascending supercell index),
// the first occurrence of a node winning (utilities.py:985-1044, restated in oracle/pw_periodic_oracle.py).  Every
// candidate therefore has the order key (parent rank, pass, adjacency slot) and a node goes to the candidate with
// the SMALLEST key.  Parents are processed 32 at a time, one per lane:
//   1. each lane walks its parent's bond list ONCE: node id, pass, visited test; the survivors are cached in shared
//      memory and entered into a small hash table node -> min key (atomicMin);
//   2. the lanes look their candidates up again: the one whose key is in the table owns the node;
//   3. the winners are compacted in key order by a warp scan -- straight into the output list, whose tail IS the
//      next frontier (levels are contiguous in breadth-first order) -- and marked visited; every lane wipes the
//      table slots it used.  Later parents (larger keys) see the winners as taken.
// A node is (image << 20 | atom); its visited bit is image * N + atom.  With SMEM_STATE the visited bits and the
// "still unassigned" bytes of the frame live in shared memory (4.4 bytes per atom and warp).
#ifndef DM_FILL_STATE_SMEM_MAX
#define DM_FILL_STATE_SMEM_MAX (56 * 1024)
#endif
constexpr int DM_FILL_WARPS = 4;
constexpr int DM_HASH = 512;                     // > 32 parents x 12 bond slots, power of two
constexpr int DM_CAND_STRIDE = 13;               // per-lane row of the candidate cache (odd: no bank conflicts)
constexpr unsigned DM_EMPTY = 0xffffffffu;
constexpr int DM_FILL_FIXED_WORDS = 2 * DM_HASH + 32 * DM_CAND_STRIDE + 32 * DM_CAND_STRIDE / 2;

__device__ __forceinline__ unsigned dm_hash_slot(unsigned node) { return (node * 2654435761u) >> 23; }   // 9 bits
__host__ __device__ inline int dm_vis_words(int N, int rebuild) { return ((rebuild ? 27 : 1) * N + 31) / 32; }
__host__ __device__ inline size_t dm_fill_warp_words(int N, int rebuild, bool smem_state) {
  return (size_t)DM_FILL_FIXED_WORDS + (smem_state ? (size_t)dm_vis_words(N, rebuild) + (size_t)(N + 3) / 4 : 0);
}

template <bool SMEM_STATE>
__global__ void __launch_bounds__(DM_FILL_WARPS * 32)
k_dm_fill(DmArgs A) {
  extern __shared__ __align__(16) unsigned dm_sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.x * DM_FILL_WARPS + wid;
  if (f >= A.F) return;
  const int N = A.N;
  unsigned* hn = dm_sm + (size_t)wid * dm_fill_warp_words(N, A.rebuild, SMEM_STATE);
  unsigned* hk = hn + DM_HASH;
  int* cnode = (int*)(hk + DM_HASH) + lane * DM_CAND_STRIDE;
  unsigned short* cslot = (unsigned short*)((int*)(hk + DM_HASH) + 32 * DM_CAND_STRIDE) + lane * DM_CAND_STRIDE;
  const double* fd = A.frame_d + (size_t)f * DM_FRAME_D;
  const double* xr = A.xr + (size_t)f * N * 3;
  const double* frac = A.frac + (size_t)f * N * 3;
  const unsigned* adj = A.adj + (size_t)f * N * DM_ADJ_CAP;
  const int* deg = A.deg + (size_t)f * N;
  const int vis_words = dm_vis_words(N, A.rebuild);
  unsigned* vis = SMEM_STATE ? hn + DM_FILL_FIXED_WORDS : A.vis + (size_t)f * ((27 * N + 31) / 32);
  unsigned char* rem = SMEM_STATE ? (unsigned char*)(vis + vis_words) : A.rem + (size_t)f * N;
  int* osrc = A.out_src + (size_t)f * A.out_cap;
  double* oxyz = A.out_xyz + (size_t)f * A.out_cap * 3;
  int* moff = A.mol_off + (size_t)f * (A.mol_cap + 1);
  const int cimg = A.rebuild ? 13 : 0;           // image index of the cell itself
  const double psx = fd[18], psy = fd[19], psz = fd[20];
  const double yy = __dadd_rn(__dadd_rn(__dmul_rn(psx, psx), __dmul_rn(psy, psy)), __dmul_rn(psz, psz));
  int nout = 0, nmol = 0, st = 0;
  if (lane == 0) moff[0] = 0;
  // distance of every atom to the pseudo origin, once, with sklearn's euclidean_distances operation order
  // (utilities.py:944-953); the seed searches below only compare
  double* sd = A.seed_d + (size_t)f * N;
  for (int a = lane; a < N; a += 32) {
    double x = xr[3 * a], y = xr[3 * a + 1], z = xr[3 * a + 2];
    double dot = __dadd_rn(__dadd_rn(__dmul_rn(x, psx), __dmul_rn(y, psy)), __dmul_rn(z, psz));
    double xx = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), xx), yy);
    sd[a] = A.type_terminal[A.atom_type[a]] ? PWB_INF : sqrt(fmax(d2, 0.0));     // terminal atoms never seed
    if (SMEM_STATE) rem[a] = 1;                  // (the global copy is set by k_dm_prepare)
  }
  for (int w = lane; w < vis_words; w += 32) vis[w] = 0u;
  for (int t = lane; t < 2 * DM_HASH; t += 32) hn[t] = DM_EMPTY;       // hn and hk are adjacent
  __syncwarp();
  while (true) {
    // ---- seed: remaining non-terminal atom closest to the pseudo origin, first arg-min
    ValIdx best; best.v = PWB_INF; best.i = 0x7fffffff;
    for (int a = lane; a < N; a += 32) {
      if (!rem[a]) continue;
      ValIdx c; c.v = sd[a]; c.i = a;
      if (c.v < PWB_INF) best = vi_min(best, c);
    }
    best = warp_vi_min(best);
    if (best.i == 0x7fffffff) break;
    const int start = nout;
    if (start >= A.out_cap) { st |= 2; break; }
    int overflow = 0;
    {
      const int seed = (cimg << 20) | best.i, bit = cimg * N + best.i;
      if (lane == 0) { osrc[start] = seed; vis[bit >> 5] |= 1u << (bit & 31); }
    }
    __syncwarp();
    int lvl0 = start, lvl1 = start + 1;          // the current level = osrc[lvl0, lvl1)
    while (lvl1 > lvl0) {
      int nn = lvl1;                             // children are appended behind the level
      for (int q0 = lvl0; q0 < lvl1; q0 += 32) {
        const int q = q0 + lane;
        int a = 0, sa = 0, sb = 0, sc = 0, dg = 0;
        if (q < lvl1) {
          const int p = osrc[q];
          a = p & 0xfffff;
          const int img = p >> 20;
          if (A.rebuild) { sa = img / 9 - 1; sb = (img / 3) % 3 - 1; sc = img % 3 - 1; }
          dg = deg[a];                           // (0 for terminal atoms)
        }
        // 1. candidates of this lane's parent: node -> smallest key (lane, pass, slot)
        unsigned valid = 0, pass1 = 0;
        for (int e = 0; e < dg; e++) {
          const unsigned k = adj[(size_t)a * DM_ADJ_CAP + e];
          const int b = (int)(k & 0xfffffu);
          int im = 0;
          if (A.rebuild) {
            const int ik = (int)(k >> 20);
            const int ma = (ik >> 8) - 3 + sa, mb = ((ik >> 4) & 15) - 3 + sb, mc = (ik & 15) - 3 + sc;   // image + 1
            if ((unsigned)ma > 2u || (unsigned)mb > 2u || (unsigned)mc > 2u) continue;    // outside the supercell
            im = ma * 9 + mb * 3 + mc;
          }
          const int bit = im * N + b;
          if (vis[bit >> 5] & (1u << (bit & 31))) continue;
          const unsigned pass = (im == cimg && rem[b]) ? 0u : 1u;     // pass 0: remaining cell atoms; 1: the rest
          const unsigned node = ((unsigned)im << 20) | (unsigned)b;
          const unsigned key = (unsigned)(lane * 24) + pass * 12u + (unsigned)e;
          unsigned h = dm_hash_slot(node);
          while (true) {
            const unsigned old = atomicCAS(&hn[h], DM_EMPTY, node);
            if (old == DM_EMPTY || old == node) { atomicMin(&hk[h], key); break; }
            h = (h + 1) & (DM_HASH - 1);
          }
          cnode[e] = (int)node; cslot[e] = (unsigned short)h;
          valid |= 1u << e; pass1 |= pass << e;
        }
        __syncwarp();
        // 2. winners of this lane, as a bit mask over (pass, slot)
        unsigned keep = 0;
        for (unsigned v = valid; v; v &= v - 1) {
          const int e = __ffs(v) - 1;
          const unsigned pe = ((pass1 >> e) & 1u) * 12u + (unsigned)e;
          if (hk[cslot[e]] == (unsigned)(lane * 24) + pe) keep |= 1u << pe;
        }
        // 3. ordered compaction behind the level, visited marks, table wipe
        const int cnt = __popc(keep);
        int inc = cnt;
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        int pos = nn + inc - cnt;
        nn += __shfl_sync(0xffffffffu, inc, 31);
        __syncwarp();
        for (unsigned kbits = keep; kbits; kbits &= kbits - 1) {
          const int pe = __ffs(kbits) - 1;
          const int node = cnode[pe >= 12 ? pe - 12 : pe];
          if (pos < A.out_cap) osrc[pos] = node; else overflow = 1;
          pos++;
          const int bit = (node >> 20) * N + (node & 0xfffff);
          atomicOr(&vis[bit >> 5], 1u << (bit & 31));
        }
        for (unsigned v = valid; v; v &= v - 1) {
          const int h = cslot[__ffs(v) - 1];
          hn[h] = DM_EMPTY; hk[h] = DM_EMPTY;
        }
        __syncwarp();
      }
      // the level is done: its cell atoms are assigned now (utilities.py:1040-1044)
      for (int q = lvl0 + lane; q < lvl1; q += 32) {
        const int p = osrc[q];
        if ((p >> 20) == cimg) rem[p & 0xfffff] = 0;
      }
      overflow = __any_sync(0xffffffffu, overflow);
      if (overflow || nn > A.out_cap) { overflow = 1; break; }
      __syncwarp();
      lvl0 = lvl1; lvl1 = nn;
    }
    if (overflow) { st |= 2; break; }
    nout = lvl1;
    // ---- resolve nodes to (source atom, coordinates); centre of mass for the filter
    double mx = 0, my = 0, mz = 0, mt = 0;
    for (int q = start + lane; q < nout; q += 32) {
      const int p = osrc[q];
      const int a = p & 0xfffff, img = p >> 20;
      const int bit = img * N + a;
      vis[bit >> 5] = 0u;                          // every set bit belongs to a node of this molecule
      double X, Y, Z;
      if (!A.rebuild || img == 13) { X = xr[3 * a]; Y = xr[3 * a + 1]; Z = xr[3 * a + 2]; }
      else {
        double fx = frac[3 * a] + (img / 9 - 1), fy = frac[3 * a + 1] + ((img / 3) % 3 - 1), fz = frac[3 * a + 2] + (img % 3 - 1);
        X = round8(fd[0] * fx + fd[1] * fy + fd[2] * fz);
        Y = round8(fd[3] * fx + fd[4] * fy + fd[5] * fz);
        Z = round8(fd[6] * fx + fd[7] * fy + fd[8] * fz);
      }
      oxyz[3 * q] = X; oxyz[3 * q + 1] = Y; oxyz[3 * q + 2] = Z;
      osrc[q] = a;
      double m = A.type_mass[A.atom_type[a]];
      mx += X * m; my += Y * m; mz += Z * m; mt += m;
    }
    mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz); mt = warp_sum(mt);
    bool keepmol = true;
    if (A.rebuild) {   // utilities.py:1062-1077
      double cx = mx / mt, cy = my / mt, cz = mz / mt;
      double f0 = round8(fd[9] * cx + fd[10] * cy + fd[11] * cz);
      double f1 = round8(fd[12] * cx + fd[13] * cy + fd[14] * cz);
      double f2 = round8(fd[15] * cx + fd[16] * cy + fd[17] * cz);
      const double lo = fd[24], hi = fd[25];
      keepmol = f0 >= lo && f0 < hi && f1 >= lo && f1 < hi && f2 >= lo && f2 < hi;
    }
    if (keepmol) {
      if (nmol >= A.mol_cap) { st |= 4; break; }
      nmol++;
      if (lane == 0) moff[nmol] = nout;
    } else nout = start;
    __syncwarp();
  }
  if (lane == 0) { A.n_mol[f] = nmol; if (st) atomicOr(&A.status[f], st); }
}

// ---- compaction of the per-frame outputs into ONE dense batch for the analysis kernels -----------------------
// k_dm_scan  one CTA: exclusive prefix sums over the frames of (molecules kept, atoms emitted) -> per-frame bases,
//            totals and the OR of the status words
// k_dm_emit  CTA per frame: item offsets, coordinates, atom types and (frame, molecule) keys at the frame's bases
struct DmPack {
  int F, N, out_cap, mol_cap;
  const int* n_mol; const int* mol_off; const int* out_src; const double* out_xyz; const int* status;
  const int* atom_type;
  int* base;           // [F][2] item base, atom base (scratch)
  int* item_off;       // [items + 1]
  double* coords;      // [atoms][3]
  int* types;          // [atoms]
  int* item_frame;     // [items]
  int* item_mol;       // [items]
  int* counts;         // [4] items, atoms, largest molecule, status OR
};

__global__ void __launch_bounds__(1024)
k_dm_scan(DmPack P) {
  __shared__ int s_items[1024], s_atoms[1024], s_st[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (P.F + nt - 1) / nt, f0 = tid * per, f1 = min(P.F, f0 + per);
  int items = 0, atoms = 0, st = 0;
  for (int f = f0; f < f1; f++) {
    const int nm = P.n_mol[f];
    items += nm;
    atoms += P.mol_off[(size_t)f * (P.mol_cap + 1) + nm];
    st |= P.status[f];
  }
  s_items[tid] = items; s_atoms[tid] = atoms; s_st[tid] = st;
  __syncthreads();
  for (int d = 1; d < nt; d <<= 1) {                      // inclusive Hillis-Steele scan over the strips
    const int a = tid >= d ? s_items[tid - d] : 0, b = tid >= d ? s_atoms[tid - d] : 0, c = tid >= d ? s_st[tid - d] : 0;
    __syncthreads();
    s_items[tid] += a; s_atoms[tid] += b; s_st[tid] |= c;
    __syncthreads();
  }
  int ib = s_items[tid] - items, ab = s_atoms[tid] - atoms;
  for (int f = f0; f < f1; f++) {
    const int nm = P.n_mol[f];
    P.base[2 * f] = ib; P.base[2 * f + 1] = ab;
    ib += nm;
    ab += P.mol_off[(size_t)f * (P.mol_cap + 1) + nm];
  }
  if (tid == nt - 1) {
    P.counts[0] = s_items[tid]; P.counts[1] = s_atoms[tid]; P.counts[2] = 0; P.counts[3] = s_st[tid];
    P.item_off[s_items[tid]] = s_atoms[tid];
  }
}

__global__ void __launch_bounds__(128)
k_dm_emit(DmPack P) {
  const int f = blockIdx.x, tid = threadIdx.x;
  const int nm = P.n_mol[f];
  const int* moff = P.mol_off + (size_t)f * (P.mol_cap + 1);
  const int ib = P.base[2 * f], ab = P.base[2 * f + 1], na = moff[nm];
  int big = 0;
  for (int k = tid; k < nm; k += blockDim.x) {
    P.item_off[ib + k] = ab + moff[k];
    P.item_frame[ib + k] = f;
    P.item_mol[ib + k] = k;
    big = max(big, moff[k + 1] - moff[k]);
  }
  const int* src = P.out_src + (size_t)f * P.out_cap;
  const double* xyz = P.out_xyz + (size_t)f * P.out_cap * 3;
  for (int j = tid; j < na; j += blockDim.x) P.types[ab + j] = P.atom_type[src[j]];
  for (int j = tid; j < 3 * na; j += blockDim.x) P.coords[(size_t)ab * 3 + j] = xyz[j];
  big = __reduce_max_sync(0xffffffffu, big);
  if ((tid & 31) == 0 && big > 0) atomicMax(&P.counts[2], big);
}

}  // namespace pwb

using namespace pwb;

extern "C" {

static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

int pwb_dm_workspace_layout(int F, int N, int out_cap, pwb_dm_layout* L) {
  size_t o = 0;
  L->frame_d = o; o = up256(o + sizeof(double) * (size_t)F * DM_FRAME_D);
  L->frac = o; o = up256(o + sizeof(double) * (size_t)F * N * 3);
  L->xr = o; o = up256(o + sizeof(double) * (size_t)F * N * 3);
  L->adj = o; o = up256(o + sizeof(unsigned) * (size_t)F * N * DM_ADJ_CAP);
  L->deg = o; o = up256(o + sizeof(int) * (size_t)F * N);
  L->vis = o; o = up256(o + sizeof(unsigned) * (size_t)F * ((27 * (size_t)N + 31) / 32));
  L->rem = o; o = up256(o + (size_t)F * N);
  L->front = o; o = up256(o + sizeof(int) * (size_t)F * 2 * out_cap);
  o = up256(o + sizeof(double) * (size_t)F * N);      // seed distances, see dm_seed_offset
  L->total = o;
  return 0;
}
// the per-atom distances to the pseudo origin (k_dm_fill) live behind the frontier buffers
static size_t dm_seed_offset(const pwb_dm_layout& L, int F, int out_cap) {
  return up256(L.front + sizeof(int) * (size_t)F * 2 * out_cap);
}
size_t pwb_dm_workspace_bytes(int F, int N, int out_cap) {
  pwb_dm_layout L;
  pwb_dm_workspace_layout(F, N, out_cap, &L);
  return L.total;
}

// discrete_molecules for F cell-frames of N atoms each (utilities.py:820-1085; with `rebuild` the
// supercell of create_supercell, utilities.py:768-810, is implied).  All pointers are device pointers.
//   coords [F][N][3]; lattice [F][9] row-major (one matrix if lattice_shared; NULL = non-periodic)
//   atom_type [N]; type_cov/type_mass/type_terminal [n_types]
//   outputs: n_mol [F]; mol_off [F][mol_cap+1]; out_src [F][out_cap]; out_xyz [F][out_cap][3];
//            status [F] (1 bond-table overflow, 2 output overflow, 4 molecule overflow)
int pwb_discrete_molecules(int F, int N, int rebuild, double tol, double max_cov,
                           const double* coords, const double* lattice, int lattice_shared,
                           const int* atom_type, int n_types, const double* type_cov, const double* type_mass,
                           const int* type_terminal, int out_cap, int mol_cap,
                           int* n_mol, int* mol_off, int* out_src, double* out_xyz, int* status,
                           void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  pwb_dm_layout L;
  pwb_dm_workspace_layout(F, N, out_cap, &L);
  if (workspace_bytes < L.total) { pwb_set_error("discrete_molecules: workspace too small"); return 1; }
  if (N >= (1 << 20)) { pwb_set_error("discrete_molecules: too many atoms per cell"); return 1; }
  if (rebuild && !lattice) { pwb_set_error("discrete_molecules: rebuild needs a lattice"); return 1; }
  char* p = (char*)workspace;
  DmArgs A;
  A.F = F; A.N = N; A.rebuild = rebuild; A.periodic = lattice != nullptr; A.n_types = n_types;
  A.out_cap = out_cap; A.mol_cap = mol_cap; A.tol = tol;
  A.coords = coords; A.lattice = lattice; A.lattice_shared = lattice_shared;
  A.atom_type = atom_type; A.type_cov = type_cov; A.type_mass = type_mass; A.type_terminal = type_terminal;
  A.frame_d = (double*)(p + L.frame_d); A.frac = (double*)(p + L.frac); A.xr = (double*)(p + L.xr);
  A.adj = (unsigned*)(p + L.adj); A.deg = (int*)(p + L.deg); A.vis = (unsigned*)(p + L.vis);
  A.rem = (unsigned char*)(p + L.rem); A.front = (int*)(p + L.front);
  A.seed_d = (double*)(p + dm_seed_offset(L, F, out_cap));
  A.n_mol = n_mol; A.mol_off = mol_off; A.out_src = out_src; A.out_xyz = out_xyz; A.status = status;
  k_dm_prepare<<<F, 128, 0, s>>>(A);
  const size_t cl_smem = DM_CL_SMEM(N);
  if (rebuild && lattice && cl_smem <= 110 * 1024) {       // two CTAs per SM; larger cells take the all-pairs kernel
    if (cl_smem > 48 * 1024) cudaFuncSetAttribute(k_dm_bonds_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_smem);
    k_dm_bonds_cells<<<F, DM_CL_THREADS, cl_smem, s>>>(A, 2 * max_cov + tol);
  } else {
    dim3 g((N + DM_TILE - 1) / DM_TILE, F);
    k_dm_bonds<<<g, DM_TILE, 0, s>>>(A, 2 * max_cov + tol);
  }
  {
    // visited bits + "unassigned" bytes of the frame in shared memory while four CTAs still fit on an SM
    const size_t with_state = sizeof(unsigned) * DM_FILL_WARPS * dm_fill_warp_words(N, rebuild, true);
    const size_t without = sizeof(unsigned) * DM_FILL_WARPS * dm_fill_warp_words(N, rebuild, false);
    const int grid = (F + DM_FILL_WARPS - 1) / DM_FILL_WARPS;
    if (with_state <= DM_FILL_STATE_SMEM_MAX) {
      if (with_state > 48 * 1024)
        cudaFuncSetAttribute(k_dm_fill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_state);
      k_dm_fill<true><<<grid, DM_FILL_WARPS * 32, with_state, s>>>(A);
    } else {
      k_dm_fill<false><<<grid, DM_FILL_WARPS * 32, without, s>>>(A);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pwb_set_error(cudaGetErrorString(e)); return 1; }
  return 0;
}

// Pack the per-frame outputs of pwb_discrete_molecules into one dense batch for pwb_full_analysis_batch
// (MolecularSystem.make_modular + the per-molecule loop of trajectory.py:512-522 for a whole batch of frames):
// item_off [items+1], coords [atoms][3], types [atoms] (per-atom type, i.e. shared_types = 0), and the (frame,
// molecule) key of every item.  Device pointers; the dense buffers must hold F*mol_cap + 1 / F*out_cap entries
// (or whatever bound the caller can prove); `scratch` = 2*F + 4 ints.  counts_host[4] (host memory) receives
// {items, atoms, largest molecule, OR of the status words}; the call synchronises `stream` once to deliver them --
// the analysis launch needs the batch size on the host.
int pwb_dm_compact(int F, int N, int out_cap, int mol_cap, const int* n_mol, const int* mol_off, const int* out_src,
                   const double* out_xyz, const int* status, const int* atom_type, int* item_off, double* coords,
                   int* types, int* item_frame, int* item_mol, int* scratch, int* counts_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (F <= 0) { counts_host[0] = counts_host[1] = counts_host[2] = counts_host[3] = 0; return 0; }
  DmPack P;
  P.F = F; P.N = N; P.out_cap = out_cap; P.mol_cap = mol_cap;
  P.n_mol = n_mol; P.mol_off = mol_off; P.out_src = out_src; P.out_xyz = out_xyz; P.status = status;
  P.atom_type = atom_type; P.base = scratch; P.counts = scratch + 2 * (size_t)F;
  P.item_off = item_off; P.coords = coords; P.types = types; P.item_frame = item_frame; P.item_mol = item_mol;
  k_dm_scan<<<1, 1024, 0, s>>>(P);
  k_dm_emit<<<F, 128, 0, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts_host, P.counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { pwb_set_error(cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // extern "C"
