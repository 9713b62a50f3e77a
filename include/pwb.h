/* pwb.h -- C ABI of libpwb.so: B200 (sm_100a) kernels for pywindow's per-frame pore-geometry path.
 *
 * The reference (marcinmiklitz/pywindow) is pure Python and has no FFI layer; the seam this library
 * plugs into is the Python function boundary between `pywindow/_internal/molecular.py` (callers) and
 * `pywindow/_internal/utilities.py` (geometry functions), see molecular.py:29-44, and the per-frame
 * loop of `Trajectory.analysis` (trajectory.py:473-586).  Each entry point below names the
 * reference function(s) it replaces.  All pointers in `pwb_batch` are DEVICE pointers unless the
 * entry point says "host"; the caller owns every buffer (inputs, workspace, outputs); the library
 * keeps no global state except the explicit context.  Every call is asynchronous on `stream`
 * except the `_host` variant, which synchronises the stream before returning.
 *
 * A "batch" is F independent items (one molecule in one frame each).  Item i owns atoms
 * [item_off[i], item_off[i+1]) of `coords`.  Atoms carry a small integer *type* (one per element);
 * per-type van-der-Waals radius / mass tables replace the reference's per-atom dict lookups
 * (tables.py:22-197).  If `shared_types` != 0 every item has the same atom order and
 * `atom_type` has max_atoms entries (trajectory of one molecule); otherwise it is indexed like
 * `coords` (one entry per atom of the batch).
 */
#ifndef PWB_H
#define PWB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWB_ABI_VERSION 4
#define PWB_MAX_TYPES 16

/* ---- result record: `rec_stride` doubles per item (>= PWB_REC_WIN0 + 4*win_cap) --------------- */
enum {
  PWB_REC_STATUS = 0,    /* bit flags PWB_ST_* stored as a double                                */
  PWB_REC_NATOMS = 1,    /* properties["no_of_atoms"]                         molecular.py:131   */
  PWB_REC_COM = 2,       /* [3] centre_of_mass                                utilities.py:127   */
  PWB_REC_MAXD = 5,      /* maximum_diameter.diameter                         utilities.py:355   */
  PWB_REC_MAXD_A1 = 6,   /* maximum_diameter.atom_1                                              */
  PWB_REC_MAXD_A2 = 7,   /* maximum_diameter.atom_2                                              */
  PWB_REC_AVGD = 8,      /* average_diameter                                  utilities.py:1586  */
  PWB_REC_PORE_D = 9,    /* pore_diameter.diameter                            utilities.py:375   */
  PWB_REC_PORE_ATOM = 10,
  PWB_REC_OPT_D = 11,    /* pore_diameter_opt.diameter                        utilities.py:400   */
  PWB_REC_OPT_ATOM = 12,
  PWB_REC_OPT_C = 13,    /* [3] pore_diameter_opt.centre_of_mass                                 */
  PWB_REC_NRAYS = 16,    /* number of sampling rays P                         utilities.py:1409  */
  PWB_REC_NCLEAR = 17,   /* rays that reach the sampling sphere unobstructed                     */
  PWB_REC_NWIN = 18,     /* clusters (= window slots used)                    utilities.py:1478  */
  PWB_REC_EPS = 19,      /* DBSCAN eps                                        utilities.py:1434  */
  PWB_REC_OPT_NIT = 20,  /* L-BFGS-B iterations of the pore-centre search                        */
  PWB_REC_OPT_STOP = 21, /* its stop reason (0 pgtol, 1 ftol, 2 line search, 3 maxiter)          */
  PWB_REC_RADIUS = 22,   /* sampling-sphere radius used for windows                              */
  PWB_REC_NSURV = 23,    /* rays that survive the line-sphere pre-test                           */
  PWB_REC_OPT_NFEV = 24, /* objective evaluations of the pore-centre search (4 per gradient)     */
  PWB_REC_WIN0 = 28      /* then win_cap x {diameter, cx, cy, cz}; diameter NaN = window dropped */
};

enum {
  PWB_ST_OPT_BOUNDS = 1,     /* pore_diameter(COM) < 0: reference raises ValueError (scipy bounds)  */
  PWB_ST_NO_CLEAR_RAY = 2,   /* find_windows returns None                      utilities.py:1473    */
  PWB_ST_RAY_CAP = 4,        /* P > ray_cap: windows not analysed, enlarge ray_cap                  */
  PWB_ST_WIN_CAP = 8,        /* more clusters than win_cap                                          */
  PWB_ST_WIN_DROPPED = 16,   /* a window's 0.1 A rescan was obstructed         utilities.py:1226    */
  PWB_ST_WIN_NEGATIVE = 32,  /* a window diameter < 0 (reference warns)        utilities.py:1546    */
  PWB_ST_BAD_ITEM = 64       /* too many atoms / unknown type                                        */
};

/* stages (bit mask for pwb_batch.stages) */
enum {
  PWB_STAGE_SCALARS = 1,  /* COM, max_dim, pore_diameter                                          */
  PWB_STAGE_AVGD = 2,     /* find_average_diameter                                                */
  PWB_STAGE_OPT = 4,      /* opt_pore_diameter                                                    */
  PWB_STAGE_RAYS = 8,     /* ray pre-test + path clearance (HOT LOOP 1 of find_windows)           */
  PWB_STAGE_CLUSTER = 16, /* eps + DBSCAN + per-cluster arg-max ray                               */
  PWB_STAGE_WINDOWS = 32, /* window_analysis per cluster (HOT LOOP 2)                             */
  PWB_STAGE_ALL = 63
};

typedef struct pwb_ctx pwb_ctx; /* per-device context: k-NN/eps table for the golden-spiral ray sets */

typedef struct pwb_batch {
  /* sizes */
  int32_t n_items;      /* F                                                                      */
  int32_t max_atoms;    /* max atoms of any item (sizes shared memory)                            */
  int32_t n_types;      /* <= PWB_MAX_TYPES                                                       */
  int32_t shared_types; /* see above                                                              */
  int32_t ray_cap;      /* capacity of per-item ray lists (>= largest P)                          */
  int32_t win_cap;      /* window slots per record                                                */
  int32_t rec_stride;   /* doubles per record                                                     */
  int32_t stages;       /* PWB_STAGE_* mask                                                       */
  /* parameters of find_windows / find_average_diameter (utilities.py:1364-1371, :1586-1591)      */
  double adjust;        /* default 1.0                                                            */
  double increment;     /* default 1.0                                                            */
  double increment2;    /* default 0.1 (window_analysis)                                          */
  int32_t pore_opt;     /* default 1                                                              */
  int32_t total_atoms;  /* atoms in the whole batch = item_off[F]                               */
  /* inputs */
  const double* coords;      /* [total_atoms][3] float64                                          */
  const int32_t* item_off;   /* [F+1]                                                             */
  const int32_t* atom_type;  /* see shared_types                                                  */
  const double* type_vdw;    /* [n_types]                                                         */
  const double* type_mass;   /* [n_types]                                                         */
  const double* centre_override; /* optional [F][3]: use instead of optimising (teacher forcing)  */
  /* opt_pore_diameter(bounds=, com=) (utilities.py:400-426); both optional (NULL = reference defaults) */
  const double* opt_start;   /* [F][3] start point x0 instead of the centre of mass                */
  const double* opt_bounds;  /* [F][3][2] (lo, hi) per axis, +-inf = unbounded, instead of x0 +- pore radius */
  /* outputs / scratch */
  double* records;           /* [F][rec_stride]                                                   */
  void* workspace;           /* pwb_workspace_bytes() bytes                                       */
  size_t workspace_bytes;
  /* Ray sharding of ONE huge item over several GPUs (BASELINE config 5, >= 1e5 rays): pwb_ray_search only
   * analyses the ray tiles [ray_tile_begin, ray_tile_end) (tiles of PWB_RAY_TILE = 1024 consecutive rays of the
   * golden spiral, utilities.py:1441-1470) and leaves that range's clear rays packed at the front of the lists;
   * the caller gathers the ranges of all GPUs in rank order, writes the merged lists and counts back and runs
   * pwb_cluster / pwb_window_refine on them.  0, 0 = every tile (single GPU).                               */
  int32_t ray_tile_begin, ray_tile_end;
} pwb_batch;
#define PWB_RAY_TILE 1024

/* layout of the per-item lists inside the workspace, for stage-wise tests */
typedef struct pwb_ws_layout {
  size_t atoms;      /* double[total_atoms_cap][4]: type-bucketed x,y,z,r_vdw                     */
  size_t atom_idx;   /* int32[total_atoms_cap]: original index of each bucketed atom              */
  size_t item_d;     /* double[F][16]                                                             */
  size_t item_i;     /* int32[F][32]                                                              */
  size_t clear_k;    /* int32[F][ray_cap]: ray index of each clear ray (ascending)                */
  size_t clear_pos;  /* int32[F][ray_cap]: path step of the narrowest point                       */
  size_t clear_a2;   /* double[F][ray_cap]: 2*min clearance along the ray (column 1 of the record)*/
  size_t labels;     /* int32[F][ray_cap]: DBSCAN label of each clear ray                         */
  size_t win_k;      /* int32[F][win_cap]: arg-max ray of each cluster                            */
  size_t big_pts;    /* double[F][ray_cap][4]: unit ray points, only if ray_cap > 4096 (config 5) */
  size_t big_i;      /* int32[F][2][ray_cap]: union-find scratch of the large-ray-count path      */
  size_t atom_scratch; /* per-CTA / per-warp atom arrays of molecules too large for shared memory
                          (> ~2 900 atoms; empty otherwise)                                        */
  size_t opt_hist;   /* double[F][60]: (s, y) rings of the pore-centre searches (k_opt_pore)       */
  size_t total;
} pwb_ws_layout;

int pwb_abi_version(void);
const char* pwb_last_error(void);

int pwb_ctx_create(pwb_ctx** out, int device, int ray_table_cap);
int pwb_ctx_destroy(pwb_ctx* ctx);
/* Cut batches into `chunks` (1..8) sub-batches pipelined over internal streams (default 1).      */
int pwb_ctx_set_chunks(pwb_ctx* ctx, int chunks);
/* Test switches.  PWB_OPTION_OPT_NO_LIST = 1: the pore-centre search (pwb_opt_pore) evaluates its objective over every
 * atom at every point instead of over its survivor lists -- the results must be bit-identical either way
 * (tests/test_gpu_parity.py).                                                                                   */
#define PWB_OPTION_OPT_NO_LIST 1
int pwb_ctx_set_option(pwb_ctx* ctx, int option, int value);

/* bytes of workspace for a batch of n_items with total_atoms atoms, the largest item having max_atoms */
size_t pwb_workspace_bytes(int n_items, int total_atoms, int max_atoms, int ray_cap, int win_cap);
int pwb_workspace_layout(int n_items, int total_atoms, int max_atoms, int ray_cap, int win_cap, pwb_ws_layout* out);

/* upper bound of the ray count any item of the batch can need (bounding-box bound on max_dim);
 * synchronises `stream`.  Replaces nothing in the reference: it sizes `ray_cap`.                  */
int pwb_ray_capacity(pwb_ctx* ctx, const pwb_batch* b, void* stream, int* out_cap);

/* stage entry points; each runs the stage named in its comment on the whole batch.                */
int pwb_frame_scalars(pwb_ctx*, const pwb_batch*, void* stream);    /* center_of_mass :127, max_dim :355, pore_diameter :375 */
int pwb_average_diameter(pwb_ctx*, const pwb_batch*, void* stream); /* find_average_diameter :1586 + vector_analysis_reversed :1556 */
int pwb_opt_pore(pwb_ctx*, const pwb_batch*, void* stream);         /* opt_pore_diameter :400 */
int pwb_ray_search(pwb_ctx*, const pwb_batch*, void* stream);       /* vector_preanalysis :1132 + vector_analysis :1100 over the ray set :1398-1423 */
int pwb_cluster(pwb_ctx*, const pwb_batch*, void* stream);          /* eps :1427-1434 + DBSCAN :1478 + arg-max ray :1221 */
int pwb_window_refine(pwb_ctx*, const pwb_batch*, void* stream);    /* window_analysis :1191 */

/* Molecule.full_analysis (molecular.py:156-202) for every item: runs the stages in b->stages.     */
int pwb_full_analysis_batch(pwb_ctx*, const pwb_batch*, void* stream);

/* Same, but `coords` and `records` are HOST pointers (pinned or pageable); `d_coords` is a device
 * staging buffer of total_atoms*3 doubles and `d_records` one of F*rec_stride doubles (caller
 * owned).  Copies in, analyses, copies the records out and synchronises `stream`.  This is what
 * the per-frame loop of Trajectory.analysis (trajectory.py:473-586) binds to.                      */
int pwb_full_analysis_host(pwb_ctx*, const pwb_batch* b, const double* h_coords, double* h_records,
                           double* d_coords, double* d_records, int total_atoms, void* stream);
/* Same without the final synchronisation (h_coords / h_records must be pinned): the records are valid once
 * `stream` has been synchronised.  `d_coords`, `d_records` and the workspace must be idle when the call is made
 * (no earlier queued work still uses them): the input copies go to an internal copy stream and the sub-batches to
 * the internal streams WITHOUT waiting for work already queued on `stream`, so with two sets of buffers used
 * alternately the copies of batch i+1 overlap the kernels of batch i and the tail of batch i overlaps the head of
 * batch i+1 -- which is how a trajectory reader that decodes batch i+1 while batch i is analysed uses it
 * (pywindow_b200.engine.FramePipeline).                                                              */
int pwb_full_analysis_host_async(pwb_ctx*, const pwb_batch* b, const double* h_coords, double* h_records,
                                 double* d_coords, double* d_records, int total_atoms, void* stream);

/* ---- periodic cells -------------------------------------------------------------------------
 * discrete_molecules (utilities.py:820-1085) for F cell-frames of N atoms each; with `rebuild` the
 * 27-image supercell of create_supercell (utilities.py:768-810) is implied and molecules are rebuilt
 * across the periodic boundary and filtered by their centre of mass (utilities.py:1062-1077).
 * Device pointers.  coords [F][N][3]; lattice [F][9] row-major (one matrix if lattice_shared, NULL =
 * non-periodic system); atom_type [N] with per-type covalent radius / mass / "terminal element" flag
 * (H, halogens, noble gases never expand, utilities.py:933); max_cov = largest covalent radius
 * present.  Outputs per frame f: n_mol[f] molecules; molecule k owns output atoms
 * [mol_off[f][k], mol_off[f][k+1]) whose source atom index is out_src[f][.] and whose (rebuilt,
 * 8-decimal rounded) coordinates are out_xyz[f][.][3], in the reference's atom order.
 * status[f]: 1 bond-table overflow, 2 output overflow (raise out_cap), 4 molecule overflow.          */
typedef struct pwb_dm_layout { size_t frame_d, frac, xr, adj, deg, vis, rem, front, total; } pwb_dm_layout;
size_t pwb_dm_workspace_bytes(int n_frames, int n_atoms, int out_cap);
int pwb_dm_workspace_layout(int n_frames, int n_atoms, int out_cap, pwb_dm_layout* out);
int pwb_discrete_molecules(int n_frames, int n_atoms, int rebuild, double tol, double max_cov,
                           const double* coords, const double* lattice, int lattice_shared,
                           const int32_t* atom_type, int n_types, const double* type_cov,
                           const double* type_mass, const int32_t* type_terminal, int out_cap, int mol_cap,
                           int32_t* n_mol, int32_t* mol_off, int32_t* out_src, double* out_xyz,
                           int32_t* status, void* workspace, size_t workspace_bytes, void* stream);
/* Pack those per-frame outputs into ONE dense batch for pwb_full_analysis_batch (make_modular + the per-molecule
 * loop of trajectory.py:512-522 for a whole batch of frames): item_off [items+1], coords [atoms][3], per-atom types
 * (shared_types = 0) and the (frame, molecule) key of every item.  Device pointers sized for the worst case the
 * caller allows (items <= F*mol_cap, atoms <= F*out_cap); scratch: 2*F + 4 ints.  counts_host[4] (HOST) receives
 * {items, atoms, largest molecule, OR of status}; synchronises `stream` once to deliver them.                */
int pwb_dm_compact(int n_frames, int n_atoms, int out_cap, int mol_cap, const int32_t* n_mol, const int32_t* mol_off,
                   const int32_t* out_src, const double* out_xyz, const int32_t* status, const int32_t* atom_type,
                   int32_t* item_off, double* coords, int32_t* types, int32_t* item_frame, int32_t* item_mol,
                   int32_t* scratch, int32_t* counts_host, void* stream);
void pwb_set_error(const char* msg);

/* ---- host side: DL_POLY HISTORY text -> [F][N][3] float64 (replaces the per-frame Python decode of
 * trajectory.py:217-249, :647-689, :712-766).  HOST pointers; multi-threaded; numbers are converted
 * exactly like Python's float().                                                                    */
int64_t pwb_history_index(const char* buf, int64_t len, int64_t* starts, int64_t cap);
int64_t pwb_history_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n,
                           int natoms, int keytrj, int imcon, double* xyz, double* lattice, int threads);

/* Multi-frame XYZ (trajectory.py:792-915) and multi-model PDB (trajectory.py:917-1045) trajectories, same
 * contract as the HISTORY pair: pwb_xyz_index maps the frames (byte range + atom count each; returns the number
 * of frames, or -(1 + byte offset) of a line that should hold an atom count), pwb_xyz_decode / pwb_pdb_decode
 * parse n frames of `natoms` atoms on `threads` host threads (0 = all) into xyz[n][natoms][3], bit-identical to
 * Python's float(); they return 0 or 1 + the index of the first malformed frame.                          */
int64_t pwb_xyz_index(const char* buf, int64_t len, int64_t* starts, int64_t* ends, int32_t* natoms, int64_t cap);
int64_t pwb_xyz_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n,
                       int natoms, double* xyz, int threads);
int64_t pwb_pdb_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n,
                       int natoms, double* xyz, int threads);

/* Measurement utility: line feeds of a mapped file counted by `threads` host threads -- the host's plain streaming
 * rate over the file, the ceiling of any text decoder (bench.py: api.host_stream_GBps).                       */
int64_t pwb_stream_probe(const char* buf, int64_t len, int threads);

/* Writer of the same HISTORY layout (keytrj 0; imcon 0 or 3): the synthetic trajectories of bench.py / tests that
 * feed the reference's DLPOLY reader and pwb_history_decode alike.  keys: natoms x 8 blank-padded characters;
 * lattice NULL | [9] | [n_frames][9] row-major (written transposed, as trajectory.py:723-726 reads it); coordinates
 * as %20.10E, or DL_POLY's own %12.4E if compact.  Returns the file size, or -1.                                                                 */
int64_t pwb_history_write(const char* path, const char* title, const char* keys, const double* masses,
                          const double* xyz, int64_t n_frames, int natoms, const double* lattice,
                          int lattice_per_frame, int compact, int threads);

/* Measurement utility: sustained FMA throughput of the FP32 (fp64 = 0) or FP64 (fp64 = 1) pipe in
 * TFLOP/s, timed with CUDA events (the roofline denominator of this compute-bound path).          */
int pwb_fma_peak(pwb_ctx*, int fp64, void* stream, double* tflops);

/* number of kernel launches issued by this library since the context was created */
long long pwb_launch_count(pwb_ctx*);

#ifdef __cplusplus
}
#endif
#endif
