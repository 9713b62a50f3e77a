"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the pywindow pore-geometry hot path.

A numpy restatement of the reference's algorithm (reference = /root/reference/src/pywindow/
_internal/utilities.py, cited as ``U:line``).  It is the *checker* for the CUDA path and the
``cpu_baseline`` leg of ``bench.py``; nothing in ``pywindow_b200/`` may import it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
(a) the reference's own test goldens (tests/test_validate_cc3.py, test_validate_windows.py,
test_validate_average_diameter.py, test_molecular.py) and (b) per-stage outputs of the reference
itself run in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).

Third-party arithmetic on the path (not vendored in the reference; uv.lock pins scipy 1.15.2 /
scikit-learn 1.6.1, the image has scipy 1.18.1 / scikit-learn 1.9.0):
* ``scipy.optimize.minimize`` (L-BFGS-B, 2-point forward differences) -- called here exactly as
  the reference calls it (U:422, U:1301) because its stopping point *is* the reference result;
* ``scipy.optimize.brute`` + ``fmin`` (U:1312) -- called the same way;
* ``sklearn.metrics.pairwise.euclidean_distances`` -- restated as the direct difference formula
  (differs from sklearn's ||x||^2-2xy+||y||^2 expansion by ~1e-15 A);
* ``sklearn.cluster.DBSCAN`` (U:1478) -- restated by the order-free rule of SURVEY.md 8(a) row 9
  and pinned against sklearn's labels stored in the goldens;
* ``sklearn.neighbors.KDTree.query(k=10)`` (U:1428) -- restated as brute-force k-NN.

The ray/path stages are evaluated batched (all rays x all atoms at once) instead of one Python
call per ray; results are identical, it is just not slow.
"""
from __future__ import annotations

import csv
import math
import os

import numpy as np
from scipy.optimize import brute, fmin, minimize

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pywindow_b200", "data")


def _load_tables():
    mass, vdw, cov = {}, {}, {}
    with open(os.path.join(_DATA, "elements.csv")) as f:
        for row in csv.DictReader(f):
            k = row["symbol"]
            if row["mass"]:
                mass[k] = float(row["mass"])
            if row["vdw_radius"]:
                vdw[k] = float(row["vdw_radius"])
            if row["covalent_radius"]:
                cov[k] = float(row["covalent_radius"])
    return mass, vdw, cov


ATOMIC_MASS, ATOMIC_VDW, ATOMIC_COV = _load_tables()


def masses_of(elements):
    return np.array([ATOMIC_MASS[e.upper()] for e in elements], dtype=np.float64)


def vdw_of(elements):
    return np.array([ATOMIC_VDW[e.upper()] for e in elements], dtype=np.float64)


# ----------------------------------------------------------------------------- scalars
def center_of_mass(mass, X):
    """U:127-148 -- sum(m_i x_i)/sum(m_i)."""
    return (X * mass[:, None]).sum(axis=0) / np.array([mass.sum()] * 3)


def shift_com(mass, X, com_adjust=np.zeros(3)):
    """U:344-352 -- X - (COM - com_adjust)."""
    return X - (center_of_mass(mass, X) - com_adjust)


def max_dim(r, X):
    """U:355-372 -- argmax over the upper triangle (diagonal included) of |xi-xj|+ri+rj,
    first occurrence in row-major order."""
    diff = X[:, None, :] - X[None, :, :]
    m = np.sqrt((diff * diff).sum(-1)) + r[:, None] + r[None, :]
    m = np.triu(m)
    i1, i2 = np.unravel_index(m.argmax(), m.shape)
    return int(i1), int(i2), float(m[i1, i2])


def clearances(r, X, c):
    d = X - c
    return np.sqrt((d * d).sum(-1)) - r


def pore_diameter(r, X, c):
    """U:375-388 -- 2*min_i(|x_i-c|-r_i) and first argmin."""
    a = clearances(r, X, c)
    i = int(np.argmin(a))
    return float(a[i] * 2), i


def opt_pore_diameter(r, X, com, bounds=None):
    """U:400-426 -- L-BFGS-B (scipy defaults) on -pore_diameter from ``com`` inside ``bounds`` (default: the box
    com +- pore_r, pore_r measured at ``com``)."""
    if bounds is None:
        pore_r = pore_diameter(r, X, com)[0] / 2
        bounds = tuple((com[k] - pore_r, com[k] + pore_r) for k in range(3))
    res = minimize(lambda c: -pore_diameter(r, X, c)[0], x0=com, bounds=bounds)
    d, i = pore_diameter(r, X, res.x)
    return d, i, res.x


def sphere_volume(radius):
    """U:429-431."""
    return float(4 / 3 * np.pi * radius ** 3)


# ----------------------------------------------------------------------------- ray set
def n_sampling_points(radius, adjust=1.0):
    """U:1399-1409."""
    return int(np.log10(4 * np.pi * radius ** 2) * 250 * adjust)


def spiral_points(P, radius):
    """U:1412-1423 -- golden-spiral points scaled to ``radius``."""
    golden = np.pi * (3 - np.sqrt(5))
    theta = golden * np.arange(P)
    z = np.linspace(1 - 1.0 / P, 1.0 / P - 1.0, P)
    rad = np.sqrt(1 - z * z)
    pts = np.zeros((P, 3))
    pts[:, 0] = rad * np.cos(theta) * radius
    pts[:, 1] = rad * np.sin(theta) * radius
    pts[:, 2] = z * radius
    return pts


def knn_eps(points, k=10, block=2048):
    """U:1427-1434 -- eps = m + sqrt(m), m = mean of the k smallest distances (self included)
    over all points."""
    P = points.shape[0]
    tot = 0.0
    for s in range(0, P, block):
        d = points[s:s + block, None, :] - points[None, :, :]
        dist = np.sqrt((d * d).sum(-1))
        tot += np.partition(dist, k - 1, axis=1)[:, :k].sum()
    m = tot / (P * k)
    return m + m ** 0.5


def py_floordiv(a, b):
    """Python/numpy float ``//`` (fmod based; U:1108 uses ``norm // increment``)."""
    return float(np.float64(a) // np.float64(b))


# ----------------------------------------------------------------------------- ray stage
def ray_pretest(r, X, points):
    """U:1132-1161 (and the identical head of U:1556-1577).

    Returns ``(blocked[P] bool, t_exit[P,N] with -inf where the atom does not count)``.
    Line through the centroid o=mean(X) along n=p/|p|; atom i counts iff r_i^2-d_i^2>0 and
    |o+(t_i-h_i)n| < |o+(t_i+h_i)n| (norms about the true origin)."""
    nrm = np.sqrt((points * points).sum(-1))
    n = points / nrm[:, None]
    o = X.sum(axis=0) / X.shape[0]
    L = X - o
    t = n @ L.T                                   # [P,N]
    l2 = (L * L).sum(-1)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(l2[None, :] - t * t)          # nan when slightly negative, like the reference
        under = r[None, :] ** 2 - d ** 2
        hit = under > 0
        h = np.sqrt(np.where(hit, under, 0.0))
    p0 = o[None, None, :] + (t - h)[..., None] * n[:, None, :]
    p1 = o[None, None, :] + (t + h)[..., None] * n[:, None, :]
    n0 = np.sqrt((p0 * p0).sum(-1))
    n1 = np.sqrt((p1 * p1).sum(-1))
    counts = hit & (n0 < n1)
    return counts.any(axis=1), np.where(counts, n1, -np.inf)


def path_scan(r, X, vec, increment):
    """U:1100-1129 for one ray.  Returns the 8-vector or None."""
    chunks = int(py_floordiv(np.sqrt((vec * vec).sum()), increment))
    chunk = vec / chunks
    path = chunk[None, :] * np.arange(chunks + 1)[:, None]
    d = X[None, :, :] - path[:, None, :]
    a = (np.sqrt((d * d).sum(-1)) - r[None, :]).min(axis=1)
    if (a > 0).all():
        pos = int(np.argmin(a))
        q = chunk * pos
        return np.array([np.sqrt((q * q).sum()), a[pos] * 2, *q, *vec])
    return None


def ray_stage(r, X, points, increment=1.0):
    """The whole HOT LOOP 1 of find_windows (U:1441-1470): pre-test then stepped path
    clearance for survivors.  Returns ``(clear[P] bool, records[S,8])``."""
    blocked, _ = ray_pretest(r, X, points)
    clear = np.zeros(points.shape[0], dtype=bool)
    recs = []
    for k in np.nonzero(~blocked)[0]:
        rec = path_scan(r, X, points[k], increment)
        if rec is not None:
            clear[k] = True
            recs.append(rec)
    return clear, np.array(recs).reshape(-1, 8)


# ----------------------------------------------------------------------------- clustering
def dbscan_labels(pts, eps, min_samples=5):
    """sklearn DBSCAN label semantics (U:1478; sklearn/cluster/_dbscan_inner.pyx) restated
    order-free: core = (#points with dist<=eps, self included) >= min_samples; clusters =
    connected components of the core-core graph numbered by their smallest core index;
    a border point takes the smallest cluster number among its core neighbours; else -1."""
    S = pts.shape[0]
    d = pts[:, None, :] - pts[None, :, :]
    adj = np.sqrt((d * d).sum(-1)) <= eps
    core = adj.sum(axis=1) >= min_samples
    comp = np.where(core, np.arange(S), S)
    cadj = adj & core[None, :]
    while True:                                   # min-label propagation over core points
        nb = np.where(cadj, comp[None, :], S).min(axis=1)
        new = np.where(core, np.minimum(comp, nb), S)
        if (new == comp).all():
            break
        comp = new
    roots = np.unique(comp[core])
    rank = {int(v): i for i, v in enumerate(roots)}
    labels = np.full(S, -1, dtype=np.int64)
    for i in range(S):
        if core[i]:
            labels[i] = rank[int(comp[i])]
        else:
            nb = comp[cadj[i]]
            if nb.size:
                labels[i] = min(rank[int(v)] for v in nb)
    return labels


# ----------------------------------------------------------------------------- windows
def _angle(x, y):
    """U:1088-1097."""
    c = abs(x[0] * y[0] + x[1] * y[1] + x[2] * y[2]) / (
        np.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2) * np.sqrt(y[0] ** 2 + y[1] ** 2 + y[2] ** 2))
    return np.arccos(c)


def window_angles(v):
    """Octant table U:1239-1259."""
    a1 = _angle(np.array([v[0], v[1], 0.0]), np.array([1.0, 0, 0]))
    a2 = _angle(v, np.array([0, 0, 1.0]))
    px, py, pz = v[0] >= 0, v[1] >= 0, v[2] >= 0
    if px and py and pz:
        a1, a2 = -a1, -a2
    if (not px) and py and pz:
        a1 = np.pi * 2 + a1
    if px and (not py) and pz:
        a2 = -a2
    if (not px) and (not py) and pz:
        a1 = np.pi * 2 - a1
    if px and py and (not pz):
        a1, a2 = -a1, np.pi + a2
    if (not px) and py and (not pz):
        a2 = np.pi - a2
    if px and (not py) and (not pz):
        a2 = a2 + np.pi
    if (not px) and (not py) and (not pz):
        a1, a2 = -a1, np.pi - a2
    return a1, a2


def _rot_z(a):
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])


def _rot_y(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])


def window_analysis(rows, r, X, increment2=0.1, detail=None):
    """U:1191-1361 for one cluster (``rows`` = its [k,8] ray records, ``X`` shifted coords)."""
    vec0 = rows[rows.argmax(axis=0)[1]][5:8]
    va = path_scan(r, X, vec0, increment2)
    if detail is not None:
        detail["rescan"] = va
    if va is None:
        return None
    v = va[5:8]
    a1, a2 = window_angles(v)
    Y = X @ _rot_z(a1).T
    Y = Y @ _rot_y(a2).T
    new_z = va[0]
    Y = Y - np.array([0, 0, new_z])
    d0, _ = pore_diameter(r, Y, np.zeros(3))
    zres = minimize(lambda z: pore_diameter(r, Y, np.array([0.0, 0.0, z[0]]))[0],
                    x0=0.0, bounds=[[-new_z, None]])
    wz = zres.x[0]
    b = ((-d0 / 2, d0 / 2), (-d0 / 2, d0 / 2))
    xy = brute(lambda q: -pore_diameter(r, Y, np.array([q[0], q[1], wz]))[0], b,
               full_output=True, finish=fmin)
    wc = np.array([xy[0][0], xy[0][1], wz])
    diam, _ = pore_diameter(r, Y, wc)
    if detail is not None:
        detail.update(z_opt=wz, xy_opt=np.array(xy[0]), d0=d0, grid_argmin=int(xy[3].ravel().argmin()),
                      coords_rot=Y, angles=(a1, a2), new_z=new_z)
    wc[2] = wc[2] + new_z
    wc = _rot_y(-a2) @ wc
    wc = _rot_z(-a1) @ wc
    return diam, wc


def find_windows(r, mass, X, adjust=1.0, pore_opt=True, increment=1.0, centre=None, detail=None):
    """U:1364-1553.  ``centre`` (extension) forces the pore centre instead of optimising it
    (teacher forcing).  Returns None | (diameters[W], coms[W,3])."""
    com0 = center_of_mass(mass, X)
    if pore_opt:
        c = opt_pore_diameter(r, X, com0)[2] if centre is None else np.asarray(centre, dtype=np.float64)
        com_adjust = com0 - c
        initial_com = com0 - com_adjust
        Xs = shift_com(mass, X, com_adjust)
    else:
        initial_com = com0
        Xs = shift_com(mass, X)
    R = max_dim(r, Xs)[2] / 2
    P = n_sampling_points(R, adjust)
    pts = spiral_points(P, R)
    eps = knn_eps(pts)
    clear, recs = ray_stage(r, Xs, pts, increment)
    if detail is not None:
        detail.update(shifted=Xs, R=R, P=P, eps=eps, clear=clear, records=recs)
    if recs.shape[0] == 0:
        return None
    labels = dbscan_labels(recs[:, 5:8], eps)
    if detail is not None:
        detail["labels"] = labels
        detail["windows"] = []
    diam, coms = [], []
    for lab in sorted(set(labels.tolist()) - {-1}):
        wd = {} if detail is not None else None
        res = window_analysis(recs[labels == lab], r, Xs, detail=wd)
        if detail is not None:
            detail["windows"].append(wd)
        if res is not None:
            diam.append(res[0])
            coms.append(res[1] + initial_com)
    return np.array(diam), np.array(coms)


# ----------------------------------------------------------------------------- average diameter
def find_average_diameter(r, mass, X, adjust=1.0):
    """U:1586-1650 + U:1556-1583: sphere radius = max_dim (not /2); per ray the farthest exit
    point among counting atoms; 2*mean over rays that hit anything."""
    Xs = shift_com(mass, X)
    R = max_dim(r, Xs)[2]
    P = n_sampling_points(R, adjust)
    pts = spiral_points(P, R)
    out = []
    for s in range(0, P, 4096):
        blocked, ex = ray_pretest(r, Xs, pts[s:s + 4096])
        out.append(ex.max(axis=1)[blocked])
    return float(np.mean(np.concatenate(out)) * 2)


# ----------------------------------------------------------------------------- full analysis
def full_analysis(elements, X, centre=None):
    """Molecule.full_analysis (molecular.py:156-202) -> the same ``properties`` dict."""
    X = np.asarray(X, dtype=np.float64)
    r, mass = vdw_of(elements), masses_of(elements)
    com = center_of_mass(mass, X)
    i1, i2, D = max_dim(r, X)
    pd_, pa = pore_diameter(r, X, com)
    od, oa, oc = opt_pore_diameter(r, X, com)
    props = {
        "no_of_atoms": len(elements),
        "centre_of_mass": com,
        "maximum_diameter": {"diameter": D, "atom_1": i1, "atom_2": i2},
        "average_diameter": find_average_diameter(r, mass, X),
        "pore_diameter": {"diameter": pd_, "atom": pa},
        "pore_volume": sphere_volume(pd_ / 2),
        "pore_diameter_opt": {"diameter": od, "atom_1": oa, "centre_of_mass": oc},
        "pore_volume_opt": sphere_volume(od / 2),
    }
    w = find_windows(r, mass, X, centre=oc if centre is None else centre)
    props["windows"] = ({"diameters": None, "centre_of_mass": None} if w is None
                        else {"diameters": w[0], "centre_of_mass": w[1]})
    return props
