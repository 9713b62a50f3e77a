"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the periodic decomposition (``create_supercell`` U:768-810,
``discrete_molecules`` U:820-1085; U = /root/reference/src/pywindow/_internal/utilities.py).

The reference grows each molecule with value-equality list operations; here the same result is
produced from a bond graph and an *ordered* breadth-first search:

* nodes are supercell atoms (image, atom); image (0,0,0) of an atom *is* the cell atom;
* bond rule U:1004-1010: ``Rcov_i + Rcov_j - tol < d < Rcov_i + Rcov_j + tol`` with the prefilter
  ``0.1 < d < 2*max(Rcov) + tol`` (U:999); terminal elements never expand (U:933, U:990);
* repeat: seed = remaining non-terminal atom closest to the pseudo origin (U:944-953); level
  synchronous flood fill; for every parent in frontier order the candidates are its bonded
  *remaining* cell atoms in ascending index (U:997-1012) followed (rebuild only) by its other bonded
  supercell atoms in ascending supercell index (U:1013-1033); first occurrence wins (U:1041),
  already collected atoms are dropped (U:1042-1044); the frontier is removed from the remaining set
  after the level (U:1037-1039);
* with rebuild a molecule is kept iff its centre of mass, in fractional coordinates rounded to 8
  decimals, lies in [0,1)^3 -- or [-0.5,0.5)^3 if the system COM is within 1 A of the origin
  (U:919-923, U:1062-1077).  Coordinates pass through ``round(x, 8)`` (U:188-216).

Parity status: PINNED against tests/golden/per_*.npz (reference run in the build container) and the
reference's own counts (tests/test_molecular.py:4467-4553: 33 fragments / 8 x 168 atoms).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

from pw_oracle import ATOMIC_COV, ATOMIC_MASS

TERMINAL = {"H", "CL", "BR", "F", "HE", "AR", "NE", "KR", "XE", "RN"}


def unit_cell_to_lattice_array(cryst):
    """U:653-690."""
    a_, b_, c_, alpha, beta, gamma = cryst
    ra, rb, rg = np.deg2rad(alpha), np.deg2rad(beta), np.deg2rad(gamma)
    vol = a_ * b_ * c_ * (1 - np.cos(ra) ** 2 - np.cos(rb) ** 2 - np.cos(rg) ** 2
                          + 2 * np.cos(ra) * np.cos(rb) * np.cos(rg)) ** 0.5
    return np.array([[a_, b_ * np.cos(rg), c_ * np.cos(rb)],
                     [0, b_ * np.sin(rg), c_ * (np.cos(ra) - np.cos(rb) * np.cos(rg)) / np.sin(rg)],
                     [0, 0, vol / (a_ * b_ * np.sin(rg))]])


def sk_distances(A, y):
    """sklearn.metrics.pairwise.euclidean_distances(A, y[None]) restated with its operation order
    (-2 A.y, += |A|^2, += |y|^2, clip at 0, sqrt): on crystal-symmetric cells two seed candidates tie
    to ~1e-14 A and this rounding decides the reference's arg-min (SURVEY.md 8a row 14 caveat)."""
    d = -2 * (A @ y.reshape(-1, 1))
    d += np.einsum("ij,ij->i", A, A)[:, None]
    d += np.einsum("ij,ij->i", y.reshape(1, -1), y.reshape(1, -1))[None, :]
    np.maximum(d, 0, out=d)
    return np.sqrt(d).ravel()


IMAGES = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)])


def supercell_coordinates(X, lattice):
    """U:768-810: fractional = inv(M) x, + the 27 integer shifts (a outermost), back to Cartesian."""
    frac = X @ np.linalg.inv(lattice).T
    return np.concatenate([(frac + s) @ lattice.T for s in IMAGES], axis=0)


def discrete_molecules(elements, X, lattice=None, rebuild=False, tol=0.4):
    """Returns a list of ``(elements, coordinates, source_atom_index)`` in the reference's order."""
    elements = np.asarray(elements)
    N = len(elements)
    up = [e.upper() for e in elements]
    rcov = np.array([ATOMIC_COV[e] for e in up])
    mass = np.array([ATOMIC_MASS[e] for e in up])
    heavy = np.array([e not in TERMINAL for e in up])
    Xr = np.round(X, 8)
    max_dist = 2 * max(ATOMIC_COV[e] for e in set(up)) + tol
    if lattice is not None:
        pseudo = (lattice @ np.array([0.26, 0.25, 0.25]).reshape(-1, 1)).ravel()
        com = (X * mass[:, None]).sum(0) / mass.sum()
        boundary = (-0.5, 0.5) if np.allclose(com, [0.01, 0, 0], atol=1.0) else (0.0, 1.0)
    else:
        pseudo = (X * mass[:, None]).sum(0) / mass.sum() + np.array([0.01, 0, 0])
    if rebuild:
        S = np.round(supercell_coordinates(X, lattice), 8)
        S[13 * N:14 * N] = Xr                      # image (0,0,0) is value-equal to the cell atom
    else:
        S = Xr
    n_img = S.shape[0] // N
    tree = cKDTree(S)

    def neighbours(node):
        """bonded supercell nodes of ``node``, ascending."""
        i = node % N
        out = []
        for j in sorted(tree.query_ball_point(S[node], max_dist)):
            d = float(np.sqrt(((S[node] - S[j]) ** 2).sum()))
            if not (0.1 < d < max_dist):
                continue
            rs = rcov[i] + rcov[j % N]
            if rs - tol < d < rs + tol:
                out.append(j)
        return out

    central = 13 * N if rebuild else 0
    remaining = np.ones(N, dtype=bool)
    mols = []
    while remaining.any():
        cand = np.nonzero(remaining & heavy)[0]
        if cand.size == 0:
            break
        seed = central + int(cand[int(np.argmin(sk_distances(Xr[cand], pseudo)))])
        frontier, collected, seen = [seed], [], set()
        while frontier:
            nxt = []
            for p in frontier:
                if heavy[p % N]:
                    nb = neighbours(p)
                    cell = [j for j in nb if central <= j < central + N and remaining[j - central]]
                    other = [j for j in nb if not (central <= j < central + N and remaining[j - central])]
                    nxt += cell + (other if rebuild else [])
                collected.append(p)
                seen.add(p)
            for p in frontier:
                if central <= p < central + N:
                    remaining[p - central] = False
            new = []
            for j in nxt:
                if j not in seen and j not in new:
                    new.append(j)
            frontier = new
        idx = np.array(collected)
        el, co = elements[idx % N], S[idx]
        keep = True
        if rebuild:
            m = mass[idx % N]
            c = (co * m[:, None]).sum(0) / m.sum()
            cf = np.around((np.linalg.inv(lattice) @ c.reshape(-1, 1)).ravel(), decimals=8)
            keep = bool(np.all((cf >= boundary[0]) & (cf < boundary[1])))
        if keep:
            mols.append((el, co, idx % N))
    assert n_img in (1, 27)
    return mols
