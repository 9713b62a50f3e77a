"""Synthetic trajectories of the shapes BASELINE.json names, written in the file
formats the trajectory front-ends read, so the same file can feed the reference and this package.

frame f = seed[f mod len(seed)] + N(0, sigma) per coordinate, numpy.random.default_rng(seed0 + f).
"""
from __future__ import annotations

import numpy as np

SEED0 = 20261017


def noisy_frames(seeds, first, count, sigma=0.05, seed0=SEED0, decimals=None):
    seeds = np.asarray(seeds, dtype=np.float64)
    out = np.empty((count,) + seeds.shape[1:], dtype=np.float64)
    for i in range(count):
        f = first + i
        out[i] = seeds[f % len(seeds)] + np.random.default_rng(seed0 + f).normal(0.0, sigma, seeds.shape[1:])
    return out if decimals is None else np.round(out, decimals)


def write_history(path, atom_keys, frames, lattice=None, masses=None, title="pywindow_b200 synthetic HISTORY"):
    """DL_POLY HISTORY, keytrj 0 (coordinates only), imcon 0 or 3 (reference reader: trajectory.py:647-766)."""
    frames = np.asarray(frames, dtype=np.float64)
    n = frames.shape[1]
    imcon = 0 if lattice is None else 3
    masses = np.ones(n) if masses is None else masses
    with open(path, "w") as fh:
        fh.write(f"{title}\n{0:10d}{imcon:10d}{n:10d}\n")
        for k, X in enumerate(frames):
            fh.write(f"timestep{k + 1:10d}{n:10d}{0:10d}{imcon:10d}{0.0007:12.6f}\n")
            if lattice is not None:
                for row in np.asarray(lattice, dtype=np.float64).T:     # reader transposes (trajectory.py:724)
                    fh.write("".join(f"{v:20.10f}" for v in row) + "\n")
            for i in range(n):
                fh.write(f"{atom_keys[i]:<8}{i + 1:10d}{masses[i]:12.6f}{0.0:12.6f}\n")
                fh.write("".join(f"{v:20.10E}" for v in X[i]) + "\n")


def write_history_native(path, atom_keys, frames, lattice=None, masses=None, title="pywindow_b200 synthetic HISTORY",
                         threads=0, compact=False):
    """Same file as :func:`write_history`, byte for byte, written by host threads (``pwb_history_write``):
    a 100 000-frame trajectory (1.8 GB) takes about a second instead of minutes.  ``compact``: coordinates in
    DL_POLY's own ``%12.4E`` records (the format of the reference's example HISTORY files) instead of ``%20.10E``."""
    from . import _lib
    frames = np.ascontiguousarray(frames, dtype=np.float64)
    F, n = frames.shape[0], frames.shape[1]
    keys = np.array([f"{str(k):<8}"[:8] for k in atom_keys], dtype="S8")
    m = np.ones(n) if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    lat = None if lattice is None else np.ascontiguousarray(lattice, dtype=np.float64)
    size = _lib.load().pwb_history_write(str(path).encode(), title.encode(), keys.ctypes.data, m.ctypes.data,
                                          frames.ctypes.data, F, n, lat.ctypes.data if lat is not None else None,
                                          int(lat is not None and lat.ndim == 3), int(bool(compact)), threads)
    if size < 0:
        raise OSError(f"could not write {path}")
    return int(size)


def write_xyz_trajectory(path, elements, frames):
    with open(path, "w") as fh:
        for k, X in enumerate(frames):
            fh.write(f"{len(elements)}\nframe {k}\n")
            for e, c in zip(elements, X):
                fh.write(f"{e} {c[0]:.8f} {c[1]:.8f} {c[2]:.8f}\n")


def write_pdb_trajectory(path, elements, frames, unit_cell=None):
    with open(path, "w") as fh:
        for X in frames:
            if unit_cell is not None:
                c = unit_cell
                fh.write(f"CRYST1{c[0]:9.3f}{c[1]:9.3f}{c[2]:9.3f}{c[3]:7.2f}{c[4]:7.2f}{c[5]:7.2f}\n")
            for i, (e, c) in enumerate(zip(elements, X), start=1):
                fh.write(f"HETATM{i % 100000:>5} {e + str(i % 1000):<4} MOL A   1    "
                         f"{c[0]:>8.3f}{c[1]:>8.3f}{c[2]:>8.3f}{1.0:>6.2f}{0.0:>6.2f}          {e:>2}\n")
            fh.write("END\n")
