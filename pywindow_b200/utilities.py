"""Geometry functions of the pore-analysis path, with the reference's names and call conventions
(``pywindow/_internal/utilities.py``; imported by its ``molecular.py:29-44``), executed on the GPU.

Every function takes ``(elements: array[str], coordinates: array[N,3] float64, ...)`` and returns the
same python ``float`` / ``int`` / ``numpy`` types as the reference.  They are thin single-molecule
wrappers over :class:`pywindow_b200.engine.Engine`; batched work (trajectories) should call the engine
directly (see ``trajectory.py``).  Host-only helpers (atom-key deciphering, unit-cell algebra, the
tolerance comparator) are plain numpy -- they run once per trajectory, not per frame.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .engine import default_engine, record_windows, sphere_volume  # noqa: F401  (re-exported)
from .tables import atomic_covalent_radius, atomic_mass, atomic_vdw_radius, opls_atom_keys  # noqa: F401


class _AtomKeyError(Exception):
    def __init__(self, message):
        self.message = message


class _AtomKeyConflictError(Exception):
    def __init__(self, message):
        self.message = message


class _ForceFieldError(Exception):
    def __init__(self, message):
        self.message = message


class _FunctionError(Exception):
    def __init__(self, message):
        self.message = message


# --------------------------------------------------------------------------- small host helpers
def is_number(number):
    try:
        float(number)
    except ValueError:
        return False
    return True


def unique(input_list):
    out = []
    for item in input_list:
        if item not in out:
            out.append(item)
    return out


def to_list(obj):
    """JSON serialiser hook for numpy arrays (utilities.py:72-77)."""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    raise TypeError("Not serializable")


def distance(a, b):
    return float((np.sum((np.asarray(a) - np.asarray(b)) ** 2)) ** 0.5)


def molecular_weight(elements):
    """utilities.py:96-107."""
    return float(np.array([atomic_mass[i.upper()] for i in elements]).sum())


def center_of_coor(coordinates):
    """utilities.py:110-124."""
    return np.sum(coordinates, axis=0) / coordinates.shape[0]


# --------------------------------------------------------------------------- GPU-backed functions
def _run(elements, coordinates, stages, **kw):
    X = np.ascontiguousarray(coordinates, dtype=np.float64).reshape(1, -1, 3)
    res = default_engine().analyse_frames(X, elements, stages=stages, **kw)
    return res.host_records()[0]


def center_of_mass(elements, coordinates):
    """utilities.py:127-148."""
    rec = _run(elements, coordinates, _lib.STAGE_SCALARS)
    return np.array(rec[_lib.REC_COM:_lib.REC_COM + 3])


def shift_com(elements, coordinates, com_adjust=np.zeros(3)):  # noqa: B008
    """utilities.py:344-352."""
    com = center_of_mass(elements, coordinates)
    return np.asarray(coordinates, dtype=np.float64) - (com - com_adjust)


def max_dim(elements, coordinates):
    """utilities.py:355-372 -> (atom_1, atom_2, diameter)."""
    rec = _run(elements, coordinates, _lib.STAGE_SCALARS)
    return int(rec[_lib.REC_MAXD_A1]), int(rec[_lib.REC_MAXD_A2]), float(rec[_lib.REC_MAXD])


def pore_diameter(elements, coordinates, com=None):
    """utilities.py:375-388 -> (diameter, closest atom).  ``com`` overrides the centre."""
    if com is None:
        rec = _run(elements, coordinates, _lib.STAGE_SCALARS)
        return float(rec[_lib.REC_PORE_D]), int(rec[_lib.REC_PORE_ATOM])
    rec = _run(elements, coordinates, _lib.STAGE_SCALARS | _lib.STAGE_OPT,
               centre_override=np.asarray(com, dtype=np.float64).reshape(1, 3))
    return float(rec[_lib.REC_OPT_D]), int(rec[_lib.REC_OPT_ATOM])


def opt_pore_diameter(elements, coordinates, bounds=None, com=None):
    """utilities.py:400-426 -> (diameter, closest atom, centre).  ``com`` is the optimiser's start point
    (default: the centre of mass); ``bounds`` = ((lo, hi),) * 3 with ``None`` for a missing side (default: the box
    start +- pore radius at the start point).  Inverted bounds raise scipy's ``ValueError`` like the reference."""
    kw = {}
    if com is not None:
        kw["opt_start"] = np.asarray(com, dtype=np.float64).reshape(1, 3)
    if bounds is not None:
        bb = np.array([[-np.inf if lo is None else float(lo), np.inf if hi is None else float(hi)]
                       for lo, hi in bounds], dtype=np.float64)
        if bb.shape != (3, 2):
            raise ValueError("bounds must be three (lower, upper) pairs")
        kw["opt_bounds"] = bb.reshape(1, 6)
    rec = _run(elements, coordinates, _lib.STAGE_SCALARS | _lib.STAGE_OPT, **kw)
    if int(rec[_lib.REC_STATUS]) & _lib.ST_OPT_BOUNDS:
        raise ValueError("An upper bound is less than the corresponding lower bound.")
    return (float(rec[_lib.REC_OPT_D]), int(rec[_lib.REC_OPT_ATOM]),
            np.array(rec[_lib.REC_OPT_C:_lib.REC_OPT_C + 3]))


def find_average_diameter(elements, coordinates, adjust=1, processes=None):  # noqa: ARG001
    """utilities.py:1586-1650."""
    rec = _run(elements, coordinates, _lib.STAGE_SCALARS | _lib.STAGE_AVGD, adjust=float(adjust))
    return float(rec[_lib.REC_AVGD])


def find_windows(elements, coordinates, processes=None, adjust=1, pore_opt=True, increment=1.0,  # noqa: ARG001
                 pore_centre=None, shard_rays=None):
    """utilities.py:1364-1553 -> None | (diameters[W], centres[W,3]).

    ``processes`` is accepted and ignored (the GPU analyses all rays / windows at once).
    ``pore_centre`` (extension) forces the ray origin instead of optimising it.
    ``shard_rays``: spread the rays of this ONE molecule over the ranks of the initialised ``torch.distributed``
    group (a collective call; default: when a group of more than one rank exists and ``adjust >= 32``, i.e. dense
    sampling with >= ~3e4 rays, BASELINE config 5)."""
    stages = _lib.STAGE_ALL & ~_lib.STAGE_AVGD
    kw = {"adjust": float(adjust), "increment": float(increment), "pore_opt": bool(pore_opt)}
    if pore_centre is not None:
        kw["centre_override"] = np.asarray(pore_centre, dtype=np.float64).reshape(1, 3)
    elif not pore_opt:
        # the reference never calls opt_pore_diameter here (utilities.py:1386-1393): a molecule whose COM lies
        # inside an atom's vdW sphere must still return its windows instead of scipy's inverted-bounds error
        stages &= ~_lib.STAGE_OPT
    from .distributed import world
    if shard_rays is None:
        shard_rays = world()[1] > 1 and float(adjust) >= 32
    if shard_rays and world()[1] > 1:
        res = default_engine().analyse_ray_sharded(elements, coordinates, stages=stages, **kw)
        rec = res.host_records()[0]
    else:
        rec = _run(elements, coordinates, stages, **kw)
    if int(rec[_lib.REC_STATUS]) & _lib.ST_OPT_BOUNDS:
        raise ValueError("An upper bound is less than the corresponding lower bound.")
    return record_windows(rec)


# --------------------------------------------------------------------------- force-field keys (host)
def dlf_notation(atom_key):
    """Element of a DL_F atom key: leading letters up to the first digit (utilities.py:267-285)."""
    chars = list(atom_key)
    element, count = "", 0
    while True:
        element += chars[count]
        count += 1
        if is_number(chars[count]):
            break
    element = "".join(c for c in element if not is_number(c))
    return "".join(c for c in element if c != "?")


def opls_notation(atom_key):
    """utilities.py:288-305."""
    if atom_key in ("ne", "he", "na"):
        raise _AtomKeyConflictError(
            f"One of the OPLS conflicting atom_keys has occured '{atom_key}'. For how to solve this issue see "
            "the manual or MolecularSystem._atom_key_swap() doc string.")
    for element, keys in opls_atom_keys.items():
        if atom_key in keys:
            return element
    raise _AtomKeyError(f"OPLS atom key {atom_key} was not found in OPLS keys dictionary.")


def decipher_atom_key(atom_key, forcefield):
    """utilities.py:308-341."""
    funcs = {"DLF": dlf_notation, "DL_F": dlf_notation, "OPLS": opls_notation, "OPLSAA": opls_notation,
             "OPLS2005": opls_notation, "OPLS3": opls_notation}
    if forcefield.upper() in funcs:
        return funcs[forcefield.upper()](atom_key)
    raise _ForceFieldError(f"Unfortunetely, '{forcefield}' forcefield is not supported by pyWINDOW.")


# --------------------------------------------------------------------------- unit-cell algebra (host)
def unit_cell_to_lattice_array(cryst):
    """utilities.py:653-690."""
    a_, b_, c_, alpha, beta, gamma = cryst
    ra, rb, rg = np.deg2rad(alpha), np.deg2rad(beta), np.deg2rad(gamma)
    volume = a_ * b_ * c_ * (1 - np.cos(ra) ** 2 - np.cos(rb) ** 2 - np.cos(rg) ** 2
                             + 2 * np.cos(ra) * np.cos(rb) * np.cos(rg)) ** 0.5
    return np.array([[a_, b_ * np.cos(rg), c_ * np.cos(rb)],
                     [0, b_ * np.sin(rg), c_ * (np.cos(ra) - np.cos(rb) * np.cos(rg)) / np.sin(rg)],
                     [0, 0, volume / (a_ * b_ * np.sin(rg))]])


def lattice_array_to_unit_cell(lattice_array):
    """utilities.py:693-710."""
    lengths = np.sqrt(np.sum(lattice_array ** 2, axis=0))
    gamma_r = np.arccos(lattice_array[0][1] / lengths[1])
    beta_r = np.arccos(lattice_array[0][2] / lengths[2])
    alpha_r = np.arccos(lattice_array[1][2] * np.sin(gamma_r) / lengths[2] + np.cos(beta_r) * np.cos(gamma_r))
    return np.append(lengths, [np.rad2deg(alpha_r), np.rad2deg(beta_r), np.rad2deg(gamma_r)])


def volume_from_lattice_array(lattice_array):
    return np.linalg.det(lattice_array)


def volume_from_cell_parameters(cryst):
    return volume_from_lattice_array(unit_cell_to_lattice_array(cryst))


def fractional_from_cartesian(coordinate, lattice_array):
    """utilities.py:723-730."""
    return (np.linalg.inv(lattice_array) @ np.asarray(coordinate).reshape(-1, 1)).reshape(1, -1)


def cartisian_from_fractional(coordinate, lattice_array):
    """utilities.py:733-740."""
    return (np.asarray(lattice_array) @ np.asarray(coordinate).reshape(-1, 1)).reshape(1, -1)


def cart2frac_all(coordinates, lattice_array):
    """utilities.py:743-752 (vectorised)."""
    return np.asarray(coordinates) @ np.linalg.inv(lattice_array).T


def frac2cart_all(frac_coordinates, lattice_array):
    """utilities.py:755-765 (vectorised)."""
    return np.asarray(frac_coordinates) @ np.asarray(lattice_array).T


# --------------------------------------------------------------------------- comparator (host)
_COMPARED = (("centre_of_mass", "array"), ("maximum_diameter.atom_1", "int"), ("maximum_diameter.atom_2", "int"),
             ("maximum_diameter.diameter", "float"), ("no_of_atoms", "int"), ("pore_diameter.atom", "int"),
             ("pore_diameter.diameter", "float"), ("pore_diameter_opt.atom_1", "int"),
             ("pore_diameter_opt.centre_of_mass", "array"), ("pore_diameter_opt.diameter", "float"),
             ("pore_volume", "float"), ("pore_volume_opt", "float"), ("windows.centre_of_mass", "array"),
             ("windows.diameters", "array"), ("average_diameter", "float"))


def compare_properties_dict(dict1, dict2):
    """Compare two ``Molecule.properties`` dicts the way the reference does (utilities.py:1694-1754):
    returns ``(True, "none")`` or ``(False, <first differing property>)``; arrays with ``np.allclose``,
    floats with ``np.isclose`` (numpy default tolerances), ints and ``None`` exactly; a property
    missing from both is skipped."""
    for prop, kind in _COMPARED:
        head, _, tail = prop.partition(".")
        if head not in dict1 and head not in dict2:
            continue
        if (head in dict1) != (head in dict2):
            return (False, prop)
        a, b = dict1[head], dict2[head]
        if tail:
            a, b = a[tail], b[tail]
        if (a is None) != (b is None):
            return (False, prop)
        if a is None:
            continue
        if kind == "array" and not np.allclose(a, b):
            return (False, prop)
        if kind == "float" and not np.isclose(a, b):
            return (False, prop)
        if kind == "int" and a != b:
            return (False, prop)
    return (True, "none")
