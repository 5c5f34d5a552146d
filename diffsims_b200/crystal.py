"""Light crystallography stand-ins for the third-party objects the reference's
hot path receives (``diffpy.structure.{Lattice,Atom,Structure}``,
``orix.crystal_map.Phase``, ``orix.quaternion.Rotation``).

Neither diffpy.structure nor orix is installed in this image, so
the drop-in layer is duck-typed: every consumer in this package only touches the
attribute names listed in SURVEY.md §8(b) "Input object protocol", which both the
real third-party objects and these stand-ins provide.  The conventions restated
here are the ones the reference relies on:

* diffpy ``Lattice``: row vectors, ``base = stdbase @ baserot`` with the standard
  setting a* || x, c || z; ``recbase = inv(base)``; ``rnorm(hkl) = |hkl @ recbase.T|``
  (used at diffsims/crystallography/reciprocal_lattice_vector.py:440,
  diffsims/utils/sim_utils.py:290, :457-472).
* orix ``Phase``: the structure's lattice is re-aligned to a || x, c* || z while the
  atoms keep their *Cartesian* positions (so their fractional coordinates change);
  the reference undoes exactly this at diffsims/utils/sim_utils.py:290-291.
* orix ``Rotation.from_euler``: Bunge ZXZ, lab->crystal; ``to_matrix()`` is the
  passive Bunge matrix G and ``~rot * v = G.T @ v``
  (diffsims/crystallography/_diffracting_vector.py:157-160).

These classes are host-side bookkeeping only; no arithmetic of the hot path
(rotation of g, excitation error, structure factors, rasterisation) lives here.
"""
from __future__ import annotations

import copy
import math
import re

import numpy as np

__all__ = ["Lattice", "Atom", "Structure", "Phase", "Rotation", "get_element"]


def get_element(atom_type_symbol: str) -> str:
    """Alphabetic head of a CIF ``_atom_type_symbol`` ("Fe3+" -> "Fe").

    Mirrors diffsims/structure_factor/atomic_scattering_parameters.py:145-164.
    """
    return re.match(r"^([A-Za-z]*)", atom_type_symbol).group(1)


def _cosd(x):
    # exact values at the multiples of 30/90 degrees diffpy special-cases
    x = float(x) % 360.0
    table = {0.0: 1.0, 60.0: 0.5, 90.0: 0.0, 120.0: -0.5, 180.0: -1.0,
             240.0: -0.5, 270.0: 0.0, 300.0: 0.5}
    return table.get(x, math.cos(math.radians(x)))


def _sind(x):
    # diffpy defines sind(x) = cosd(90 - x)
    return _cosd(90.0 - float(x))


class Lattice:
    """Subset of ``diffpy.structure.Lattice`` (row-vector convention)."""

    def __init__(self, a=None, b=None, c=None, alpha=None, beta=None, gamma=None,
                 baserot=None, base=None):
        self.baserot = np.identity(3)
        if base is not None:
            self.setLatBase(base)
        elif a is None:
            self.setLatPar(1.0, 1.0, 1.0, 90.0, 90.0, 90.0, baserot)
        else:
            self.setLatPar(a, b, c, alpha, beta, gamma, baserot)

    # -- construction -----------------------------------------------------
    def setLatPar(self, a=None, b=None, c=None, alpha=None, beta=None, gamma=None,
                  baserot=None):
        for name, val in (("a", a), ("b", b), ("c", c), ("alpha", alpha),
                          ("beta", beta), ("gamma", gamma)):
            if val is not None:
                setattr(self, "_" + name, float(val))
        if baserot is not None:
            self.baserot = np.array(baserot, dtype=float)
        a, b, c = self._a, self._b, self._c
        ca, sa = _cosd(self._alpha), _sind(self._alpha)
        cb, sb = _cosd(self._beta), _sind(self._beta)
        cg, sg = _cosd(self._gamma), _sind(self._gamma)
        vunit = math.sqrt(1.0 + 2.0 * ca * cb * cg - ca * ca - cb * cb - cg * cg)
        self.ar = sa / (a * vunit)
        self.br = sb / (b * vunit)
        self.cr = sg / (c * vunit)
        car = (cb * cg - ca) / (sb * sg)
        cbr = (ca * cg - cb) / (sa * sg)
        cgr = (ca * cb - cg) / (sa * sb)
        sgr = math.sqrt(1.0 - cgr * cgr)
        self.alphar = math.degrees(math.acos(car))
        self.betar = math.degrees(math.acos(cbr))
        self.gammar = math.degrees(math.acos(cgr))
        # standard setting: a* || x, c || z, b in the y-z plane
        self.stdbase = np.array(
            [[1.0 / self.ar, -cgr / sgr / self.ar, cb * a],
             [0.0, b * sa, b * ca],
             [0.0, 0.0, c]], dtype=float)
        self.base = self.stdbase @ self.baserot
        self.recbase = np.linalg.inv(self.base)
        self.volume = abs(np.linalg.det(self.base))
        return self

    def setLatBase(self, base):
        base = np.array(base, dtype=float)
        va, vb, vc = base
        a, b, c = (float(np.linalg.norm(v)) for v in (va, vb, vc))
        ca = float(vb @ vc) / (b * c)
        cb = float(va @ vc) / (a * c)
        cg = float(va @ vb) / (a * b)
        self._a, self._b, self._c = a, b, c
        self._alpha = math.degrees(math.acos(ca))
        self._beta = math.degrees(math.acos(cb))
        self._gamma = math.degrees(math.acos(cg))
        self.baserot = np.identity(3)
        self.setLatPar()
        self.baserot = np.linalg.inv(self.stdbase) @ base
        self.base = base
        self.recbase = np.linalg.inv(base)
        self.volume = abs(np.linalg.det(base))
        return self

    a = property(lambda s: s._a, lambda s, v: s.setLatPar(a=v))
    b = property(lambda s: s._b, lambda s, v: s.setLatPar(b=v))
    c = property(lambda s: s._c, lambda s, v: s.setLatPar(c=v))
    alpha = property(lambda s: s._alpha, lambda s, v: s.setLatPar(alpha=v))
    beta = property(lambda s: s._beta, lambda s, v: s.setLatPar(beta=v))
    gamma = property(lambda s: s._gamma, lambda s, v: s.setLatPar(gamma=v))

    def abcABG(self):
        return (self._a, self._b, self._c, self._alpha, self._beta, self._gamma)

    # -- geometry ---------------------------------------------------------
    def reciprocal(self):
        return Lattice(base=self.recbase.T)

    def cartesian(self, u):
        return np.asarray(u, dtype=float) @ self.base

    def fractional(self, rc):
        return np.asarray(rc, dtype=float) @ self.recbase

    def norm(self, xyz):
        return np.sqrt((self.cartesian(xyz) ** 2).sum(axis=-1))

    def rnorm(self, hkl):
        g = np.asarray(hkl, dtype=float) @ self.recbase.T
        return np.sqrt((g ** 2).sum(axis=-1))

    def dist(self, u, v):
        d = np.asarray(u, dtype=float) - np.asarray(v, dtype=float)
        return self.norm(d)

    def __repr__(self):
        return ("Lattice(a=%g, b=%g, c=%g, alpha=%g, beta=%g, gamma=%g)" % self.abcABG())


class Atom:
    """Subset of ``diffpy.structure.Atom``: element label, fractional xyz, occupancy."""

    def __init__(self, atype="", xyz=(0.0, 0.0, 0.0), occupancy=1.0, lattice=None,
                 Uisoequiv=0.0):
        self.element = atype
        self.xyz = np.array(xyz, dtype=float)
        self.occupancy = float(occupancy)
        self.lattice = lattice
        self.Uisoequiv = Uisoequiv

    def __repr__(self):
        return "%-4s %8.6f %8.6f %8.6f %6.4f" % (self.element, *self.xyz, self.occupancy)


class Structure(list):
    """Subset of ``diffpy.structure.Structure``: a list of atoms plus a lattice."""

    def __init__(self, atoms=None, lattice=None, title=""):
        super().__init__()
        self.lattice = lattice if lattice is not None else Lattice()
        self.title = title
        for a in atoms or []:
            b = copy.copy(a)
            b.xyz = np.array(a.xyz, dtype=float)
            b.lattice = self.lattice
            self.append(b)

    @property
    def xyz(self):
        return np.array([a.xyz for a in self], dtype=float).reshape(-1, 3)

    @property
    def xyz_cartn(self):
        return self.xyz @ self.lattice.base

    @property
    def element(self):
        return np.array([a.element for a in self])

    @property
    def occupancy(self):
        return np.array([a.occupancy for a in self], dtype=float)

    def copy(self):
        return Structure(atoms=list(self), lattice=copy.deepcopy(self.lattice),
                         title=self.title)

    def __deepcopy__(self, memo):
        return self.copy()


def _aligned_base(old_base):
    """Lattice base re-aligned to a || x, c* || z (what orix ``Phase`` enforces)."""
    lat = Lattice(base=old_base)
    a, b, c, al, be, ga = lat.abcABG()
    ca, cb, cg, sg = _cosd(al), _cosd(be), _cosd(ga), _sind(ga)
    vunit = math.sqrt(1.0 + 2.0 * ca * cb * cg - ca * ca - cb * cb - cg * cg)
    return np.array([[a, 0.0, 0.0],
                     [b * cg, b * sg, 0.0],
                     [c * cb, c * (ca - cb * cg) / sg, c * vunit / sg]], dtype=float)


class _PointGroup:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return str(self.name)


# Laue/point-group name from the international space-group number (coarse: crystal
# system's holohedry; only used for repr and ``get_beam_directions_grid`` defaults).
def _holohedry_from_space_group(n):
    n = int(n)
    if n <= 2:
        return "-1"
    if n <= 15:
        return "2/m"
    if n <= 74:
        return "mmm"
    if n <= 142:
        return "4/mmm"
    if n <= 167:
        return "-3m"
    if n <= 194:
        return "6/mmm"
    return "m-3m"


class Phase:
    """Subset of ``orix.crystal_map.Phase`` (name, structure, point group)."""

    def __init__(self, name=None, space_group=None, point_group=None, structure=None,
                 color=None):
        self._structure = None
        self.name = name if name is not None else ""
        self.space_group = space_group
        if point_group is None and space_group is not None:
            point_group = _holohedry_from_space_group(space_group)
        self.point_group = _PointGroup(point_group) if point_group is not None else None
        self.color = color
        if structure is not None:
            self.structure = structure
            if not self.name:
                self.name = getattr(structure, "title", "") or ""
        else:
            self._structure = Structure()

    @property
    def structure(self):
        return self._structure

    @structure.setter
    def structure(self, value):
        old_base = np.array(value.lattice.base, dtype=float)
        new_base = _aligned_base(old_base)
        cart = np.array([np.asarray(a.xyz, dtype=float) for a in value]).reshape(-1, 3) @ old_base
        new_lat = Lattice(base=new_base)
        new_frac = cart @ new_lat.recbase
        atoms = []
        for a, f in zip(value, new_frac):
            atoms.append(Atom(a.element, f, getattr(a, "occupancy", 1.0)))
        self._structure = Structure(atoms=atoms, lattice=new_lat,
                                    title=getattr(value, "title", ""))

    def deepcopy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return f"<name: {self.name}. point group: {self.point_group}>"


class Rotation:
    """Subset of ``orix.quaternion.Rotation``: unit quaternions ``(a, b, c, d)``.

    ``to_matrix()`` follows orix: for ``from_euler`` (Bunge ZXZ, lab->crystal) it
    is the passive matrix G, and ``rot * v`` is ``G @ v``.
    """

    def __init__(self, data):
        data = np.atleast_2d(np.asarray(getattr(data, "data", data), dtype=float))
        if data.shape[-1] != 4:
            raise ValueError("Rotation data must have a last dimension of 4")
        self.data = data.reshape(-1, 4)

    # -- constructors -----------------------------------------------------
    @classmethod
    def identity(cls, shape=(1,)):
        n = int(np.prod(shape))
        q = np.zeros((n, 4))
        q[:, 0] = 1.0
        return cls(q)

    @classmethod
    def from_euler(cls, euler, direction="lab2crystal", degrees=False):
        e = np.atleast_2d(np.asarray(euler, dtype=float)).reshape(-1, 3)
        if degrees:
            e = np.deg2rad(e)
        phi1, Phi, phi2 = e[:, 0], e[:, 1], e[:, 2]
        sigma = 0.5 * (phi1 + phi2)
        delta = 0.5 * (phi1 - phi2)
        c, s = np.cos(Phi / 2), np.sin(Phi / 2)
        q = np.stack([c * np.cos(sigma), -s * np.cos(delta), -s * np.sin(delta),
                      -c * np.sin(sigma)], axis=1)
        q[q[:, 0] < 0] *= -1
        rot = cls(q)
        if direction == "crystal2lab":
            rot = ~rot
        return rot

    @classmethod
    def from_matrix(cls, matrix):
        om = np.asarray(matrix, dtype=float).reshape(-1, 3, 3)
        # passive convention consistent with to_matrix(): om = R(q) below
        t = om[:, 0, 0] + om[:, 1, 1] + om[:, 2, 2]
        a = 0.5 * np.sqrt(np.maximum(1 + t, 0))
        b = 0.5 * np.sqrt(np.maximum(1 + om[:, 0, 0] - om[:, 1, 1] - om[:, 2, 2], 0))
        c = 0.5 * np.sqrt(np.maximum(1 - om[:, 0, 0] + om[:, 1, 1] - om[:, 2, 2], 0))
        d = 0.5 * np.sqrt(np.maximum(1 - om[:, 0, 0] - om[:, 1, 1] + om[:, 2, 2], 0))
        b = np.where(om[:, 2, 1] < om[:, 1, 2], -b, b)
        c = np.where(om[:, 0, 2] < om[:, 2, 0], -c, c)
        d = np.where(om[:, 1, 0] < om[:, 0, 1], -d, d)
        q = np.stack([a, b, c, d], axis=1)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        return cls(q)

    @classmethod
    def random(cls, shape=(1,), rng=None):
        n = int(np.prod(shape))
        rng = np.random.default_rng(rng)
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        q[q[:, 0] < 0] *= -1
        return cls(q)

    # -- conversions ------------------------------------------------------
    def to_matrix(self):
        a, b, c, d = self.data.T
        om = np.empty((self.size, 3, 3))
        om[:, 0, 0] = a * a + b * b - c * c - d * d
        om[:, 1, 1] = a * a - b * b + c * c - d * d
        om[:, 2, 2] = a * a - b * b - c * c + d * d
        om[:, 0, 1] = 2 * (b * c - a * d)
        om[:, 1, 0] = 2 * (b * c + a * d)
        om[:, 0, 2] = 2 * (b * d + a * c)
        om[:, 2, 0] = 2 * (b * d - a * c)
        om[:, 1, 2] = 2 * (c * d - a * b)
        om[:, 2, 1] = 2 * (c * d + a * b)
        return om

    def to_euler(self, degrees=False):
        om = self.to_matrix()
        Phi = np.arccos(np.clip(om[:, 2, 2], -1, 1))
        sing = np.isclose(np.abs(om[:, 2, 2]), 1.0)
        phi1 = np.where(sing, np.arctan2(om[:, 0, 1], om[:, 0, 0]),
                        np.arctan2(om[:, 2, 0], -om[:, 2, 1]))
        phi2 = np.where(sing, 0.0, np.arctan2(om[:, 0, 2], om[:, 1, 2]))
        e = np.stack([phi1 % (2 * np.pi), Phi, phi2 % (2 * np.pi)], axis=1)
        return np.rad2deg(e) if degrees else e

    # -- container protocol -------------------------------------------------
    @property
    def size(self):
        return self.data.shape[0]

    @property
    def shape(self):
        return (self.data.shape[0],)

    def __len__(self):
        return self.size

    def __getitem__(self, key):
        return Rotation(np.atleast_2d(self.data[key]))

    def __iter__(self):
        for i in range(self.size):
            yield Rotation(self.data[i:i + 1])

    def __invert__(self):
        q = self.data.copy()
        q[:, 1:] *= -1
        return Rotation(q)

    def __mul__(self, other):
        if isinstance(other, Rotation):
            a1, b1, c1, d1 = self.data.T
            a2, b2, c2, d2 = other.data.T
            return Rotation(np.stack([
                a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
                a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
                a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2], axis=1))
        v = np.asarray(getattr(other, "data", other), dtype=float)
        if self.size != 1:
            raise ValueError("rotation * vectors needs a single rotation")
        return v @ self.to_matrix()[0].T

    def __repr__(self):
        return f"Rotation {self.shape}\n{self.data}"
