"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the SO(3) grid samplers behind
diffsims/generators/rotation_list_generators.py:85-134 (get_fundamental_zone_grid / get_local_grid).

Those reference functions delegate to orix.sampling.get_sample_fundamental / get_sample_local (orix >= 0.12.1, setup.py:84),
whose source is NOT under /root/reference and which is not installed here: PARITY UNPINNED against orix's point lists.
This file restates the published algorithm orix implements -- cubochoric sampling, Rosca, Morawiec and De Graef (2014)
Modelling Simul. Mater. Sci. Eng. 22 075013, section 3 (cube -> homochoric ball, equal volume), and Singh and De Graef
(2016) ibid. 24 085013 (N = round(131.97049 / (resolution - 0.03732)), cell-centred points) -- and is pinned by properties
of that algorithm (tests/test_so3_grid.py): unit Jacobian, cube surface -> sphere of radius (3 pi / 4)^(1/3), volume fraction
1 / |G| of a fundamental zone, symmetry-reduction invariants.  Only tests/ may import it.
"""
import numpy as np

AP = np.pi ** (2.0 / 3.0)                  # cube edge
R1 = (3.0 * np.pi / 4.0) ** (1.0 / 3.0)    # radius of the homochoric ball
SC = 0.897772786961286
PREK = 1.6434564029725040
PREF = np.sqrt(6.0 / np.pi)


def cu2ho(xyz):
    """Cubochoric -> homochoric, one point [3]."""
    x, y, z = (float(v) for v in xyz)
    if max(abs(x), abs(y), abs(z)) == 0.0:
        return np.zeros(3)
    if abs(x) <= abs(z) and abs(y) <= abs(z):
        pyr, (a, b, c) = 0, (x, y, z)
    elif abs(y) <= abs(x) and abs(z) <= abs(x):
        pyr, (a, b, c) = 1, (y, z, x)
    else:
        pyr, (a, b, c) = 2, (z, x, y)
    a, b, c = a * SC, b * SC, c * SC
    if max(abs(a), abs(b)) == 0.0:
        la, lb, lc = 0.0, 0.0, PREF * c
    else:
        if abs(b) <= abs(a):
            q = (np.pi / 12.0) * b / a
            f = PREK * a / np.sqrt(np.sqrt(2.0) - np.cos(q))
            t1, t2 = (np.sqrt(2.0) * np.cos(q) - 1.0) * f, np.sqrt(2.0) * np.sin(q) * f
        else:
            q = (np.pi / 12.0) * a / b
            f = PREK * b / np.sqrt(np.sqrt(2.0) - np.cos(q))
            t1, t2 = np.sqrt(2.0) * np.sin(q) * f, (np.sqrt(2.0) * np.cos(q) - 1.0) * f
        cc = t1 * t1 + t2 * t2
        s = np.pi * cc / (24.0 * c * c)
        d = np.sqrt(np.pi) * cc / np.sqrt(24.0) / c
        q = np.sqrt(1.0 - s)
        la, lb, lc = t1 * q, t2 * q, PREF * c - d
    return np.array([(la, lb, lc), (lc, la, lb), (lb, lc, la)][pyr])


def ho2qu(h):
    """Homochoric -> unit quaternion (a >= 0): |h|^3 = 3/4 (w - sin w), Newton iteration."""
    hm = float(np.linalg.norm(h))
    if hm == 0.0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    target = (4.0 / 3.0) * hm ** 3
    w = min(np.cbrt(6.0 * target), np.pi)
    for _ in range(50):
        fp = 1.0 - np.cos(w)
        if fp < 1e-300:
            break
        step = (w - np.sin(w) - target) / fp
        w -= step
        if abs(step) < 1e-15 * max(1.0, w):
            break
    w = min(max(w, 0.0), np.pi)
    return np.concatenate([[np.cos(w / 2)], np.sin(w / 2) * np.asarray(h) / hm])


def cubochoric_grid(n_steps):
    """All (2 N)^3 cell-centred grid rotations as quaternions [n, 4], x slowest / z fastest."""
    n = 2 * n_steps
    delta = AP / n
    axis = (np.arange(n) - n_steps + 0.5) * delta
    out = np.empty((n ** 3, 4))
    t = 0
    for x in axis:
        for y in axis:
            for z in axis:
                out[t] = ho2qu(cu2ho((x, y, z)))
                t += 1
    return out


def fundamental_zone_mask(q, sym, tol=1e-9):
    """q [n, 4] (a >= 0) is in the fundamental zone of the proper group ``sym`` [m, 4] iff no s q has a larger |scalar|."""
    w = sym[:, :1] * q[:, 0] - sym[:, 1:2] * q[:, 1] - sym[:, 2:3] * q[:, 2] - sym[:, 3:4] * q[:, 3]   # [m, n]
    return np.all(np.abs(w) <= q[:, 0] + tol, axis=0)


def rotation_angle(q):
    return 2.0 * np.arccos(np.minimum(np.abs(q[:, 0]), 1.0))
