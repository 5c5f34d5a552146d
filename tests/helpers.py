"""Shared comparison helpers for the parity tests."""
import numpy as np

from oracle import kinematical as K

# tolerances of BASELINE.json north_star
RTOL = 1e-5          # intensities and coordinates
CUT_EPS = 1e-6       # reflections this close to the excitation-error cut may differ
IMG_ATOL = 1e-4      # rendered templates, fraction of peak


def compare_spots(ref, got, s_max, rr, prec=False, noise_floor=1e-10, abs_floor=0.0):
    """ref: oracle dict(g_index, xyz, intensity, excitation_error); got: same keys from the device.

    Reflection sets must be identical except for reflections within CUT_EPS of the cut and for
    round-off "reflections" (I < noise_floor * max I, SURVEY.md section 7 hard part 3)."""
    rI, gI = np.asarray(ref["intensity"]), np.asarray(got["intensity"])
    # ``abs_floor``: when every candidate is a forbidden reflection the pattern consists of |F|^2 ~ 1e-30
    # round-off only (noise against noise, SURVEY.md section 7 hard part 3); such patterns compare as empty
    big = max(rI.max() if rI.size else 0.0, gI.max() if gI.size else 0.0, abs_floor / max(noise_floor, 1e-300))
    rk = {int(k): i for i, k in enumerate(ref["g_index"])}
    gk = {int(k): i for i, k in enumerate(got["g_index"])}
    for k in set(rk) ^ set(gk):
        if k in rk:
            I, s = rI[rk[k]], ref["excitation_error"][rk[k]]
        else:
            I, s = gI[gk[k]], got["excitation_error"][gk[k]]
        near_cut = (not prec) and abs(abs(s) - s_max) < CUT_EPS
        assert near_cut or I < noise_floor * big, f"reflection {k} only on one side (I={I}, s={s})"
    common = sorted(set(rk) & set(gk))
    ri = [rk[k] for k in common]
    gi = [gk[k] for k in common]
    keep = rI[ri] >= noise_floor * big
    np.testing.assert_allclose(gI[gi][keep], rI[ri][keep], rtol=RTOL, atol=0)
    np.testing.assert_allclose(np.asarray(got["xyz"])[gi], np.asarray(ref["xyz"])[ri], rtol=RTOL,
                               atol=1e-6 * rr)
    # order: both are in g-table order
    assert [k for k in ref["g_index"] if k in gk] == [k for k in got["g_index"] if k in rk]
    return len(common)


def row(spots, r):
    """Row r of an engine.SpotTable as host numpy dict."""
    n = int(spots.count[r].item())
    d = dict(g_index=spots.g_index[r, :n].cpu().numpy(), xyz=spots.xyz[r, :n].cpu().numpy(),
             intensity=spots.intensity[r, :n].cpu().numpy())
    d["excitation_error"] = spots.exc[r, :n].cpu().numpy() if spots.exc is not None else np.zeros(n)
    return d


def random_quats(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    return q


def oracle_G_from_active_quat(q):
    """The kernel applies g_lab = R(q) g; the oracle applies g @ G, so G = R(q).T."""
    return K.quat_to_matrix(q).T
