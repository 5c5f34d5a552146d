"""SO(3) rotation-list producers (get_fundamental_zone_grid / get_local_grid, reference rotation_list_generators.py:85-134).

orix's sampler is third-party and absent: parity with its point lists is UNPINNED.  Pinned here: the properties of the
published cubochoric construction the oracle (and the kernel) restate, and the reference's own test of these functions."""
import numpy as np
import pytest

from oracle import so3


def test_cubochoric_map_is_volume_preserving():
    """|det J| = 1 everywhere (cube volume pi^2 = ball volume): a wrong constant or branch in the map breaks this."""
    rng = np.random.default_rng(0)
    half = so3.AP / 2
    for p in rng.uniform(-0.97 * half, 0.97 * half, size=(200, 3)):
        e = 1e-6
        J = np.stack([(so3.cu2ho(p + e * np.eye(3)[k]) - so3.cu2ho(p - e * np.eye(3)[k])) / (2 * e) for k in range(3)], axis=1)
        # (finite differences across a pyramid boundary are not meaningful: skip points within e of one)
        a = np.sort(np.abs(p))
        if a[2] - a[1] < 10 * e or a[1] - a[0] < 10 * e:
            continue
        assert abs(abs(np.linalg.det(J)) - 1.0) < 1e-5


def test_cube_surface_maps_to_the_sphere_and_centre_to_identity():
    rng = np.random.default_rng(1)
    half = so3.AP / 2
    for _ in range(100):
        p = rng.uniform(-half, half, 3)
        p[rng.integers(3)] = half * rng.choice([-1, 1])
        assert abs(np.linalg.norm(so3.cu2ho(p)) - so3.R1) < 1e-12
    np.testing.assert_allclose(so3.ho2qu(so3.cu2ho((0, 0, 0))), [1, 0, 0, 0])
    # the ball surface is the rotation angle pi
    q = so3.ho2qu(so3.cu2ho((half, 0.1, -0.2)))
    assert abs(q[0]) < 1e-12 and abs(np.linalg.norm(q) - 1) < 1e-12


def test_fundamental_zone_volume_fractions_and_symmetry_reduction():
    from diffsims_b200.generators.rotation_list_generators import _PROPER_GENERATORS, proper_point_group_quaternions
    q = so3.cubochoric_grid(12)                 # 13 824 rotations, equal-volume cells
    assert np.all(q[:, 0] >= 0) and np.allclose(np.linalg.norm(q, axis=1), 1)
    for name in _PROPER_GENERATORS:
        sym = proper_point_group_quaternions(name)
        inside = so3.fundamental_zone_mask(q, sym)
        frac = inside.mean()
        assert abs(frac * len(sym) - 1.0) < 0.15, (name, frac)      # volume of SO(3) / |G| (boundary cells included)
        # every rotation has a symmetric equivalent inside the zone, and the one inside has the smallest angle
        for r in q[::997]:
            eq = np.array([[s[0] * r[0] - s[1] * r[1] - s[2] * r[2] - s[3] * r[3],
                            s[0] * r[1] + s[1] * r[0] + s[2] * r[3] - s[3] * r[2],
                            s[0] * r[2] - s[1] * r[3] + s[2] * r[0] + s[3] * r[1],
                            s[0] * r[3] + s[1] * r[2] - s[2] * r[1] + s[3] * r[0]] for s in sym])
            eq[eq[:, 0] < 0] *= -1
            m = so3.fundamental_zone_mask(eq, sym)
            assert m.any()
            assert np.isclose(so3.rotation_angle(eq[m]).min(), so3.rotation_angle(eq).min())


def test_group_tables():
    from diffsims_b200.generators.rotation_list_generators import (_proper_group_of_space_group, _resolve_proper_group,
                                                                   proper_point_group_quaternions, resolution_to_semi_edge_steps)
    orders = {"1": 1, "2": 2, "222": 4, "4": 4, "422": 8, "3": 3, "32": 6, "6": 6, "622": 12, "23": 12, "432": 24}
    for name, n in orders.items():
        g = proper_point_group_quaternions(name)
        assert g.shape == (n, 4) and np.allclose(np.linalg.norm(g, axis=1), 1)
    assert [_proper_group_of_space_group(s) for s in (1, 20, 62, 99, 150, 194, 195, 225, 227)] == \
        ["1", "222", "222", "422", "32", "622", "23", "432", "432"]
    assert _resolve_proper_group("m-3m", None) == "432" and _resolve_proper_group(None, 194) == "622"
    with pytest.raises(ValueError):
        _resolve_proper_group(None, None)
    assert resolution_to_semi_edge_steps(2) == 67 and resolution_to_semi_edge_steps(20) == 7


@pytest.mark.gpu
@pytest.mark.parametrize("group,res", [("432", 10), ("622", 12), ("1", 25), ("32", 15)])
def test_device_fundamental_zone_equals_oracle(group, res):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffsims_b200.crystal import Rotation
    from diffsims_b200.generators import rotation_list_generators as rlg
    n = rlg.resolution_to_semi_edge_steps(res)
    q = so3.cubochoric_grid(n)
    sym = rlg.proper_point_group_quaternions(group)
    ref = q[so3.fundamental_zone_mask(q, sym)]
    euler, quat = rlg.fundamental_zone_device(res, point_group=group)
    got = quat.cpu().numpy()
    got[:, 1:] *= -1                                   # the kernel writes active quaternions (conjugates)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, atol=1e-12)   # same points, same (grid) order
    np.testing.assert_allclose(euler.cpu().numpy(), Rotation(ref).to_euler(degrees=True), atol=1e-8)
    lst = rlg.get_fundamental_zone_grid(res, point_group=group)
    assert isinstance(lst, list) and isinstance(lst[0], tuple) and len(lst) == len(ref)


@pytest.mark.gpu
def test_reference_test_get_grid_and_local_grid():
    """diffsims/tests/generators/test_rotation_list_generator.py:30-40, plus the local grid's defining properties."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffsims_b200.crystal import Rotation
    from diffsims_b200.generators.rotation_list_generators import (get_fundamental_zone_grid, get_list_from_orix, get_local_grid,
                                                                   local_grid_device)
    for grid in (get_local_grid(resolution=30, center=(0, 1, 0), grid_width=35),
                 get_fundamental_zone_grid(space_group=20, resolution=20)):
        assert isinstance(grid, list) and len(grid) > 0 and isinstance(grid[0], tuple)
    # every rotation of a local grid lies within grid_width of the centre, and the grid is the ball's share of SO(3)
    centre = Rotation.from_euler([[30, 40, 50]], degrees=True)
    _, quat = local_grid_device(resolution=5, center=(30, 40, 50), grid_width=25)
    q = quat.cpu().numpy()
    q[:, 1:] *= -1
    rel = (~centre * Rotation(q)) if len(q) == 1 else Rotation(np.array([(~centre * Rotation(r)).data[0] for r in q]))
    ang = np.rad2deg(so3.rotation_angle(rel.data))
    assert ang.max() <= 25 + 1e-9
    w = np.deg2rad(25)
    n_all = (2 * 27) ** 3                               # resolution 5 -> N = 27
    assert abs(len(q) / n_all - (w - np.sin(w)) / np.pi) < 0.01
    assert get_list_from_orix(Rotation(q[:3])) == [tuple(np.round(e, 2)) for e in Rotation(q[:3]).to_euler(degrees=True)]
