"""CPU tests: the oracle (oracle/step_oracle.c) pinned bit-for-bit against golden vectors recorded from the
UNMODIFIED reference (oracle/make_golden.py), and against the reference's own known-answer tests
(/root/reference/tests/test_distance.py, 8 cases) restated on the oracle API."""
import itertools
from math import ceil, floor

import numpy as np
import pytest

from conftest import golden, neighbor_case_tags, step_goldens
from oracle import oracle as O


@pytest.mark.parametrize("tag", neighbor_case_tags())
def test_neighbor_search_matches_reference(tag):
    g = golden("neighbors_cases.npz")
    rows, order, counts, idx = O.detect_particle_collisions(g[f"{tag}__pts"], float(g[f"{tag}__d"]))
    assert np.array_equal(rows, g[f"{tag}__rows"])
    assert np.array_equal(order, g[f"{tag}__order"])
    assert np.array_equal(counts, g[f"{tag}__counts"])
    assert np.array_equal(idx, g[f"{tag}__idx"])


def test_geometry_matches_reference():
    g = golden("geometry_cases.npz")
    for k in ("row", "rnd"):
        near, dist = O.points_to_segments_distance(g[f"{k}_p"], g[f"{k}_segs"])
        assert np.array_equal(near, g[f"{k}_near"]) and np.array_equal(dist, g[f"{k}_dist"])
    assert np.array_equal(O.pad_segments(g["rnd_segs"], float(g["rnd_pad_r"])), g["rnd_pad"])


@pytest.mark.parametrize("name", step_goldens())
def test_step_matches_reference_bit_for_bit(name):
    g = golden(name)
    out = O.step(g["coeffs"], g["pos_in"], g["vel_in"], g["segments"], g["body_len"], g["body_kin"],
                 noise_mode=2, noise=g["noise"])
    for key in ("pos_search", "nbr_count", "nbr_idx", "pressure", "tension_vec", "vel_out", "pos_out"):
        assert np.array_equal(out[key], g[key]), key
    rows, order, _, _ = O.detect_particle_collisions(g["pos_search"], 2 * float(g["coeffs"][1]))
    assert np.array_equal(rows, g["rows_sorted"]) and np.array_equal(order, g["order"])


def test_goldens_exercise_the_quirks():
    """The fixtures must actually contain what parity is fragile on: walls, CCD scaling, K >= 8 (pairwise sum)."""
    seen = {"walls": 0, "ccd": 0, "k8": 0, "multi_contact": 0}
    for name in step_goldens():
        g = golden(name)
        out = O.step(g["coeffs"], g["pos_in"], g["vel_in"], g["segments"], g["body_len"], g["body_kin"],
                     noise_mode=2, noise=g["noise"])
        seen["walls"] += int((out["wall_count"] > 0).sum())
        seen["multi_contact"] += int((out["wall_count"] > 1).sum())
        seen["ccd"] += int((out["ccd_factor"] < 1).sum())
        seen["k8"] += int((g["nbr_count"] >= 8).sum())
    assert all(v > 0 for v in seen.values()), seen


# ---- the reference's own tests (tests/test_distance.py), restated on the oracle ---------------------------------
PARTICLES_COUNT, SEGMENTS_COUNT = 35, 5


def test_row_distance():  # test_distance.py:16-25
    p = np.array([[i, 0] for i in range(PARTICLES_COUNT)])
    segments = np.array([[[i, -1], [i, 1]] for i in range(SEGMENTS_COUNT)])
    _, distances = O.points_to_segments_distance(p, segments)
    assert distances.shape == (PARTICLES_COUNT, SEGMENTS_COUNT)
    for i in range(SEGMENTS_COUNT):
        for j in range(PARTICLES_COUNT):
            assert distances[j, i] == abs(j - i)


def _lists(p, d):
    _, _, counts, idx = O.detect_particle_collisions(p, d)
    return O.neighbor_lists(counts, idx)


@pytest.mark.parametrize("diameter,lo,hi", [(0.5, 0, 0), (1, 1, 2), (2, 2, 4)])
def test_collider_particles_row(diameter, lo, hi):  # test_distance.py:38-48
    p = np.array([[i, 0] for i in range(PARTICLES_COUNT)])
    nb = _lists(p, diameter)
    for i, n in enumerate(nb):
        for j in range(max(0, ceil(i - diameter)), min(floor(i + diameter), PARTICLES_COUNT - 1)):
            assert j in n or j == i
    assert len(nb) == p.shape[0]
    assert all(lo <= len(n) <= hi for n in nb)
    assert any(lo == len(n) for n in nb) and any(hi == len(n) for n in nb)


@pytest.mark.parametrize("diameter,lo,hi", [(0.5, 0, 0), (1, 2, 4), (2, 5, 12)])
def test_collider_particles_grid(diameter, lo, hi):  # test_distance.py:51-58
    p = np.array([[i, j] for i, j in itertools.product(range(PARTICLES_COUNT), range(PARTICLES_COUNT))])
    nb = _lists(p, diameter)
    assert len(nb) == p.shape[0]
    assert all(lo <= len(n) <= hi for n in nb)
    assert any(lo == len(n) for n in nb) and any(hi == len(n) for n in nb)


def test_collider_random_space():  # test_distance.py:61-70
    diameter = 0.1
    ps = np.random.RandomState(0).rand(PARTICLES_COUNT, 2)
    nb = _lists(ps, diameter)
    for i, p in enumerate(ps):
        if nb[i]:
            assert all(np.linalg.norm(ps[nb[i]] - p, axis=1) <= diameter * 3)


def test_neighbor_set_is_brute_force_when_untrimmed():
    rs = np.random.RandomState(3)
    ps, d = rs.rand(500, 2), 0.04
    nb = _lists(ps, d)
    for i in range(len(ps)):
        dist = np.sqrt(((ps - ps[i]) ** 2).sum(1))
        want = set(np.where(dist <= d)[0]) - {i}
        assert len(nb[i]) < 20 and set(nb[i]) == want


def test_remove_particles_is_stable():
    pos = np.array([[0.5, 0.5], [-0.1, 0.5], [0.2, 1.006], [0.9, 0.1], [1.2, 0.3], [-0.005, 1.005]])
    vel = np.arange(12, dtype=float).reshape(6, 2)
    p, v, mask = O.remove_particles(pos, vel, 0.005)
    assert mask.tolist() == [False, True, True, False, True, False]
    assert np.array_equal(p, pos[~mask]) and np.array_equal(v, vel[~mask])


def test_counter_noise_is_uniform_and_keyed():
    k = O.tick_key(0, 7)
    assert k != O.tick_key(0, 8) and k != O.tick_key(1, 7)
    u = np.array([O.pair_noise(k, i, j) for i in range(40) for j in range(40)])
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    assert O.pair_noise(k, 3, 4) != O.pair_noise(k, 4, 3)  # per DIRECTED pair, like the reference
