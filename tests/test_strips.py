"""Strip decomposition, CPU side: partitioning and the per-tick pack / exchange / unpack protocol over
`torch.distributed` (gloo, world_size 2 and 3), with the device work done by the oracle-backed test double.  The
N-rank run must reproduce the single-domain oracle run bit for bit - that is what proves the halo width, the
migration rule and the ghost retention rule (csrc/sc_dist.cuh) before any GPU is involved."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import oracle as O
from oracle_backend import OracleContext
from sand_crate_b200.scenes import box_fill, dam_break
from sand_crate_b200.strips import HALO_ROWS, StripDomain, partition_rows, rows_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_partition_rows_balances_and_respects_halo():
    world, pos, _ = dam_break(40000)
    d = 2 * world.coefficients["particle_radius"]
    rows = rows_of(pos, d)
    for n in (2, 3, 4):
        cuts = partition_rows(rows, n)
        assert len(cuts) == n + 1 and cuts[0] < rows.min() and cuts[-1] > rows.max()
        counts = [int(((rows >= cuts[k]) & (rows < cuts[k + 1])).sum()) for k in range(n)]
        assert sum(counts) == len(rows)
        assert max(counts) - min(counts) <= 0.1 * len(rows) / n
        assert all(cuts[k + 1] - cuts[k] >= 2 * HALO_ROWS for k in range(1, n - 1))
    with pytest.raises(ValueError):
        partition_rows(rows[rows < rows.min() + 10], 4)


def _coeff_vec(c):
    return np.array([c["dt"], c["particle_radius"], c["wall_collision_decay"], c["pressure_amplifier"],
                     c["ignored_pressure"], c["collider_noise_level"], c["viscosity"], c["surface_smoothing"],
                     c["target_pressure"], c["gravity"][0], c["gravity"][1]], dtype=np.float64)


def _worker(rank, world_size, port, scene, n, ticks, out_dir, rebalance_every=0, adaptive=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        world, pos, vel = (box_fill if scene == "box_fill" else dam_break)(n)
        cuts = None
        if scene == "dam_break_shifted":   # start from a partition that is wrong for the scene: cuts must move
            pos = pos.copy()
            pos[:, 1] -= 0.1
        if scene == "dam_break_wrong_cuts":  # the equal-count cuts of ANOTHER scene (tests/mgpu_check.py, case 3)
            from sand_crate_b200.strips import partition_rows, rows_of
            shifted = pos.copy()
            shifted[:, 1] -= 0.1
            cuts = partition_rows(rows_of(shifted, 2 * world.coefficients["particle_radius"]), world_size)
        vel = vel + np.random.RandomState(7).randn(*vel.shape) * 3.0  # fast particles: migration every tick
        dom = StripDomain(world, pos, vel, rank=rank, world_size=world_size, precision="f64", noise="counter",
                          noise_seed=11, context_factory=OracleContext, tensor_device=torch.device("cpu"),
                          rebalance_every=rebalance_every, cuts=cuts, adaptive_rebalance=adaptive)
        cuts0 = list(dom.cuts)
        migrated = 0
        for _ in range(ticks):
            before = set(dom.ctx.dist_get_owned()[2].tolist())
            dom.physics_tick()
            migrated += len(set(dom.ctx.dist_get_owned()[2].tolist()) - before)
        st = dom.status()
        uid, p, v = dom.gather()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), uid=uid, pos=p, vel=v)
        stats = [None] * world_size
        dist.all_gather_object(stats, (migrated, st["too_far"], st["n_local"], int(dom.cuts != cuts0),
                                       len(dom.ctx.dist_get_owned()[2])))
        logs = [None] * world_size
        dist.all_gather_object(logs, [tuple(e[:3]) for e in dom.rebalance_log])
        if rank == 0:
            np.save(os.path.join(out_dir, "stats.npy"), np.array(stats, dtype=np.int64))
            with open(os.path.join(out_dir, "recuts.json"), "w") as f:
                json.dump(logs, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size,scene", [(2, "dam_break"), (3, "box_fill")])
def test_strip_protocol_matches_single_domain(tmp_path, world_size, scene):
    n, ticks = 6000, 6
    mp.spawn(_worker, args=(world_size, _free_port(), scene, n, ticks, str(tmp_path)), nprocs=world_size, join=True)
    got = np.load(tmp_path / "out.npz")
    stats = np.load(tmp_path / "stats.npy")
    # reference: the same ticks on the whole scene in one domain
    world, pos, vel = (dam_break if scene == "dam_break" else box_fill)(n)
    vel = vel + np.random.RandomState(7).randn(*vel.shape) * 3.0
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = _coeff_vec(world.coefficients)
    uid = np.arange(n, dtype=np.uint32)
    for tick in range(ticks):
        pos, vel, mask = O.remove_particles(pos, vel, world.coefficients["particle_radius"])
        uid = uid[~mask]
        out = O.step(cv, pos, vel, seg, [4], np.zeros((1, 5)), noise_mode=1, tkey=O.tick_key(11, tick), uid=uid,
                     want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
    assert np.array_equal(got["uid"], uid)
    assert np.array_equal(got["pos"], pos) and np.array_equal(got["vel"], vel)
    assert stats[:, 0].sum() > 0, "the test scene must exercise migration"
    assert not stats[:, 1].any(), "no particle may cross a whole halo in one tick"


def test_cuts_from_histogram_balance_spacing_and_agreement_with_partition_rows():
    from sand_crate_b200.strips import HALO_ROWS, cuts_from_histogram, partition_rows
    rng = np.random.RandomState(5)
    rows = np.concatenate([rng.randint(-3, 400, 20000), rng.randint(300, 400, 30000)])  # a dense bottom layer
    for nranks in (2, 3, 8):
        lo = int(rows.min())
        hist = np.bincount(rows - lo)
        cuts = cuts_from_histogram(hist, lo, nranks)
        assert cuts == partition_rows(rows, nranks)
        assert cuts[0] < -(1 << 61) and cuts[-1] > (1 << 61) and len(cuts) == nranks + 1
        inner = cuts[1:-1]
        assert all(b - a >= 2 * HALO_ROWS for a, b in zip([lo] + inner[:-1], inner))
        edges = [0] + [c - lo for c in inner] + [len(hist)]
        share = np.array([hist[a:b].sum() for a, b in zip(edges[:-1], edges[1:])])
        assert share.sum() == len(rows)
        assert np.all(np.abs(share - len(rows) / nranks) <= hist.max()), share   # within one row of equal
    # weights instead of counts (the device's histogram: base + pair count per particle): same rule
    w = np.where(np.arange(800) < 600, 7000 * 6, 19000 * 12).astype(np.uint64)
    cuts = cuts_from_histogram(w, -2, 8)
    edges = [0] + [c + 2 for c in cuts[1:-1]] + [800]
    share = np.array([int(w[a:b].sum()) for a, b in zip(edges[:-1], edges[1:])])
    assert share.max() - share.min() <= 2 * int(w.max()), share
    # an empty histogram and one with all the weight in a single row still give legal, ordered cuts
    for hist in (np.zeros(100, np.int64), np.bincount([57] * 1000, minlength=100)):
        cuts = cuts_from_histogram(hist, 0, 4)
        assert all(b - a >= 2 * HALO_ROWS for a, b in zip(cuts[1:-2], cuts[2:-1])), cuts
    assert cuts_from_histogram(np.ones(10), 0, 1)[1:] == [cuts[-1]]


def test_rebalance_interval_rule():
    from sand_crate_b200.strips import next_rebalance_interval as nxt
    assert nxt(250, 40, 250) == 125 and nxt(125, 40, 250) == 62 and nxt(31, 40, 250) == 25 and nxt(25, 40, 250) == 25
    assert nxt(250, 0, 250) == 500 and nxt(500, 2, 250) == 1000 and nxt(1000, 0, 250) == 1000
    assert nxt(250, 5, 250) == 250
    assert nxt(2, 40, 2) == 2 and nxt(3, 9, 3) == 3, "a caller who starts below 25 is never slowed down by a far-off cut"
    assert nxt(2000, 0, 2000) == 2000
    assert nxt(250, 40, 250, floor=80) == 125 and nxt(125, 40, 250, floor=80) == 80 and nxt(80, 40, 250, floor=80) == 80
    assert nxt(50, 40, 250, floor=80) == 80, "an expensive re-cut lengthens the interval even when the cuts are far off"
    assert nxt(25, 13, 250, floor=111) == 111 and nxt(500, 0, 250, floor=5000) == 1000


def test_sliding_cuts_rebalance_a_collapsing_column(tmp_path):
    """Re-balancing: the dam-break column collapses, rows change population, the cuts slide (<= halo - 2 rows per
    tick) towards equal counts - and the result is still bit-identical to the single-domain run."""
    n, ticks, world_size = 6000, 14, 3
    mp.spawn(_worker, args=(world_size, _free_port(), "dam_break_shifted", n, ticks, str(tmp_path), 2),
             nprocs=world_size, join=True)
    got = np.load(tmp_path / "out.npz")
    stats = np.load(tmp_path / "stats.npy")
    world, pos, vel = dam_break(n)
    pos = pos.copy()
    pos[:, 1] -= 0.1
    vel = vel + np.random.RandomState(7).randn(*vel.shape) * 3.0
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = _coeff_vec(world.coefficients)
    uid = np.arange(n, dtype=np.uint32)
    for tick in range(ticks):
        pos, vel, mask = O.remove_particles(pos, vel, world.coefficients["particle_radius"])
        uid = uid[~mask]
        out = O.step(cv, pos, vel, seg, [4], np.zeros((1, 5)), noise_mode=1, tkey=O.tick_key(11, tick), uid=uid,
                     want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
    assert np.array_equal(got["uid"], uid) and np.array_equal(got["pos"], pos) and np.array_equal(got["vel"], vel)
    assert stats[:, 3].all(), "the cuts must have moved"
    assert not stats[:, 1].any()
    owned = stats[:, 4]
    assert owned.max() - owned.min() < 0.25 * n / world_size, owned


def test_adaptive_recut_interval_same_on_every_rank_and_still_bit_identical(tmp_path):
    """The adaptive re-cut interval (measured idle and tick times ride in the histogram's all-reduce): every rank must
    log the same (tick, shift, interval) sequence - a rank that decided differently would miss the next collective -
    and where the cuts are at any tick never changes the result."""
    n, ticks, world_size = 6000, 14, 3
    mp.spawn(_worker, args=(world_size, _free_port(), "dam_break_shifted", n, ticks, str(tmp_path), 2, True),
             nprocs=world_size, join=True)
    got = np.load(tmp_path / "out.npz")
    logs = json.load(open(tmp_path / "recuts.json"))
    assert len(logs[0]) >= 1 and logs[0][0][0] == 2, logs[0]
    assert all(log == logs[0] for log in logs), logs
    assert all(e[2] >= 2 for e in logs[0])
    world, pos, vel = dam_break(n)
    pos = pos.copy()
    pos[:, 1] -= 0.1
    vel = vel + np.random.RandomState(7).randn(*vel.shape) * 3.0
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = _coeff_vec(world.coefficients)
    uid = np.arange(n, dtype=np.uint32)
    for tick in range(ticks):
        pos, vel, mask = O.remove_particles(pos, vel, world.coefficients["particle_radius"])
        uid = uid[~mask]
        out = O.step(cv, pos, vel, seg, [4], np.zeros((1, 5)), noise_mode=1, tkey=O.tick_key(11, tick), uid=uid,
                     want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
    assert np.array_equal(got["uid"], uid) and np.array_equal(got["pos"], pos) and np.array_equal(got["vel"], vel)


def test_sliding_cuts_four_ranks_two_interior(tmp_path):
    """Four strips (two interior ranks with two moving cuts each) started from the cuts of another scene and re-cut
    every 3 ticks: still bit-identical to the single-domain run (the case tests/mgpu_check.py runs on 4 GPUs)."""
    n, ticks, world_size = 8000, 10, 4
    mp.spawn(_worker, args=(world_size, _free_port(), "dam_break_wrong_cuts", n, ticks, str(tmp_path), 3),
             nprocs=world_size, join=True)
    got = np.load(tmp_path / "out.npz")
    stats = np.load(tmp_path / "stats.npy")
    world, pos, vel = dam_break(n)
    vel = vel + np.random.RandomState(7).randn(*vel.shape) * 3.0
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = _coeff_vec(world.coefficients)
    uid = np.arange(n, dtype=np.uint32)
    for tick in range(ticks):
        pos, vel, mask = O.remove_particles(pos, vel, world.coefficients["particle_radius"])
        uid = uid[~mask]
        out = O.step(cv, pos, vel, seg, [4], np.zeros((1, 5)), noise_mode=1, tkey=O.tick_key(11, tick), uid=uid,
                     want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
    assert np.array_equal(got["uid"], uid) and np.array_equal(got["pos"], pos) and np.array_equal(got["vel"], vel)
    assert stats[:, 3].all(), "the cuts must have moved"
    assert not stats[:, 1].any()
