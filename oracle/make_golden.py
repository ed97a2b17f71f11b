"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by executing the UNMODIFIED reference
(David-Taub/sand_crate, loaded from /root/reference through oracle/ref_shim.py) in the build container.

    python -m oracle.make_golden            # everything (~30 minutes: wave_machine to tick 3000)
    python -m oracle.make_golden --quick    # skip wave_machine beyond tick 100

The fixtures are committed; the GPU box never needs /root/reference.
What is recorded:
  step_<config>_t<N>.npz    one tick: inputs after create/remove/apply_bodies_velocity (pos_in, vel_in, segments,
                            body_len, body_kin, coeffs), the reference's own np.random.rand draws in CSR order
                            (noise), every intermediate (pos_search, rows_sorted, order, nbr_count, nbr_idx,
                            pressure, tension_vec) and the outputs (pos_out, vel_out)
  freerun_<config>.npz      whole-trajectory checkpoints (pos, vel at given ticks) for the drop-in Crate test
  neighbors_cases.npz       detect_particle_collisions on the reference's own test inputs
                            (tests/test_distance.py) plus dense / duplicate / on-boundary cases that exercise
                            the 20-trim
  geometry_cases.npz        points_to_segments_distance and pad_segments
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_shim import RecordingCrate, load_reference  # noqa: E402

STEP_TICKS = {
    "stirring_cup": [1, 20, 60, 150, 199, 260, 400, 800, 1199],
    "wave_machine": [5, 100, 300, 500, 1000, 2000, 2999],
    "free_body": [],
}
MONITOR_SECTIONS = ("tension", "gravity", "pressure", "viscosity", "wall_bounce", "continuous_collision")
FREERUN_TICKS = {
    # up to the configs' own `ticks_to_record` (config/stirring_cup.yaml:3 = 1200, config/wave_machine.yaml:3 = 3000)
    "stirring_cup": [1, 5, 20, 40, 80, 200, 400, 800, 1200],
    "wave_machine": [1, 5, 20, 40, 500, 1000, 2000, 3000],
    "free_body": [1, 10, 30, 60],
}


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}  ({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


def free_body_config(ref):
    """wave_machine with its motored paddle replaced by a FREE body (rigid_body.py:18-49, "free" in
    BODY_TYPE_TO_CLASS): a bar that falls under gravity (crate.py:311-314) and spins - the one body type neither
    shipped config exercises."""
    cfg = ref.load_config(os.path.join(ref.config_dir, "wave_machine.yaml"))
    bodies = [b for b in cfg.world_config.rigid_bodies if "fixed" in b]
    bodies.append({"free": {"name": "bar", "segments": [[[-0.15, 0.0], [0.15, 0.0]], [[0.0, -0.05], [0.0, 0.05]]],
                            "position": [0.3, 0.75], "angular_clockwise_velocity": 0.8}})
    cfg.world_config.rigid_bodies = bodies
    return cfg


def record_config(ref, name, quick):
    cfg = free_body_config(ref) if name == "free_body" else ref.load_config(os.path.join(ref.config_dir, f"{name}.yaml"))
    rc = RecordingCrate(cfg.world_config)
    step_ticks = [t for t in STEP_TICKS[name] if not (quick and t > 100)]
    free_ticks = FREERUN_TICKS[name]
    free = {}
    last = max(step_ticks + free_ticks)
    t0 = time.time()
    for tick in range(1, last + 1):  # `tick` = value of crate.tick AFTER this call
        rec = rc.tick(record=tick in step_ticks)
        if rec is not None:
            save(f"step_{name}_t{tick}.npz", **rec)
        if tick in free_ticks:
            free[f"pos_t{tick}"] = rc.crate.particles.copy()
            free[f"vel_t{tick}"] = rc.crate.particle_velocities.copy()
            free[f"segments_t{tick}"] = rc.crate.segments.copy()
            free[f"pressure_t{tick}"] = np.asarray(rc.crate.particles_pressure, dtype=np.float64).copy()
            fm = rc.crate.force_monitor.context_to_velocity  # utils/force_monitor.py: EMA of mean |dv| per section
            free[f"monitor_t{tick}"] = np.array([float(fm[k]) for k in MONITOR_SECTIONS])
        if tick % 50 == 0:
            print(f"  {name}: tick {tick}/{last}  P={rc.crate.particle_count}  {time.time() - t0:.0f}s", flush=True)
    world = {"coefficients": cfg.world_config.coefficients, "particle_sources": cfg.world_config.particle_sources,
             "rigid_bodies": cfg.world_config.rigid_bodies}
    save(f"freerun_{name}.npz", ticks=np.array(free_ticks), world_json=np.array(json.dumps(world)), **free)


def neighbor_cases(ref):
    cd = ref.collision_detector
    cases = {}

    def add(tag, pts, d):
        pts = np.asarray(pts, dtype=np.float64)
        lists = cd.detect_particle_collisions(particles=pts, diameter=d)
        _, rows, order = cd.strip_sort_particles(particles=pts, diameter=d)
        counts = np.array([len(x) for x in lists], np.int32)
        idx = np.full((len(lists), 20), -1, np.int32)
        for i, x in enumerate(lists):
            idx[i, :len(x)] = x
        cases[f"{tag}__pts"] = pts
        cases[f"{tag}__d"] = np.float64(d)
        cases[f"{tag}__rows"] = np.asarray(rows, np.int64)
        cases[f"{tag}__order"] = np.asarray(order, np.int64)
        cases[f"{tag}__counts"] = counts
        cases[f"{tag}__idx"] = idx

    n = 35
    row = np.array([[i, 0] for i in range(n)])                                   # tests/test_distance.py:40
    grid = np.array([[i, j] for i in range(n) for j in range(n)])                # tests/test_distance.py:53
    for d in (0.5, 1, 2):
        add(f"row_d{d}", row, d)
        add(f"grid_d{d}", grid, d)
    rs = np.random.RandomState(0)
    add("random35_d0.1", rs.rand(n, 2), 0.1)                                     # tests/test_distance.py:61-63
    add("dense400_d0.2", rs.rand(400, 2), 0.2)           # ~45 neighbors each: trim fires everywhere
    add("dense1500_d0.05", rs.rand(1500, 2), 0.05)       # ~11 neighbors: trim fires sometimes
    pts = rs.rand(600, 2)
    pts[:, 0] = np.round(pts[:, 0] * 20) / 20            # many duplicate x, many exactly-d-apart pairs
    pts[:, 1] = np.round(pts[:, 1] * 40) / 40
    add("dupes600_d0.05", pts, 0.05)
    add("unitbox3000_d0.01", rs.rand(3000, 2), 0.01)     # sparse, the YAML configs' diameter
    pts = rs.rand(800, 2) * np.array([1.0, 0.03])        # 3 strips only, long rows
    add("flat800_d0.01", pts, 0.01)
    pts = rs.rand(300, 2) * 0.2 - 0.005                  # straddles 0: negative rows/cols
    add("neg300_d0.01", pts, 0.01)
    save("neighbors_cases.npz", **cases)


def geometry_cases(ref):
    geo = ref.geometry_utils
    rs = np.random.RandomState(1)
    p = np.array([[i, 0] for i in range(35)], dtype=np.float64)                  # tests/test_distance.py:17-18
    segs = np.array([[[i, -1], [i, 1]] for i in range(5)], dtype=np.float64)
    near, dist = geo.points_to_segments_distance(p, segs)
    p2 = rs.rand(200, 2)
    segs2 = rs.rand(9, 2, 2)
    near2, dist2 = geo.points_to_segments_distance(p2, segs2)
    pad2 = geo.pad_segments(segs2, 0.005)
    save("geometry_cases.npz", row_p=p, row_segs=segs, row_near=near, row_dist=dist, rnd_p=p2, rnd_segs=segs2,
         rnd_near=near2, rnd_dist=dist2, rnd_pad=pad2, rnd_pad_r=np.float64(0.005))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ref = load_reference()
    if a.only in ("", "neighbors"):
        neighbor_cases(ref)
    if a.only in ("", "geometry"):
        geometry_cases(ref)
    for name in ("stirring_cup", "wave_machine", "free_body"):
        if a.only in ("", name):
            print(f"recording {name}", flush=True)
            record_config(ref, name, a.quick)


if __name__ == "__main__":
    main()
