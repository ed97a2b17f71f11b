"""Headless runner: `python -m sand_crate_b200.run config/stirring_cup.yaml --ticks 1200 --out run.npz`

The reference's only entry point opens a PyGame window (`src/main.py:19-23` -> `Playback.run_live_simulation`,
playback.py:51-65) and can only record rendered frames (GIF/AVI, playback.py:109-138; the particle dumps at 112-113
are commented out).  This runner drives the same `Crate(world_config).physics_tick()` loop without a display and
records what a renderer needs - positions, pressures and wall segments every `--every` ticks - so a run made on a GPU
box can be replayed anywhere.  Output: one compressed .npz with
`ticks`, `count[t]`, `offsets[t + 1]`, `pos` (sum(count) x 2), `pressure`, `segments` (T x S x 2 x 2), `config_yaml`.
"""
from __future__ import annotations

import argparse
import json
import sys
import time

import numpy as np
import yaml

from . import Crate, load_config


def run(config_path, ticks=None, every=10, out=None, precision="f64", noise="reference", device=0, quiet=False):
    cfg = load_config(config_path)
    ticks = int(ticks if ticks is not None else cfg.playback_config.ticks_to_record)
    crate = Crate(cfg.world_config, precision=precision, noise=noise, device=device)
    rec_ticks, counts, pos, prs, segs = [], [], [], [], []
    t0 = time.perf_counter()
    steps = 0
    for tick in range(1, ticks + 1):
        crate.physics_tick()
        steps += crate.particle_count if noise == "reference" else 0
        if out and (tick % every == 0 or tick == ticks):
            rec_ticks.append(tick)
            counts.append(crate.particle_count)
            pos.append(crate.particles.copy())
            prs.append(crate.particles_pressure.copy())
            segs.append(crate.segments.copy())
    n_final = crate.particle_count  # synchronises
    elapsed = time.perf_counter() - t0
    summary = {"config": str(config_path), "ticks": ticks, "particles_final": int(n_final),
               "seconds": round(elapsed, 3), "ms_per_tick": round(1e3 * elapsed / max(ticks, 1), 4),
               "precision": precision, "noise": noise}
    if steps:
        summary["particle_steps_per_sec"] = round(steps / elapsed, 1)
    if out:
        offsets = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        with open(config_path) as f:
            config_yaml = f.read()
        np.savez_compressed(out, ticks=np.array(rec_ticks), count=np.array(counts), offsets=offsets,
                            pos=np.concatenate(pos) if pos else np.zeros((0, 2)),
                            pressure=np.concatenate(prs) if prs else np.zeros(0), segments=np.array(segs),
                            config_yaml=np.array(config_yaml), summary=np.array(json.dumps(summary)))
        summary["out"] = str(out)
    if not quiet:
        print(json.dumps(summary))
    crate.close()
    return summary


def load_recording(path):
    """Frames of a recording made by `run`: yields (tick, positions, pressures, segments)."""
    z = np.load(path, allow_pickle=False)
    off = z["offsets"]
    for k, tick in enumerate(z["ticks"]):
        yield int(tick), z["pos"][off[k]:off[k + 1]], z["pressure"][off[k]:off[k + 1]], z["segments"][k]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("config")
    ap.add_argument("--ticks", type=int, default=None, help="default: playback.ticks_to_record of the config")
    ap.add_argument("--every", type=int, default=10, help="record every N ticks")
    ap.add_argument("--out", default=None)
    ap.add_argument("--precision", default="f64", choices=["f64", "mixed"])
    ap.add_argument("--noise", default="reference", choices=["reference", "counter", "none"])
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    run(a.config, a.ticks, a.every, a.out, a.precision, a.noise, a.device)


if __name__ == "__main__":
    main(sys.argv[1:])
