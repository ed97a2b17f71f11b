"""Diagnostic: the hydrostatic regime of the 64M dam break (8 300 cell rows of liquid) on ONE GPU - a column of the same
height and particle diameter but only `cols` lattice columns wide.  Prints, every `every` ticks, the fastest particle in
cell rows per tick (the strip exchange needs < halo_rows = 4), the largest neighbor count, NaNs and the live count."""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
from sand_crate_b200 import _lib  # noqa: E402
from sand_crate_b200.scenes import LATTICE_FRACTION, _lattice, _world  # noqa: E402

cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 250
spacing = math.sqrt(0.5 * 0.98 / 64_000_000)
d = spacing / LATTICE_FRACTION
n = cols * int(0.98 / spacing)
pts = _lattice(n, d, d + cols * spacing, 1.0 - d, spacing, 42)
world = _world(n, d)
c = world.coefficients
ctx = _lib.Context(n, _lib.PRECISION_MIXED)
ctx.set_params(dt=c["dt"], particle_radius=c["particle_radius"], wall_collision_decay=c["wall_collision_decay"],
               pressure_amplifier=c["pressure_amplifier"], ignored_pressure=c["ignored_pressure"],
               collider_noise_level=c["collider_noise_level"], viscosity=c["viscosity"],
               surface_smoothing=c["surface_smoothing"], target_pressure=c["target_pressure"],
               gravity_x=c["gravity"][0], gravity_y=c["gravity"][1])
seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
ctx.set_walls(seg, [4], np.zeros((1, 5)))
ctx.set_noise(_lib.NOISE_COUNTER, 0)
ctx.set_state(pts, np.zeros_like(pts))
print(f"n={n} d={d:.4e} dt={c['dt']:.4e} rows={int(0.98 / d)}", flush=True)
for t in range(0, ticks, every):
    ctx.step(every)
    pos, vel, prs = ctx.get_state()
    speed = np.sqrt((vel ** 2).sum(1))
    bad = int((~np.isfinite(pos)).any(1).sum())
    k = int(np.nanargmax(speed))
    counts, _ = ctx.get_neighbors(len(pos))
    print(f"tick {t + every}: n={len(pos)} max rows/tick={np.nanmax(speed) * c['dt'] / d:.3f} at y={pos[k, 1]:.3f} x={pos[k, 0]:.4f} "
          f"p99.99={np.nanpercentile(speed, 99.99) * c['dt'] / d:.3f} maxK={counts.max()} meanK={counts.mean():.2f} "
          f"pmax={np.nanmax(prs):.2f} nonfinite={bad} untiled={ctx.untiled_blocks()}", flush=True)
