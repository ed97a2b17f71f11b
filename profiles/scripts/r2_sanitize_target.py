"""What the compute-sanitizer runs execute: __graft_entry__.smoke() (one recorded wave_machine tick in fp64 and mixed
mode through the untiled kernels and the host-noise protocol, one 200k dam-break tick through the tiled production
kernels) plus 40 ticks of wave_machine in the production mode (device-side sources, tiled kernels, walls, removal)."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import __graft_entry__ as g  # noqa: E402
from conftest import world_from_freerun  # noqa: E402
from sand_crate_b200 import Crate  # noqa: E402

g.smoke()
world, _ = world_from_freerun("wave_machine")
crate = Crate(world, precision="mixed", noise="counter")
for _ in range(40):
    crate.physics_tick()
assert np.isfinite(crate.particles).all() and crate.particle_count > 100
print("sanitize target ok:", crate.particle_count, "particles after 40 ticks")
