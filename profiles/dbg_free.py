import sys, os, json, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_parity import world_from_freerun
from sand_crate_b200 import Crate
world, g = world_from_freerun(sys.argv[1] if len(sys.argv) > 1 else "free_body")
crate = Crate(world)
print("capacity", crate._ctx.capacity)
for tick in range(1, 3):
    crate.physics_tick()
    print("tick", tick, "count", crate.particle_count, "cap", crate._ctx.capacity)
    print("uids", crate._ctx.get_uids())
    print("pos", crate.particles[:12])
    if f"pos_t{tick}" in g.files: print("gold", g[f"pos_t{tick}"][:12])
