import sys, os, hashlib, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import world_from_freerun
from sand_crate_b200 import Crate
world, _ = world_from_freerun(sys.argv[3] if len(sys.argv) > 3 else "wave_machine")
prec, nt = sys.argv[1], int(sys.argv[2])
every = int(os.environ.get("EVERY", "10"))
np.random.seed(1234)
crate = Crate(world, precision=prec, noise="counter")
out = []
for t in range(nt):
    crate.physics_tick()
    if (t + 1) % every == 0:
        p = crate.particles; v = crate.particle_velocities
        out.append(f"{t+1}:{len(p)}:{hashlib.md5(p.tobytes() + v.tobytes()).hexdigest()[:8]}")
print(prec, " ".join(out))
