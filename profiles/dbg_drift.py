import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import world_from_freerun
from sand_crate_b200 import Crate
world, _ = world_from_freerun("wave_machine")
prec = sys.argv[1]
np.random.seed(1234)
crate = Crate(world, precision=prec, noise="counter")
for t in range(1000):
    try:
        crate.physics_tick()
        if t % 50 == 0 or t > 0 and False:
            n = crate.particle_count
            print(prec, "tick", t, "n", n, "cap", crate._ctx.capacity, flush=True)
    except Exception as e:
        print(prec, "FAILED at tick", t, str(e)[:200]); break
else:
    print(prec, "ok", crate.particle_count)
