import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import world_from_freerun
from sand_crate_b200 import Crate
world, _ = world_from_freerun("wave_machine")
seq = sys.argv[1].split(",")
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
do_close = (sys.argv[3] != "noclose") if len(sys.argv) > 3 else True
keep = []
for prec in seq:
    np.random.seed(1234)
    crate = Crate(world, precision=prec, noise="counter")
    ok = True
    for t in range(nt):
        try:
            crate.physics_tick()
        except Exception as e:
            print(prec, "FAILED at tick", t, str(e)[:160], flush=True); ok = False; break
    if not ok: break
    print(prec, "ok", crate.particle_count, flush=True)
    if do_close: crate.close()
    else: keep.append(crate)
