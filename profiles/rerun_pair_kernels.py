import sys, ctypes as C, numpy as np
sys.path.insert(0, "/root/repo")
import torch
from sand_crate_b200 import _lib
from sand_crate_b200.scenes import dam_break
from bench import scene_params
world, pos, vel = dam_break(1_000_000)
ctx = _lib.Context(len(pos), _lib.PRECISION_MIXED)
ctx.set_params(**scene_params(world)); seg = np.array(world.rigid_bodies[0]["fixed"]["segments"]); ctx.set_walls(seg, [4], np.zeros((1,5)))
ctx.set_noise(_lib.NOISE_COUNTER, 0); ctx.set_state(pos, vel); ctx.step(300); ctx.synchronize()
L = _lib.load(); L.sc_debug_rerun.restype = C.c_double; L.sc_debug_rerun.argtypes = [C.c_void_p, C.c_int, C.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
import os
for which in ((4, 24, 5) if os.environ.get('SC_RERUN_REPEAT') else (4, 5)):
    warm = L.sc_debug_rerun(ctx._h, which, 20)
    cold = []
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize()
        cold.append(L.sc_debug_rerun(ctx._h, which, 1))
    print("K%d warm (back-to-back) %.1f us   cold (L2 flushed) %.1f us" % (which, 1e3 * warm, 1e3 * np.mean(cold)))
# in-pipeline order: K4 then K5 immediately
t = []
for _ in range(10):
    flush.fill_(1); torch.cuda.synchronize()
    a = L.sc_debug_rerun(ctx._h, 4, 1); b = L.sc_debug_rerun(ctx._h, 5, 1); t.append((a, b))
print("cold K4 then K5 right after: %.1f us, %.1f us" % tuple(1e3 * np.mean(t, 0)))
