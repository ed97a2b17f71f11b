"""Developer aid: does a second domain in the same process (recycled device memory) give the same bits as a fresh one?"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sand_crate_b200.scenes import dam_break
from sand_crate_b200.strips import StripDomain
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
def run(ticks, n=200_000):
    cfg, pos, vel = dam_break(n)
    d = 2 * cfg.coefficients["particle_radius"]
    vel = vel + np.random.RandomState(3).randn(*vel.shape) * (d / cfg.coefficients["dt"]) * 0.3
    s = StripDomain(cfg, pos, vel, rank=0, world_size=1, precision="f64", noise="counter", noise_seed=5, device=0,
                    stream=stream.cuda_stream)
    s.step(ticks)
    out = s.gather()
    s.close()
    return out
mode = sys.argv[1]
if mode == "fresh":
    u, p, v = run(16); np.savez("gpurun_out/dbg_seq_fresh.npz", u=u, p=p, v=v)
else:
    run(12)
    u, p, v = run(16)
    g = np.load("gpurun_out/dbg_seq_fresh.npz")
    print("second-in-process == fresh:", np.array_equal(u, g["u"]) and np.array_equal(p, g["p"]) and np.array_equal(v, g["v"]),
          "ndiff", int((np.any(p != g["p"], 1)).sum()) if len(p) == len(g["p"]) else -1)
