"""Developer aid: per-tick comparison of an N-rank strip run with the single-GPU run (same launch line as mgpu_check.py).
Prints, per tick, whether the gathered state is bit-identical and, if not, which particles differ and where they sit
relative to the cuts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sand_crate_b200.scenes import dam_break  # noqa: E402
from sand_crate_b200.strips import StripDomain, partition_rows, rows_of  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    n, ticks, rebalance = 200_000, 16, int(os.environ.get("SC_REBALANCE", "3"))
    transport = os.environ.get("SC_TRANSPORT", "nccl")
    cfg, pos, vel = dam_break(n)
    d = 2 * cfg.coefficients["particle_radius"]
    shifted = pos.copy()
    shifted[:, 1] -= 0.1
    cuts = partition_rows(rows_of(shifted, d), world)
    vel = vel + np.random.RandomState(3).randn(*vel.shape) * (d / cfg.coefficients["dt"]) * 0.3
    dom = StripDomain(cfg, pos, vel, rank=rank, world_size=world, precision="f64", noise="counter", noise_seed=5,
                      device=local, stream=stream.cuda_stream, transport=transport, rebalance_every=rebalance, cuts=cuts)
    single = None
    if rank == 0:
        single = StripDomain(cfg, pos, vel, rank=0, world_size=1, precision="f64", noise="counter", noise_seed=5,
                             device=local, stream=stream.cuda_stream)
    every = int(os.environ.get("SC_DEBUG_EVERY", "1"))  # compare every N ticks (host syncs only then)
    for t in range(ticks):
        cuts_before = list(dom.cuts)
        dom.physics_tick()
        if rank == 0 and (t + 1) % every:
            single.physics_tick()
        if (t + 1) % every:
            continue
        uid, gp, gv = dom.gather()
        st = dom.status()
        allst = [None] * world
        dist.all_gather_object(allst, (st["n_local"], st["overflow"], st["too_far"]))
        if rank == 0:
            single.physics_tick()
            suid, sp, sv = single.gather()
            same = np.array_equal(uid, suid) and np.array_equal(gp, sp) and np.array_equal(gv, sv)
            msg = f"[dbg] tick {t} cuts {cuts_before[1:-1]} -> {dom.cuts[1:-1]} same={same} n={len(uid)}/{len(suid)} st={allst}"
            if not same:
                if len(uid) != len(suid) or not np.array_equal(uid, suid):
                    u, c = np.unique(uid, return_counts=True)
                    dup = u[c > 1]
                    missing = np.setdiff1d(suid, uid)
                    msg += f" dup={len(dup)} missing={len(missing)}"
                    idx = np.searchsorted(suid, np.concatenate([dup, missing])[:8])
                    msg += f" rows={np.floor(sp[idx, 1] / d).astype(int).tolist()}"
                else:
                    bad = np.nonzero(np.any(gp != sp, 1) | np.any(gv != sv, 1))[0]
                    rows = np.floor(sp[bad, 1] / d).astype(int)
                    msg += f" ndiff={len(bad)} rows[min,max]={rows.min()},{rows.max()} sample_rows={rows[:10].tolist()}"
                    msg += f" maxdv={np.abs(gv[bad] - sv[bad]).max():.3e}"
            print(msg, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
