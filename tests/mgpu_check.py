"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py

Every rank runs its strip of a dam-break scene (fp64 mode, counter noise, NCCL halo + migration exchange); rank 0
also runs the whole scene on one GPU.  The gathered N-rank state must equal the single-GPU state bit for bit, and
the single-GPU state must equal the oracle's (fp64 cases the oracle finishes in seconds).  SC_CHECK_SCALE=1 runs the
re-cutter at scale instead: dam-break 2M, 200 ticks, re-cut every 25 ticks, mixed precision.  Prints one line per check
and exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sand_crate_b200 import _lib  # noqa: E402
from sand_crate_b200.scenes import box_fill, dam_break  # noqa: E402
from sand_crate_b200.strips import StripDomain  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ok = True
    transport = os.environ.get("SC_TRANSPORT", "nccl")
    cases = [(dam_break, 200_000, 12, "f64", 0), (box_fill, 300_000, 8, "f64", 0),
             (dam_break, 200_000, 12, "mixed", 0), (dam_break, 200_000, 16, "f64", 3)]
    if os.environ.get("SC_CHECK_SCALE"):  # the re-cutter at the size it was built for
        cases = [(dam_break, 2_000_000, 200, "mixed", 25), (dam_break, 200_000, 12, "f64", 0)]
    only = os.environ.get("SC_CHECK_ONLY")  # developer aid: run some of the cases, by index ("2,3")
    if only is not None:
        cases = tuple(cases[int(k)] for k in only.split(","))
    for maker, n, ticks, precision, rebalance in cases:
        world_cfg, pos, vel = maker(n)
        cuts = None
        if rebalance:  # start from the equal-count cuts of ANOTHER scene (the column 0.1 higher): unbalanced here
            from sand_crate_b200.strips import partition_rows, rows_of
            shifted = pos.copy()
            shifted[:, 1] -= 0.1
            cuts = partition_rows(rows_of(shifted, 2 * world_cfg.coefficients["particle_radius"]), world)
        # fast particles (0.3 d per tick): rows change hands every tick
        vel = vel + np.random.RandomState(3).randn(*vel.shape) * (
            2.0 * world_cfg.coefficients["particle_radius"] / world_cfg.coefficients["dt"]) * 0.3
        dom = StripDomain(world_cfg, pos, vel, rank=rank, world_size=world, precision=precision, noise="counter",
                          noise_seed=5, device=local, stream=stream.cuda_stream, transport=transport,
                          rebalance_every=rebalance, cuts=cuts, adaptive_rebalance=False)
        cuts0 = list(dom.cuts)
        own0 = set(dom.owned()[0].tolist())
        dom.step(ticks)
        st = dom.status()
        uid, gp, gv = dom.gather()
        moved = len(set(dom.owned()[0].tolist()) - own0)
        moved_all = [None] * world
        dist.all_gather_object(moved_all, (moved, st))
        if rank == 0:
            single = StripDomain(world_cfg, pos, vel, rank=0, world_size=1, precision=precision, noise="counter",
                                 noise_seed=5, device=local, stream=stream.cuda_stream)
            single.step(ticks)
            suid, sp, sv = single.gather()
            same = np.array_equal(uid, suid) and np.array_equal(gp, sp) and np.array_equal(gv, sv)
            if not same:  # say what differs
                if not np.array_equal(uid, suid):
                    u, c = np.unique(uid, return_counts=True)
                    print(f"[mgpu]   uid sets differ: {len(uid)} vs {len(suid)}, duplicated {int((c > 1).sum())}, "
                          f"missing {len(np.setdiff1d(suid, uid))}", flush=True)
                else:
                    bad = np.nonzero(np.any(gp != sp, 1) | np.any(gv != sv, 1))[0]
                    print(f"[mgpu]   {len(bad)} particles differ, max |dpos| {np.abs(gp - sp).max():.3e}, "
                          f"max |dvel| {np.abs(gv - sv).max():.3e}", flush=True)
            vs_oracle = ""
            if precision == "f64" and n <= 300_000:  # the single-GPU state against the oracle, bit for bit
                from oracle import oracle as O
                c = world_cfg.coefficients
                cv = np.array([c["dt"], c["particle_radius"], c["wall_collision_decay"], c["pressure_amplifier"],
                               c["ignored_pressure"], c["collider_noise_level"], c["viscosity"], c["surface_smoothing"],
                               c["target_pressure"], c["gravity"][0], c["gravity"][1]])
                seg = np.array(world_cfg.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
                rp, rv = pos.copy(), vel.copy()
                for tick in range(ticks):
                    out = O.step(cv, rp, rv, seg, [4], np.zeros((1, 5)), noise_mode=1, tkey=O.tick_key(5, tick),
                                 want_all=False)
                    rp, rv = out["pos_out"], out["vel_out"]
                exact = np.array_equal(sp, rp) and np.array_equal(sv, rv)
                vs_oracle = f"; single GPU == oracle after {ticks} ticks: {exact}"
                same = same and exact
            flags = any(s["overflow"] or s["too_far"] for _, s in moved_all)
            print(f"[mgpu] transport={transport} {maker.__name__} n={n} ticks={ticks} {precision} ranks={world}: "
                  f"bit-identical to single GPU = {same}; migrated = {[m for m, _ in moved_all]}; "
                  f"local = {[s['n_local'] for _, s in moved_all]}; flags = {flags}; "
                  f"rebalance_every = {rebalance}, cuts moved = {dom.cuts != cuts0}{vs_oracle}", flush=True)
            ok = ok and same and not flags and sum(m for m, _ in moved_all) > 0
            single.close()
        dom.close()
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(res, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(res.item()) else 1)


if __name__ == "__main__":
    main()
