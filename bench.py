#!/usr/bin/env python
"""bench.py - particle-steps/sec of the SandCrate step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--particles P] [--scene dam_break|box_fill]
    python bench.py --impl reference ...        # the CPU port of the reference step on the host cores

One "step" = one `physics_tick` over the whole synthetic scene.  N = 1: dam-break, 1M particles (BASELINE.json
configs[2]).  Timed with CUDA events on the stream the kernels are launched on; L2 is flushed between timed steps
(the 1M scene's working set is smaller than the 126 MB L2).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
# algorithmic bytes per particle per kernel for the mixed layout (pos f64x2, vel f32x2, id, p, s,
# cell id), each record moved once per kernel.  These are the figures `roofline.achieved` is computed from.
SURVEY_BYTES = {"prepass_wall_key": 20, "place": 14, "rank_gather": 56, "density": 28, "force_integrate": 60}


# What THIS design moves per particle per launch: the same accounting plus the records K4 hands
# to K5 (8 bytes per directed pair, K = measured mean pairs per particle) and the packed (p, s) record.
def algo_bytes(K):
    return {
        "prepass_wall_key": 16 + 4 + 4,                       # R pos; W key, slot
        "place": 4 + 4 + 4,                                   # R key, slot; W index
        "rank_gather": (4 + 4 + 16 + 8 + 4) + (16 + 8 + 8 + 4 + 4),   # R idx, key, pos, vel, uid; W pos, rel, vel, uid, key
        "density": (8 + 4 + 4) + (16 + 4 + 1 + 8 * K),        # R rel, key, uid; W (p, s), pair offset/count, pair records
        "force_integrate": (16 + 8 + 16 + 4 + 1 + 8 * K) + (16 + 8),   # R pos, vel, (p, s), offset/count, records; W pos, vel
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def scene_params(world):
    c = world.coefficients
    return dict(dt=c["dt"], particle_radius=c["particle_radius"], wall_collision_decay=c["wall_collision_decay"],
                pressure_amplifier=c["pressure_amplifier"], ignored_pressure=c["ignored_pressure"],
                collider_noise_level=c["collider_noise_level"], viscosity=c["viscosity"],
                surface_smoothing=c["surface_smoothing"], target_pressure=c["target_pressure"],
                gravity_x=c["gravity"][0], gravity_y=c["gravity"][1])


def coeff_vec(world):
    p = scene_params(world)
    return np.array([p[k] for k in ("dt", "particle_radius", "wall_collision_decay", "pressure_amplifier",
                                    "ignored_pressure", "collider_noise_level", "viscosity", "surface_smoothing",
                                    "target_pressure", "gravity_x", "gravity_y")])


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
            # nvidia-smi takes a few hundred ms to come up and the timed region is short: wait for its first line
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9 or not f[1].isdigit():
                    continue
                sm.append(int(f[1]))
                out["sm_max_mhz"] = int(f[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm = sm[1:] if len(sm) > 2 else sm   # the first line predates the load
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def run_reference_arm(a):
    """The reference's CPU implementation of the path on this box's host cores.  The reference itself is pure
    Python and cannot travel to the GPU box, so this is the oracle port (oracle/step_oracle.c; OpenMP over
    particles for the force loops, scalar neighbor search), on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from sand_crate_b200.scenes import SCENES
    n = a.cpu_particles
    world, pos, vel = SCENES[a.scene](n)
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = coeff_vec(world)
    kin = np.zeros((1, 5))
    tick = 0
    for _ in range(a.warmup):
        out = O.step(cv, pos, vel, seg, [4], kin, noise_mode=1, tkey=O.tick_key(0, tick), want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
        tick += 1
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = O.step(cv, pos, vel, seg, [4], kin, noise_mode=1, tkey=O.tick_key(0, tick), want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
        tick += 1
    dt = time.perf_counter() - t0
    value = n * a.steps / dt
    cores = O.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{a.scene} {a.particles} particles (CPU sample: {n} particles)", "scene": a.scene,
                   "particles": a.particles},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{a.scene} {n} particles x {a.steps} ticks, oracle/step_oracle.c, {cores} OpenMP threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count()},
    }
    emit(line)


def cpu_baseline_sample(scene, budget_s=15.0):
    """cpu_baseline leg: the oracle port on a bounded sample (about budget_s of CPU work)."""
    from oracle import oracle as O
    from sand_crate_b200.scenes import SCENES
    n = 200_000
    world, pos, vel = SCENES[scene](n)
    seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
    cv = coeff_vec(world)
    kin = np.zeros((1, 5))
    out = O.step(cv, pos, vel, seg, [4], kin, noise_mode=1, tkey=O.tick_key(0, 0), want_all=False)  # warm
    pos, vel = out["pos_out"], out["vel_out"]
    t0 = time.perf_counter()
    ticks = 0
    while True:
        out = O.step(cv, pos, vel, seg, [4], kin, noise_mode=1, tkey=O.tick_key(0, 1 + ticks), want_all=False)
        pos, vel = out["pos_out"], out["vel_out"]
        ticks += 1
        el = time.perf_counter() - t0
        if el > budget_s or ticks >= 200:
            break
    cores = O.num_threads()
    return {"value": n * ticks / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{scene} {n} particles x {ticks} ticks in {el:.1f}s, oracle/step_oracle.c, {cores} OpenMP threads "
                      f"(neighbor search scalar).  The reference itself is pure Python: 6.1-7.0e3 particle-steps/s "
                      f"on one core in the build container (BASELINE.md section 2); it cannot run on the GPU box."}


def emit(line: dict) -> None:
    """The ONE JSON line, written to the process's original stdout (see main: libraries such as NCCL print banners
    to fd 1, so everything else is sent to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # anything a library prints to stdout from here on goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="dam_break", choices=["dam_break", "box_fill"])
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--precision", default="mixed", choices=["mixed", "f64"])
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "p2p"], help="strip exchange: NCCL send/recv or "
                    "direct NVLink stores into the neighbor's symmetric-memory buffer")
    ap.add_argument("--rebalance-every", type=int, default=0, help="strips: re-cut the partition every N ticks")
    ap.add_argument("--mgpu-particles", type=int, default=2_000_000, help="particles per GPU when --gpus > 1")
    ap.add_argument("--cpu-particles", type=int, default=200_000)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=20)
    a = ap.parse_args()
    if a.impl == "reference":
        if a.steps == 200 and a.warmup == 100:  # defaults sized for the GPU arm; keep the CPU arm to ~a minute
            a.steps, a.warmup = 20, 3
        a.warmup = max(a.warmup, 1)
        return run_reference_arm(a)
    a.warmup = max(a.warmup, 3)

    import torch
    import torch.distributed as dist
    from sand_crate_b200 import Crate, _lib
    from sand_crate_b200.scenes import SCENES

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a non-default torch stream: the library launches on it and torch.cuda.Event records on it (handle 0, the
    # legacy default stream, would make sc_create open a private stream that torch events cannot see)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    precision = _lib.PRECISION_MIXED if a.precision == "mixed" else _lib.PRECISION_F64
    dom = None
    if world_size == 1:
        n = a.particles
        scene = a.scene
        world, pos, vel = SCENES[scene](n)
        ctx = _lib.Context(n, precision, local_rank, stream)
        ctx.set_params(**scene_params(world))
        seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
        ctx.set_walls(seg, [4], np.zeros((1, 5)))
        ctx.set_noise(_lib.NOISE_COUNTER, 0)
        ctx.set_state(pos, vel)
        step_fn = ctx.step
        n_total = n
    else:
        # one scene cut into horizontal strips of cell rows, NCCL halo + migration exchange every tick
        # (BASELINE.json configs[3]: box-fill, 2M particles per GPU = 16M on 8 GPUs)
        from sand_crate_b200.strips import StripDomain
        scene = "box_fill" if a.scene == "dam_break" and a.particles == 1_000_000 else a.scene
        per_gpu = a.mgpu_particles if a.particles == 1_000_000 else a.particles
        n_total = per_gpu * world_size
        world, pos, vel = SCENES[scene](n_total)
        dom = StripDomain(world, pos, vel, rank=rank, world_size=world_size, precision=a.precision, noise="counter",
                          device=local_rank, stream=stream, transport=a.transport,
                          rebalance_every=a.rebalance_every)
        ctx = dom.ctx
        step_fn = dom.physics_tick
        n = n_total // world_size
        del pos, vel

    flush = None if a.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(a.warmup):
        step_fn()
    barrier()

    def timed_pass(profile):
        """K steps, each bracketed by CUDA events on the launch stream, L2 flushed (untimed) before each."""
        ctx.profile_enable(profile)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        barrier()
        w0 = time.perf_counter()
        for e0, e1 in ev:
            if flush is not None:
                flush.fill_(1)          # untimed: evicts the previous step's lines from the 126 MB L2
            e0.record()
            step_fn()
            e1.record()
        barrier()
        w = time.perf_counter() - w0
        ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
        kern = ctx.profile_read() if profile else {}
        ctx.profile_enable(False)
        return float(ms.sum()), w, kern

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    # pass 1 = THE timed region (value, ms_per_step): plain launches, nothing but the step kernels on the stream
    total_ms, wall, _ = timed_pass(False)
    launches = ctx.launch_count() - launches0
    # pass 2 = the same K steps again with a CUDA-event pair around every kernel launch (per-kernel durations for
    # the roofline); the extra event records stretch the gaps between kernels, so its step time is not the headline
    total_ms_prof, _, kernels = timed_pass(True)
    clocks = sampler.stop()
    n_live = ctx.particle_count() if dom is None else dom.status()["n_local"]
    if dom is None:   # mean directed pairs per particle of the last tick (sizes the pair records in the byte model)
        counts, _ = ctx.get_neighbors(n_live)
        mean_pairs = float(counts.mean())
        del counts
    else:
        mean_pairs = 5.3  # strip mode has no tap; rest-density value measured on one GPU
    dist_status = None if dom is None else dom.status()

    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n_total * a.steps / (total_ms_max * 1e-3)

    # ---- e2e: the public API with host buffers: upload state, tick, read the result back, every step ----------
    if dom is None:
        crate = Crate(world, precision=a.precision, noise="counter", device=local_rank, capacity=n, stream=stream)
        hp = torch.empty((n, 2), dtype=torch.float64, pin_memory=True).numpy()
        hv = torch.empty((n, 2), dtype=torch.float64, pin_memory=True).numpy()
        gp, gv, _ = ctx.get_state(want_pressure=False)
        hp[:], hv[:] = gp, gv

        def e2e_step():
            crate.set_particles(hp, hv)
            crate.physics_tick()
            return crate.particles
        e2e_api = "Crate.set_particles(host, page-locked) -> physics_tick() -> Crate.particles (host, page-locked)"
    else:
        uid0, gp, gv = dom.owned()
        m = len(uid0)
        hp = torch.empty((m, 2), dtype=torch.float64, pin_memory=True).numpy()
        hv = torch.empty((m, 2), dtype=torch.float64, pin_memory=True).numpy()
        hp[:], hv[:] = gp, gv

        def e2e_step():
            dom.ctx.set_state_uids(hp, hv, uid0)
            dom.physics_tick()
            return dom.ctx.dist_get_owned(want_vel=False, want_uid=False, reuse=True)[0]
        e2e_api = ("Context.set_state_uids(host, page-locked) -> StripDomain.physics_tick() -> "
                   "dist_get_owned (positions, host, page-locked), per rank")
    for _ in range(3):
        out_pos = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        out_pos = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * a.e2e_steps / float(t.item())
    assert np.isfinite(out_pos).all()
    h2d = torch.tensor([float(hp.nbytes + hv.nbytes), float(out_pos.nbytes)], device="cuda", dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(h2d)
    h2d_bytes, d2h_bytes = int(h2d[0].item()), int(h2d[1].item())

    if rank == 0:
        peak, peak_src = peaks()
        ALGO_BYTES = algo_bytes(mean_pairs)
        top = max((k for k in kernels if k in ALGO_BYTES), key=lambda k: kernels[k]["ms"], default="density")
        k = kernels.get(top, {"launches": 1, "ms": float("nan")})
        k_ms = k["ms"] / max(k["launches"], 1)
        achieved = SURVEY_BYTES[top] * n_live / (k_ms * 1e-3) / 1e9
        design_achieved = ALGO_BYTES[top] * n_live / (k_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top)
        per_kernel = {}
        for name, v in kernels.items():
            ms = v["ms"] / max(v["launches"], 1)
            per_kernel[name] = {"ms": round(ms, 5), "launches_per_step": round(v["launches"] / a.steps, 2)}
            if name in SURVEY_BYTES:
                per_kernel[name]["algo_gbs"] = round(SURVEY_BYTES[name] * n_live / (ms * 1e-3) / 1e9, 1)
                per_kernel[name]["design_gbs"] = round(ALGO_BYTES[name] * n_live / (ms * 1e-3) / 1e9, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 forces / f64 positions" if a.precision == "mixed" else "f64", "data": "synthetic",
            "config": {"workload": f"{scene} {n_total} particles ({n} per GPU), closed unit box, counter noise 0.1",
                       "scene": scene, "particles_total": n_total, "particles_per_gpu": n,
                       "local_particles_rank0": n_live,
                       "parallelism": "single GPU" if world_size == 1 else
                       f"{world_size} horizontal strips of cell rows, NCCL send/recv halo + migration with rank+-1 "
                       f"every tick (halo {dom.halo_rows} rows, wire buffer {dom.wire_capacity} records, transport {dom.transport}, re-cut every {dom.rebalance_every or 'never'})",
                       "dist_status_rank0": dist_status,
                       "l2": "flushed between timed steps (256 MiB write)" if flush is not None else "not flushed",
                       "timing": "CUDA events per step on the launch stream, summed; max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": a.e2e_steps, "api": e2e_api},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": SURVEY_BYTES[top], "kernel_ms": k_ms,
                         "note": "K4 is issue-bound, not HBM-bound (ncu: 74 % issue-active at 22 of 32 lanes, 6 % DRAM; profiles/r1u_ncu_full_summary.csv)",
                         "design": {"bytes_per_particle": ALGO_BYTES[top], "achieved": design_achieved,
                                    "frac": design_achieved / peak, "mean_pairs_per_particle": mean_pairs},
                         "whole_step": {"survey_bytes_per_particle": 178,
                                        "achieved": 178 * n_live / (total_ms_max / a.steps * 1e-3) / 1e9,
                                        "design_bytes_per_particle": sum(ALGO_BYTES.values()),
                                        "design_achieved": sum(ALGO_BYTES.values()) * n_live / (total_ms_max / a.steps * 1e-3) / 1e9}},
            "kernels": per_kernel,
            "wall_s_timed_region": wall,
            "ms_per_step_with_per_kernel_events": total_ms_prof / a.steps,
        }
        if not a.no_cpu_baseline and world_size == 1:
            line["cpu_baseline"] = cpu_baseline_sample(a.scene)
        emit(line)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
