"""TEST INFRASTRUCTURE ONLY - ctypes wrapper around ``oracle/step_oracle.c`` (the CPU restatement of
`Crate.physics_tick`, reference src/crate/crate.py:91-129).  See the header of ``step_oracle.c`` for the parity
status.  The product package never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "step_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")
MAX_NEIGHBORS = 20

COEFF_FIELDS = ("dt", "radius", "wall_collision_decay", "pressure_amplifier", "ignored_pressure",
                "collider_noise_level", "viscosity", "surface_smoothing", "target_pressure", "gx", "gy")


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in COEFF_FIELDS]

    @classmethod
    def from_array(cls, a):
        a = np.asarray(a, dtype=np.float64)
        assert a.shape == (len(COEFF_FIELDS),)
        return cls(*[float(v) for v in a])


def build(force: bool = False) -> str:
    """gcc recipe for the restatement.  -ffp-contract=off: NumPy never fuses multiply-add."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
               "-o", _LIB, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.oc_detect_particle_collisions.argtypes = [dp, C.c_int64, C.c_double, lp, lp, ip, ip]
        L.oc_detect_particle_collisions.restype = C.c_int
        L.oc_step.argtypes = [C.POINTER(Params), C.c_int64, dp, dp, dp, C.c_int, ip, dp, C.c_int, C.c_int, dp,
                              C.c_uint64, C.POINTER(C.c_uint32), dp, ip, ip, dp, dp, dp, ip, dp]
        L.oc_step.restype = C.c_int
        L.oc_remove_particles.argtypes = [dp, dp, C.c_int64, C.c_double, C.POINTER(C.c_uint8)]
        L.oc_remove_particles.restype = C.c_int64
        L.oc_points_to_segments_distance.argtypes = [dp, C.c_int64, dp, C.c_int, dp, dp]
        L.oc_points_to_segments_distance.restype = None
        L.oc_pad_segments.argtypes = [dp, C.c_int, C.c_double, dp]
        L.oc_pad_segments.restype = None
        L.oc_tick_key.argtypes = [C.c_uint64, C.c_uint64]
        L.oc_tick_key.restype = C.c_uint64
        L.oc_pair_noise.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, dp, dp]
        L.oc_pair_noise.restype = None
        L.oc_num_threads.restype = C.c_int
        L.oc_source_key.argtypes = [C.c_uint64, C.c_uint32]
        L.oc_source_key.restype = C.c_uint64
        L.oc_source_uniform.argtypes = [C.c_uint64, C.c_uint64]
        L.oc_source_uniform.restype = C.c_double
        L.oc_emit_counter.argtypes = [C.c_uint64, C.c_int, dp, ip, C.POINTER(C.c_uint32), C.c_int64, C.c_int64, dp, dp]
        L.oc_emit_counter.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64)) if a is not None else None


def detect_particle_collisions(particles, diameter):
    """Restatement of collision_detector.py:9-49.  Returns (rows_sorted, order, counts, idx[P, 20])."""
    pos = np.ascontiguousarray(particles, dtype=np.float64).reshape(-1, 2)
    P = pos.shape[0]
    rows = np.zeros(P, np.int64)
    order = np.zeros(P, np.int64)
    counts = np.zeros(P, np.int32)
    idx = np.full((P, MAX_NEIGHBORS), -1, np.int32)
    rc = lib().oc_detect_particle_collisions(_dp(pos), P, float(diameter), _lp(rows), _lp(order), _ip(counts), _ip(idx))
    if rc:
        raise MemoryError("oracle allocation failed")
    return rows, order, counts, idx


def neighbor_lists(counts, idx):
    return [list(map(int, idx[i, :counts[i]])) for i in range(len(counts))]


def points_to_segments_distance(p, segments):
    """Restatement of geometry_utils.py:7-39.  Returns (nearest[P, S, 2], distance[P, S])."""
    pos = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 2)
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
    P, S = pos.shape[0], seg.shape[0]
    nearest = np.zeros((P, S, 2))
    dist = np.zeros((P, S))
    lib().oc_points_to_segments_distance(_dp(pos), P, _dp(seg), S, _dp(nearest), _dp(dist))
    return nearest, dist


def pad_segments(segments, pad):
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((2 * seg.shape[0], 2, 2))
    lib().oc_pad_segments(_dp(seg), seg.shape[0], float(pad), _dp(out))
    return out


def remove_particles(pos, vel, radius):
    pos = np.array(pos, dtype=np.float64).reshape(-1, 2)
    vel = np.array(vel, dtype=np.float64).reshape(-1, 2)
    mask = np.zeros(pos.shape[0], np.uint8)
    n = lib().oc_remove_particles(_dp(pos), _dp(vel), pos.shape[0], float(radius),
                                  mask.ctypes.data_as(C.POINTER(C.c_uint8)))
    return pos[:n].copy(), vel[:n].copy(), mask.astype(bool)


def tick_key(seed: int, tick: int) -> int:
    return int(lib().oc_tick_key(C.c_uint64(seed), C.c_uint64(tick)))


def source_uniform(seed: int, tick: int, source_index: int, j: int) -> float:
    L = lib()
    return float(L.oc_source_uniform(L.oc_source_key(C.c_uint64(tick_key(seed, tick)), C.c_uint32(source_index)),
                                     C.c_uint64(j)))


def source_count(u: float, flow: float, dt: float) -> int:
    """Binomial(flow, dt) through its inverse CDF at u (the emission count of particle_source.py:18 before the clamp):
    the smallest k whose cumulative probability reaches u."""
    trials, p = int(flow), float(dt)
    if trials <= 0 or p <= 0:
        return 0
    if p >= 1:
        return trials
    import math
    q = 1.0 - p
    mass = math.pow(q, trials)
    cdf, k = mass, 0
    while u > cdf and k < trials:
        k += 1
        mass *= (trials - k + 1) / k * (p / q)
        cdf += mass
    return k


def emit_counter(seed: int, tick: int, sources, dt: float, particle_count: int, max_particles: int):
    """The counter-stream version of create_new_particles for one tick.  sources: the world's particle_sources dicts
    (radius, position, velocity, flow, active_ticks, noise).  Returns (pos, vel) of the appended rows."""
    rows, counts, index = [], [], []
    for q, s in enumerate(sources):
        if s["active_ticks"] <= tick:           # crate.py:140
            continue
        n = source_count(source_uniform(seed, tick, q, 0), s["flow"], dt)
        if n == 0:
            continue
        rows.append([s["position"][0], s["position"][1], s["radius"], s["velocity"][0], s["velocity"][1],
                     s.get("noise", 0.05)])
        counts.append(n)
        index.append(q)
    if not rows:
        return np.zeros((0, 2)), np.zeros((0, 2))
    src = np.ascontiguousarray(rows, dtype=np.float64)
    cnt = np.ascontiguousarray(counts, dtype=np.int32)
    idx = np.ascontiguousarray(index, dtype=np.uint32)
    pos = np.zeros((int(cnt.sum()), 2))
    vel = np.zeros((int(cnt.sum()), 2))
    m = lib().oc_emit_counter(C.c_uint64(tick_key(seed, tick)), len(rows), _dp(src), _ip(cnt),
                              idx.ctypes.data_as(C.POINTER(C.c_uint32)), int(particle_count), int(max_particles),
                              _dp(pos), _dp(vel))
    return pos[:m].copy(), vel[:m].copy()


def pair_noise(tkey: int, uid_i: int, uid_j: int):
    ux, uy = C.c_double(), C.c_double()
    lib().oc_pair_noise(C.c_uint64(tkey), C.c_uint32(uid_i), C.c_uint32(uid_j), C.byref(ux), C.byref(uy))
    return ux.value, uy.value


def step(coeffs, pos, vel, segments, body_len, body_kin, noise_mode=0, noise=None, tkey=0, uid=None,
         want_all=True):
    """One step of the restatement on inputs taken after create/remove/apply_bodies_velocity.

    coeffs: the 11-vector of ``COEFF_FIELDS``.  Returns a dict with pos_out, vel_out and the intermediates."""
    pos = np.array(pos, dtype=np.float64).reshape(-1, 2)
    vel = np.array(vel, dtype=np.float64).reshape(-1, 2)
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
    body_len = np.ascontiguousarray(body_len, dtype=np.int32)
    body_kin = np.ascontiguousarray(body_kin, dtype=np.float64).reshape(-1, 5)
    assert int(body_len.sum()) == seg.shape[0]
    P = pos.shape[0]
    prm = Params.from_array(coeffs)
    out = {}
    if want_all:
        out["pos_search"] = np.zeros((P, 2))
        out["nbr_count"] = np.zeros(P, np.int32)
        out["nbr_idx_padded"] = np.full((P, MAX_NEIGHBORS), -1, np.int32)
        out["pressure"] = np.zeros(P)
        out["tension_vec"] = np.zeros((P, 2))
        out["ccd_factor"] = np.ones(P)
        out["wall_count"] = np.zeros(P, np.int32)
        out["force_monitor"] = np.zeros(6)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
    if uid is not None:
        uid = np.ascontiguousarray(uid, dtype=np.uint32)
    rc = lib().oc_step(C.byref(prm), P, _dp(pos), _dp(vel), _dp(seg), seg.shape[0], _ip(body_len), _dp(body_kin),
                       body_len.shape[0], int(noise_mode), _dp(noise), C.c_uint64(tkey),
                       uid.ctypes.data_as(C.POINTER(C.c_uint32)) if uid is not None else None,
                       _dp(out.get("pos_search")), _ip(out.get("nbr_count")), _ip(out.get("nbr_idx_padded")),
                       _dp(out.get("pressure")), _dp(out.get("tension_vec")), _dp(out.get("ccd_factor")),
                       _ip(out.get("wall_count")), _dp(out.get("force_monitor")))
    if rc:
        raise MemoryError("oracle allocation failed")
    out["pos_out"] = pos
    out["vel_out"] = vel
    if want_all:
        c = out["nbr_count"]
        out["nbr_idx"] = (np.concatenate([out["nbr_idx_padded"][i, :c[i]] for i in range(P)]).astype(np.int32)
                          if P and c.sum() else np.zeros(0, np.int32))
    return out


def num_threads() -> int:
    return int(lib().oc_num_threads())
