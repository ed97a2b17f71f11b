"""TEST INFRASTRUCTURE ONLY - loader for the *unmodified* reference step.

This module imports David-Taub/sand_crate's ``Crate`` from ``/root/reference`` (read-only, present in the
build container only, never on the GPU box) so that golden vectors can be generated from the real reference
(``oracle/make_golden.py``) and the CPU restatement in ``oracle/step_oracle.c`` can be pinned against it.
Nothing in the product package (``sand_crate_b200/``) imports it.

The reference does not import as shipped on this image:
  * ``nptyping`` is absent (used only as the alias ``NDArray``)           -> stub module;
  * ``pygame`` is absent (used only for ``Vector2(x, y).rotate(deg)`` in
    ``src/crate/rigid_body.py:38-39``)                                   -> stub module (``_Vector2`` below);
  * ``src/crate/rigid_body.py:22`` has a mutable dataclass default that
    Python >= 3.11 rejects                                               -> one-line in-memory patch at load time.
No reference source is copied into this repository; the files are executed where they lie.
"""
from __future__ import annotations

import importlib
import importlib.util
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SANDCRATE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "crate", "crate.py"))


def rotate_like_pygame(x: float, y: float, angle_deg: float) -> tuple[float, float]:
    """Restatement of pygame 2.x ``Vector2.rotate`` (math.c, ``_vector2_rotate_helper``): angle reduced to
    [0, 360), exact quarter turns special-cased, otherwise ``cos*x - sin*y, sin*x + cos*y`` with the angle
    converted as ``deg * pi / 180``.  pygame itself is not installed here, so the last ulp of placed wall
    endpoints is *defined* by this function on both the oracle and the product side; placed segments are
    recorded as oracle inputs, so it cannot affect step parity."""
    eps = 1e-6
    angle = math.fmod(angle_deg, 360.0)
    if angle < 0:
        angle += 360.0
    if math.fmod(angle + eps, 90.0) < 2 * eps:
        quarter = int((angle + eps) / 90.0)
        if quarter in (0, 4):
            return (x, y)
        if quarter == 1:
            return (-y, x)
        if quarter == 2:
            return (-x, -y)
        return (y, -x)
    rad = angle * math.pi / 180.0
    s, c = math.sin(rad), math.cos(rad)
    return (c * x - s * y, s * x + c * y)


class _Vector2:
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = float(x), float(y)

    def rotate(self, angle_deg):
        return rotate_like_pygame(self.x, self.y, angle_deg)


def _install_stubs() -> None:
    if "nptyping" not in sys.modules:
        m = types.ModuleType("nptyping")
        m.NDArray = np.ndarray
        sys.modules["nptyping"] = m
    if "pygame" not in sys.modules:
        m = types.ModuleType("pygame")
        m.Vector2 = _Vector2
        sys.modules["pygame"] = m


_loaded = None


def load_reference():
    """Returns the namespace ``(Crate, load_config, collision_detector, geometry_utils)`` of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT} (it only exists in the build container)")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # packages first (empty __init__ files), then the patched rigid_body, then everything else unmodified
    importlib.import_module("src")
    importlib.import_module("src.crate")
    rb_path = os.path.join(REFERENCE_ROOT, "src", "crate", "rigid_body.py")
    with open(rb_path, "r") as f:
        src = f.read()
    bad = "center_velocity: NDArray = np.array([0.0, 0.0])"
    assert bad in src, "reference rigid_body.py changed; shim needs review"
    src = src.replace(bad, "center_velocity: NDArray = field(default_factory=lambda: np.array([0.0, 0.0]))")
    spec = importlib.util.spec_from_file_location("src.crate.rigid_body", rb_path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["src.crate.rigid_body"] = mod
    exec(compile(src, rb_path, "exec"), mod.__dict__)
    crate_mod = importlib.import_module("src.crate.crate")
    cfg_mod = importlib.import_module("src.crate.load_config")
    cd_mod = importlib.import_module("src.crate.collision_detector")
    geo_mod = importlib.import_module("src.crate.utils.geometry_utils")
    _loaded = types.SimpleNamespace(
        Crate=crate_mod.Crate,
        load_config=cfg_mod.load_config,
        collision_detector=cd_mod,
        geometry_utils=geo_mod,
        crate_module=crate_mod,
        config_dir=os.path.join(REFERENCE_ROOT, "config"),
    )
    return _loaded


class RecordingCrate:
    """Wraps a reference ``Crate`` and records, per tick, everything the GPU step takes as input and produces as
    output.  Instance-level method wrapping only; reference code is untouched."""

    def __init__(self, world_config):
        ref = load_reference()
        self.ref = ref
        self.crate = ref.Crate(world_config)
        self.record = None
        c = self.crate
        orig_cvc = c.calc_virtual_colliders
        orig_pop = c.populate_colliders
        orig_cpp = c.compute_particle_pressures
        orig_tension = c.apply_tension

        def calc_virtual_colliders():
            if self.record is not None:
                r = self.record
                r["pos_in"] = c.particles.copy()
                r["vel_in"] = c.particle_velocities.copy()
                r["segments"] = c.segments.copy()
                r["body_len"] = np.array([len(b) for b in c.rigid_bodies], dtype=np.int32)
                r["body_kin"] = np.array(
                    [[b.center_velocity[0], b.center_velocity[1], b.angular_clockwise_velocity,
                      b.position[0], b.position[1]] for b in c.rigid_bodies], dtype=np.float64).reshape(-1, 5)
            orig_cvc()

        def populate_colliders():
            if self.record is None:
                return orig_pop()
            r = self.record
            r["pos_search"] = c.particles.copy()  # after apply_hard_wall_fix
            draws = []
            real_rand = np.random.rand

            def rec_rand(*shape):
                out = real_rand(*shape)
                draws.append(out.reshape(-1, 2).copy())
                return out

            np.random.rand = rec_rand
            try:
                orig_pop()
            finally:
                np.random.rand = real_rand
            r["noise"] = np.concatenate(draws, 0) if draws else np.zeros((0, 2))
            counts = np.array([len(n) for n in c.colliders_indices], dtype=np.int32)
            r["nbr_count"] = counts
            r["nbr_idx"] = (np.concatenate([np.asarray(n, dtype=np.int32) for n in c.colliders_indices])
                            if counts.sum() else np.zeros(0, np.int32))

        def compute_particle_pressures():
            orig_cpp()
            if self.record is not None:
                self.record["pressure"] = np.asarray(c.particles_pressure, dtype=np.float64).copy()

        def apply_tension():
            if self.record is not None:
                # the surface normals are a local of apply_tension (crate.py:337-342); restate the same
                # expression here from the reference's own per-particle arrays so the intermediate is pinned
                s = np.zeros((c.particle_count, 2))
                for i in range(c.particle_count):
                    if c.colliders_count(i) == 0:
                        continue
                    s[i] = np.sum(((1 - c.collider_overlaps[i]) * c.collider_overlaps[i])[:, None] * c.colliders[i], 0)
                self.record["tension_vec"] = s
            orig_tension()

        c.calc_virtual_colliders = calc_virtual_colliders
        c.populate_colliders = populate_colliders
        c.compute_particle_pressures = compute_particle_pressures
        c.apply_tension = apply_tension

    def tick(self, record: bool = False):
        c = self.crate
        self.record = {} if record else None
        if record:
            self.record["coeffs"] = coefficients_of(c)
        c.physics_tick()
        rec = self.record
        self.record = None
        if rec is not None:
            rec["pos_out"] = c.particles.copy()
            rec["vel_out"] = c.particle_velocities.copy()
            rec["tick_after"] = np.int64(c.tick)
            d = c.diameter
            cd = self.ref.collision_detector
            if len(rec["pos_search"]):
                _, yf, order = cd.strip_sort_particles(particles=rec["pos_search"], diameter=d)
                rec["rows_sorted"] = np.asarray(yf, dtype=np.int64)
                rec["order"] = np.asarray(order, dtype=np.int64)
            else:
                rec["rows_sorted"] = np.zeros(0, np.int64)
                rec["order"] = np.zeros(0, np.int64)
        return rec


COEFF_NAMES = ("dt", "particle_radius", "wall_collision_decay", "pressure_amplifier", "ignored_pressure",
               "collider_noise_level", "viscosity", "surface_smoothing", "target_pressure")


def coefficients_of(crate) -> np.ndarray:
    """[dt, r, decay, amp, ignored, noise_level, visc, smoothing, target, gx, gy] as float64."""
    vals = [float(getattr(crate, n)) for n in COEFF_NAMES]
    vals += [float(crate.gravity[0]), float(crate.gravity[1])]
    return np.array(vals, dtype=np.float64)
