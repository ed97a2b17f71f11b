"""Test double: a `Context` look-alike that runs the step through the CPU oracle (oracle/step_oracle.c).

It exists so the HOST logic of `sand_crate_b200.Crate` (RNG protocol, particle sources, rigid-body motion, lazy
counts, capacity growth) can be tested on a machine without a GPU.  It lives under tests/ because only tests may
touch oracle/; the product never sees it."""
import numpy as np

from oracle import oracle as O


class OracleContext:
    def __init__(self, capacity, precision=0, device=0, stream=None):
        self.capacity, self.precision, self.device = int(capacity), precision, device
        self.pos = np.zeros((0, 2))
        self.vel = np.zeros((0, 2))
        self.prs = np.zeros(0)
        self.uid = np.zeros(0, np.uint32)
        self.next_uid = 0
        self.params = None
        self.walls = None
        self.noise_mode, self.seed, self.tick = 0, 0, 0
        self._pending = None

    def close(self):
        pass

    def set_params(self, **kw):
        self.params = kw

    def set_walls(self, segments, body_len, body_kin):
        self.walls = (np.array(segments, dtype=np.float64).reshape(-1, 4), np.array(body_len, np.int32),
                      np.array(body_kin, dtype=np.float64).reshape(-1, 5))

    def set_noise(self, mode, seed=0):
        self.noise_mode, self.seed = mode, seed

    def set_tick(self, tick):
        self.tick = tick

    def profile_enable(self, on=True):
        pass

    def profile_read(self):
        return {}

    def set_state(self, pos, vel):
        self.pos, self.vel = np.array(pos, dtype=np.float64), np.array(vel, dtype=np.float64)
        self.uid = np.arange(len(self.pos), dtype=np.uint32)
        self.next_uid = len(self.pos)
        self.prs = np.zeros(len(self.pos))

    def append_particles(self, pos, vel):
        assert len(self.pos) + len(pos) <= self.capacity
        self.pos = np.vstack((self.pos, pos))
        self.vel = np.vstack((self.vel, vel))
        self.uid = np.concatenate((self.uid, np.arange(self.next_uid, self.next_uid + len(pos), dtype=np.uint32)))
        self.next_uid += len(pos)
        self.prs = np.zeros(len(self.pos))

    def particle_count(self):
        return len(self.pos)

    def get_state(self, want_vel=True, want_pressure=True):
        return self.pos.copy(), self.vel.copy(), self.prs.copy()

    def _coeffs(self):
        p = self.params
        return np.array([p[k] for k in ("dt", "particle_radius", "wall_collision_decay", "pressure_amplifier",
                                        "ignored_pressure", "collider_noise_level", "viscosity",
                                        "surface_smoothing", "target_pressure", "gravity_x", "gravity_y")])

    def _remove(self):
        pos, vel, mask = O.remove_particles(self.pos, self.vel, self.params["particle_radius"])
        self.uid = self.uid[~mask]
        self.pos, self.vel = pos, vel

    def step_begin(self):
        self._remove()
        seg, bl, bk = self.walls
        probe = O.step(self._coeffs(), self.pos, self.vel, seg, bl, bk, noise_mode=0)
        self._pending = True
        return len(self.pos), int(probe["nbr_count"].sum())

    def step_finish(self, noise=None):
        seg, bl, bk = self.walls
        mode = self.noise_mode if self.params["collider_noise_level"] != 0 else 0
        out = O.step(self._coeffs(), self.pos, self.vel, seg, bl, bk, noise_mode=mode, noise=noise,
                     tkey=O.tick_key(self.seed, self.tick), uid=self.uid)
        self.pos, self.vel, self.prs = out["pos_out"], out["vel_out"], out["pressure"]
        self.tick += 1
        self._pending = None

    def step(self, n=1):
        for _ in range(n):
            self.step_begin()
            self.step_finish(None)
