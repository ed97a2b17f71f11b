"""Host-side particle sources (reference `src/crate/particle_source.py:9-28`).

They stay on the host on purpose: a source emits a handful of particles per tick and draws from the reference's
global NumPy MT19937 stream *between* the step's own draws (SURVEY.md section 8(c) "noise protocol"), so keeping
them here is what makes whole-run parity with the reference possible.  Their output is an input of the GPU step
(`sc_append_particles`)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class ParticleSource:
    radius: float
    position: list
    velocity: list
    flow: float
    active_ticks: int
    noise: float = 0.05

    def generate_particles(self, dt: float, max_particles: int) -> tuple[Optional[np.ndarray], Optional[np.ndarray]]:
        # same draws in the same order as particle_source.py:18-23: binomial, rand(n, 2), rand(n, 2)
        emitted = min(np.round(np.random.binomial(self.flow, dt)), max_particles)
        if emitted == 0:
            return None, None
        jitter = np.random.rand(emitted, 2) - 0.5
        positions = jitter * self.radius + np.array(self.position)
        velocities = np.ones_like(positions) * np.array(self.velocity)[None]
        velocities += (np.random.rand(emitted, 2) - 0.5) * self.noise
        return positions, velocities


def build_particle_sources(particle_source_configs) -> list[ParticleSource]:
    return [ParticleSource(**cfg) for cfg in (particle_source_configs or [])]
