"""TEST INFRASTRUCTURE ONLY - the reference's own sensitivity to a 1e-13 perturbation, as the noise floor for the
free-run drift bounds.

    python tests/make_spread.py            # writes tests/golden/spread_<config>.json (a minute or two)

The reference trajectory is chaotic on a ~50-tick horizon, so a production-mode (fp32 forces) run can only be compared
with it on aggregates.  How far apart may aggregates be?  This script answers with the reference itself: the step is
run through the CPU oracle (bit-identical to the reference, tests/test_oracle_golden.py) under the drop-in `Crate`
host protocol, once unperturbed and several times with positions shifted by 1e-13 * N(0, 1) at tick 5 (as early as there
are particles: a reduced-precision run departs from the reference in its first tick, so the perturbed runs must be given
the same time to decorrelate - a perturbation at tick 100 leaves them only ~50 ticks of macroscopic divergence at the
tick-200 checkpoint), and the aggregates of tests/conftest.py::aggregates are recorded at the free-run checkpoints.  The GPU test
(tests/test_gpu_parity.py::test_mixed_free_run_vs_reference_aggregates) allows the mixed-precision run the SURVEY
bound or twice this spread, whichever is larger."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from conftest import aggregates, world_from_freerun  # noqa: E402
from oracle_backend import OracleContext  # noqa: E402
from sand_crate_b200 import Crate, crate as crate_mod  # noqa: E402

CHECKPOINTS = {"stirring_cup": [200, 400, 800, 1200], "wave_machine": [500, 1000, 2000, 3000]}
PERTURB_AT, PERTURBED_RUNS = 5, 8


def run(name, seed):
    world, g = world_from_freerun(name)
    crate = Crate(world)
    out = {}
    for tick in range(1, max(CHECKPOINTS[name]) + 1):
        crate.physics_tick()
        if tick == PERTURB_AT and seed is not None:
            rs = np.random.RandomState(seed)   # a private generator: the global stream belongs to the simulation
            crate._ctx.pos = crate._ctx.pos + 1e-13 * rs.randn(*crate._ctx.pos.shape)
            crate._cache = {}
        if tick in CHECKPOINTS[name]:
            out[str(tick)] = aggregates(crate.particles, crate.particle_velocities, crate.particles_pressure)
            if seed is None:  # the unperturbed run IS the recorded reference trajectory
                assert np.array_equal(crate.particles, g[f"pos_t{tick}"]), (name, tick)
    return out


def main():
    crate_mod._lib.Context = OracleContext
    for name in CHECKPOINTS:
        base = run(name, None)
        runs = [run(name, 1000 + k) for k in range(PERTURBED_RUNS)]
        spread = {t: {k: max(abs(r[t][k] - base[t][k]) for r in runs) for k in base[t]} for t in base}
        path = os.path.join(HERE, "golden", f"spread_{name}.json")
        json.dump({"perturbation": f"1e-13 * N(0,1) on positions at tick {PERTURB_AT}, {PERTURBED_RUNS} runs",
                   "reference": base, "max_abs_spread": spread}, open(path, "w"), indent=1)
        print(f"wrote {path}")
        for t in base:
            print(f"  {name} t={t}: " + ", ".join(f"{k} {base[t][k]:.4g} +- {spread[t][k]:.2g}" for k in base[t]))


if __name__ == "__main__":
    main()
