import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def step_goldens():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "step_*.npz")))


def neighbor_case_tags():
    g = golden("neighbors_cases.npz")
    return sorted({k.split("__")[0] for k in g.files})


def world_from_freerun(name):
    from sand_crate_b200 import WorldConfig
    g = golden(f"freerun_{name}.npz")
    w = json.loads(str(g["world_json"]))
    return WorldConfig(rigid_bodies=w["rigid_bodies"], particle_sources=w["particle_sources"],
                       coefficients=w["coefficients"]), g


def params_from_coeffs(c):
    """golden `coeffs` 11-vector (oracle.ref_shim.coefficients_of) -> sc_params keyword dict."""
    names = ("dt", "particle_radius", "wall_collision_decay", "pressure_amplifier", "ignored_pressure",
             "collider_noise_level", "viscosity", "surface_smoothing", "target_pressure", "gravity_x", "gravity_y")
    return {n: float(v) for n, v in zip(names, c)}


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def aggregates(pos, vel, prs):
    """The bulk quantities the free-run drift bounds are stated on ("1000-step free run")."""
    pos, vel, prs = np.asarray(pos), np.asarray(vel), np.asarray(prs)
    return {"count": int(len(pos)), "com_x": float(pos[:, 0].mean()), "com_y": float(pos[:, 1].mean()),
            "kinetic": float(0.5 * (vel ** 2).sum()), "mean_speed": float(np.sqrt((vel ** 2).sum(1)).mean()),
            "p_mean": float(prs.mean()), "p_max": float(prs.max())}
