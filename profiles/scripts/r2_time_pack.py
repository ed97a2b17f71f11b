"""Diagnostic: how long does the strip pack kernel take by itself?  One GPU plays rank 0 of 2 over the lower... upper half
of a 4M box-fill (its own buffers stand in for the neighbor's), steps a few ticks and times sc_dist_pack /
sc_dist_pack_push with CUDA events, with and without a valid search state (zone-restricted vs full pass)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sand_crate_b200 import _lib  # noqa: E402
from sand_crate_b200.scenes import box_fill  # noqa: E402
from sand_crate_b200.strips import partition_rows, rows_of  # noqa: E402

world, pos, vel = box_fill(4_000_000)
c = world.coefficients
d = 2 * c["particle_radius"]
rows = rows_of(pos, d)
cuts = partition_rows(rows, 2)
mine = np.nonzero(rows < cuts[1])[0]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = _lib.Context(int(len(mine) * 1.3) + 300_000, _lib.PRECISION_MIXED, 0, stream.cuda_stream)
ctx.set_params(dt=c["dt"], particle_radius=c["particle_radius"], wall_collision_decay=c["wall_collision_decay"],
               pressure_amplifier=c["pressure_amplifier"], ignored_pressure=c["ignored_pressure"],
               collider_noise_level=c["collider_noise_level"], viscosity=c["viscosity"],
               surface_smoothing=c["surface_smoothing"], target_pressure=c["target_pressure"],
               gravity_x=c["gravity"][0], gravity_y=c["gravity"][1])
seg = np.array(world.rigid_bodies[0]["fixed"]["segments"], dtype=np.float64)
ctx.set_walls(seg, [4], np.zeros((1, 5)))
ctx.set_noise(_lib.NOISE_COUNTER, 0)
ctx.set_state_uids(pos[mine], vel[mine], mine.astype(np.uint32))
wire = 96_000
ctx.dist_configure(0, 2, cuts[0], cuts[1], 4, wire)
nbytes = _lib.wire_bytes(wire)
send = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
recv = torch.zeros(nbytes + 256, dtype=torch.uint8, device="cuda")
flag = torch.zeros(64, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    out = []
    for _ in range(reps):
        ctx.step()                      # a tick: valid search state, arrays in sorted order
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3)
        # throw the packed records away: re-arm the buffer by a zero-size unpack substitute
        send.zero_()
    return np.median(out), np.min(out)


tick = [0]


def pack_local():
    ctx.dist_pack(None, send)


def pack_direct():
    tick[0] += 1
    ctx.dist_pack_push(None, (send, recv.data_ptr(), flag.data_ptr()), tick[0])


print("n_local", ctx.dist_status()["n_local"])
print("pack (local buffer, zone-restricted): median %.1f us, min %.1f us" % timed(pack_local))
print("pack+push (direct to a device buffer, zone-restricted): median %.1f us, min %.1f us" % timed(pack_direct))
st = ctx.dist_status(None, send)
print("flags", st)
