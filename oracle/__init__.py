"""TEST INFRASTRUCTURE ONLY - the parity oracle for the SandCrate particle step.

``oracle.step_oracle.c``  CPU restatement of the reference step (pinned bit-for-bit against goldens recorded
                          from the unmodified reference; see its header).
``oracle.oracle``         ctypes wrapper + build recipe for the restatement.
``oracle.ref_shim``       loader for the unmodified reference (build container only).
``oracle.make_golden``    generates ``tests/golden/*.npz`` from the unmodified reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  ``sand_crate_b200`` never does.
"""
