# 1xB200, final build of the round: GPU suite (plain, then with every allocation poisoned), the reference arm (C port + the
# unmodified NumPy step from baseline/_ref), the bench line, the ncu launch list and the full-set capture of one tick.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log; tail -2 gpurun_out/r2s_tests.log
SC_POISON=0xFF python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests_poisoned.log 2>&1; echo "poisoned tests rc=$?" >> gpurun_out/r2s_tests_poisoned.log; tail -2 gpurun_out/r2s_tests_poisoned.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2s_bench_reference_arm.json 2> gpurun_out/r2s_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --steps 200 --warmup 10 > gpurun_out/r2s_bench_1gpu_final.json 2> gpurun_out/r2s_bench_1gpu_final.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_1gpu_driver_args.json 2> gpurun_out/r2s_bench_1gpu_driver_args.err; echo "bench (driver args) rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/r2s_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 618 -c 36 --csv --log-file gpurun_out/r2s_launches.csv $CMD > gpurun_out/r2s_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/r2s_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_prepass|k_scan_lookback|k_place|k_rank_gather|k_density_tile|k_force_tile" -s 618 -c 12 -o gpurun_out/prof_r2s $CMD > gpurun_out/r2s_ncu_full.log 2>&1; echo "ncu full rc=$?"
