# 1xB200: the bench line, the same scene size as one strip of the multi-GPU runs, then the ncu launch list (durations only)
# and the full-set capture of one tick's six kernels.  Each ncu run follows a plain run of the same command.
set -x
python bench.py --steps 200 --warmup 10 > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; echo "bench rc=$?"
python bench.py --steps 200 --warmup 10 --scene box_fill --particles 2000000 --no-cpu-baseline > gpurun_out/r2m_bench_1gpu_boxfill2m.json 2> gpurun_out/r2m_bench_1gpu_boxfill2m.err; echo "bench 2M rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/r2m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 618 -c 36 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/r2m_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_prepass|k_scan_lookback|k_place|k_rank_gather|k_density_tile|k_force_tile" -s 618 -c 12 -o gpurun_out/prof_r2m $CMD > gpurun_out/r2m_ncu_full.log 2>&1; echo "ncu full rc=$?"
