# 2xB200: the re-cut histogram test, then the 64M dam break cut in two with the adaptive re-cut interval (work base 2).
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "recut_histogram or strips_two" 2>&1 | tail -n 5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
wb=${1:-2}
SC_WORK_BASE=$wb timeout 600 $TR --master-port 29541 bench.py --gpus 2 --scene dam_break_wide --particles 32000000 --relax 4000 --warmup 10 --steps 100 --rebalance-every 250 --e2e-steps 1 \
    > gpurun_out/r3g_bench_2gpu_dam64m_adaptive_wb${wb}.json 2> gpurun_out/r3g_bench_2gpu_dam64m_adaptive_wb${wb}.err; echo "wb=$wb rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r3g_bench_2gpu_dam64m_adaptive_wb${wb}.json"))
print("work base $wb:", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e9, 2), "G", [(r["n_local"], round(r["mean_pairs"], 2)) for r in d["strips"]["per_rank"]], d["strips"]["recuts"], d["strips"]["recuts_tick_shift_interval_idleus_tickus"][-4:], d["kernels"].get("io_scatter"))
PY
