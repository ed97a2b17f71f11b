# 8xB200: the 64M dam break with the adaptive re-cut interval (starts at 250) and the block-aggregated histogram
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 75 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --scene dam_break_wide --relax 4000 --warmup 10 --steps 200 --particles 8000000 --rebalance-every 250 --e2e-steps 1 \
    > gpurun_out/r3i_bench_8gpu_dam64m_adaptive.json 2> gpurun_out/r3i_bench_8gpu_dam64m_adaptive.err; echo "64M adaptive rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r3i_bench_8gpu_dam64m_adaptive.json"))
print(round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e9, 2), "G", [(r["n_local"], round(r["mean_pairs"], 2)) for r in d["strips"]["per_rank"]], d["strips"]["recuts"], d["strips"]["recuts_tick_shift_interval_idleus_tickus"][-5:])
PY
