# 2xB200: the 64M dam break cut in two, adaptive re-cut interval (starts at 250), work base 2 against 13.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for wb in 2 13; do
  SC_WORK_BASE=$wb timeout 600 $TR --master-port $((29530 + wb)) bench.py --gpus 2 --scene dam_break_wide --particles 32000000 --relax 4000 --warmup 10 --steps 100 --rebalance-every 250 --e2e-steps 1 \
    > gpurun_out/r3f_bench_2gpu_dam64m_adaptive_wb${wb}.json 2> gpurun_out/r3f_bench_2gpu_dam64m_adaptive_wb${wb}.err; echo "wb=$wb rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r3f_bench_2gpu_dam64m_adaptive_wb${wb}.json"))
print("work base $wb:", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e9, 2), "G", [(r["n_local"], round(r["mean_pairs"], 2)) for r in d["strips"]["per_rank"]], d["strips"]["recuts"], d["strips"]["recuts_tick_shift_interval_idleus_tickus"])
PY
done
