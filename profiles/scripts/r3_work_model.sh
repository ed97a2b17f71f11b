# 2xB200: the 64M dam break cut in two (32M per GPU) under different work models of the re-cutter:
# weight of a particle = SC_WORK_BASE + its pair count.  13 = the first guess; smaller = the pairs weigh more.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in "0 0" "2 8" "0 5"; do
  set -- $cfg; wb=$1; wq=$2
  SC_WORK_BASE=$wb SC_WORK_QUAD=$wq timeout 600 $TR --master-port 2952$wq bench.py --gpus 2 --scene dam_break_wide --particles 32000000 --relax 4000 --warmup 10 --steps 100 --rebalance-every 250 --e2e-steps 1 \
    > gpurun_out/r3d_bench_2gpu_dam64m_wb${wb}_q${wq}.json 2> gpurun_out/r3d_bench_2gpu_dam64m_wb${wb}_q${wq}.err; echo "wb=$wb wq=$wq rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r3d_bench_2gpu_dam64m_wb${wb}_q${wq}.json"))
print("work base $wb quad $wq:", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e9, 2), "G", [(r["n_local"], round(r["mean_pairs"], 2)) for r in d["strips"]["per_rank"]])
PY
done
