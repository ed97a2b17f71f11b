# 8xB200: the 64M dam break with the calibrated work model of the re-cutter (2 + pairs), then box-fill 16M again
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --scene dam_break_wide --relax 4000 --warmup 10 --steps 200 --particles 8000000 --rebalance-every 250 \
    > gpurun_out/r3e_bench_8gpu_dam64m_recut250.json 2> gpurun_out/r3e_bench_8gpu_dam64m_recut250.err; echo "64M recut rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 200 --warmup 10 \
    > gpurun_out/r3e_bench_8gpu_boxfill16m.json 2> gpurun_out/r3e_bench_8gpu_boxfill16m.err; echo "16M box rc=$?"
