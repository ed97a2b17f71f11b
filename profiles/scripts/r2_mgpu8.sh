# 8xB200: BASELINE.json configs[4] (dam-break 64M) with / without re-cutting, configs[3] (box-fill 16M), then - side by
# side on disjoint GPUs - a 4-GPU strong-scaling point of the 64M scene and a 2-GPU run whose exchange kernels are
# launched without programmatic serialization (so their event times are their own).  One gpurun --gpus 8 call.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
W="--scene dam_break_wide --relax 4000 --warmup 10 --steps 200"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 $W --particles 8000000 --rebalance-every 250 \
    > gpurun_out/r2k_bench_8gpu_dam64m_recut250.json 2> gpurun_out/r2k_bench_8gpu_dam64m_recut250.err; echo "64M recut rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 $W --particles 8000000 --rebalance-every 0 --no-weak-baseline \
    > gpurun_out/r2k_bench_8gpu_dam64m_norecut.json 2> gpurun_out/r2k_bench_8gpu_dam64m_norecut.err; echo "64M no recut rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 200 --warmup 10 \
    > gpurun_out/r2k_bench_8gpu_boxfill16m.json 2> gpurun_out/r2k_bench_8gpu_boxfill16m.err; echo "16M box rc=$?"
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 $W --particles 16000000 --rebalance-every 250 --no-weak-baseline \
    > gpurun_out/r2k_bench_4gpu_dam64m_recut250.json 2> gpurun_out/r2k_bench_4gpu_dam64m_recut250.err; echo "64M 4gpu rc=$?") &
(CUDA_VISIBLE_DEVICES=4,5 SC_DIST_PDL=0 timeout 600 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --steps 200 --warmup 10 --no-weak-baseline \
    > gpurun_out/r2k_bench_2gpu_nopdl_exchange.json 2> gpurun_out/r2k_bench_2gpu_nopdl_exchange.err; echo "2gpu nopdl rc=$?") &
(CUDA_VISIBLE_DEVICES=6,7 timeout 600 $TR --nproc-per-node 2 --master-port 29526 bench.py --gpus 2 --steps 200 --warmup 10 \
    > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "2gpu rc=$?") &
wait
for f in gpurun_out/r2k_*.err; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
