# 8xB200, final build of the round: configs[3] (box-fill 16M) and configs[4] (dam-break 64M, re-cut every 250), then on
# disjoint GPUs the 4- and 2-GPU points of the box-fill weak-scaling series and the strips' bit-parity checks.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 200 --warmup 10 \
    > gpurun_out/r2z_bench_8gpu_boxfill16m.json 2> gpurun_out/r2z_bench_8gpu_boxfill16m.err; echo "16M box rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --scene dam_break_wide --relax 4000 --warmup 10 --steps 200 --particles 8000000 --rebalance-every 250 \
    > gpurun_out/r2z_bench_8gpu_dam64m_recut250.json 2> gpurun_out/r2z_bench_8gpu_dam64m_recut250.err; echo "64M recut rc=$?"
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 200 --warmup 10 \
    > gpurun_out/r2z_bench_4gpu_boxfill8m.json 2> gpurun_out/r2z_bench_4gpu_boxfill8m.err; echo "4gpu rc=$?") &
(CUDA_VISIBLE_DEVICES=4,5 timeout 600 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --steps 200 --warmup 10 \
    > gpurun_out/r2z_bench_2gpu_final.json 2> gpurun_out/r2z_bench_2gpu_final.err; echo "2gpu rc=$?") &
(CUDA_VISIBLE_DEVICES=6,7 SC_TRANSPORT=p2p timeout 600 $TR --nproc-per-node 2 --master-port 29526 tests/mgpu_check.py \
    > gpurun_out/r2z_mgpu_check_2gpu_p2p.log 2>&1; echo "mgpu rc=$?"; grep -c "bit-identical to single GPU = True" gpurun_out/r2z_mgpu_check_2gpu_p2p.log) &
wait
for f in gpurun_out/r2z_*.err; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
