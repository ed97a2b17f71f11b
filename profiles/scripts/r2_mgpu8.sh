# 8xB200: BASELINE.json configs[4] (dam-break 64M) with / without re-cutting, a 4-GPU strong-scaling point of the same
# scene, the re-cutter's bit-parity at scale on 2 GPUs, and configs[3] (box-fill 16M).  One gpurun --gpus 8 call.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
W="--scene dam_break_wide --relax 4000 --warmup 10 --steps 200"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 $W --particles 8000000 --rebalance-every 50 \
    > gpurun_out/r2g_bench_8gpu_dam64m_recut50.json 2> gpurun_out/r2g_bench_8gpu_dam64m_recut50.err; echo "64M recut rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 $W --particles 8000000 --rebalance-every 0 --no-weak-baseline \
    > gpurun_out/r2g_bench_8gpu_dam64m_norecut.json 2> gpurun_out/r2g_bench_8gpu_dam64m_norecut.err; echo "64M no recut rc=$?"
# strong scaling: the same 64M scene on 4 GPUs (0-3), while GPUs 4,5 check the re-cutter's bit-parity at 2M
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 $W --particles 16000000 --rebalance-every 50 --no-weak-baseline \
    > gpurun_out/r2g_bench_4gpu_dam64m_recut50.json 2> gpurun_out/r2g_bench_4gpu_dam64m_recut50.err; echo "64M 4gpu rc=$?") &
(CUDA_VISIBLE_DEVICES=4,5 SC_CHECK_SCALE=1 SC_TRANSPORT=p2p timeout 900 $TR --nproc-per-node 2 --master-port 29524 tests/mgpu_check.py \
    > gpurun_out/r2g_mgpu_scale_2gpu.log 2>&1; echo "mgpu scale rc=$?"; grep "\[mgpu\]" gpurun_out/r2g_mgpu_scale_2gpu.log) &
wait
timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 200 --warmup 10 \
    > gpurun_out/r2g_bench_8gpu_boxfill16m.json 2> gpurun_out/r2g_bench_8gpu_boxfill16m.err; echo "16M box rc=$?"
tail -2 gpurun_out/*.err | cut -c1-300
