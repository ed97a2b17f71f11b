# 2xB200: the single-GPU suite, the strips' bit-parity checks (both transports, and the re-cutter at scale) and the 2-GPU bench
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SC_TRANSPORT=p2p timeout 600 $TR tests/mgpu_check.py > gpurun_out/r2j_mgpu_p2p.log 2>&1; echo "mgpu p2p rc=$?"; grep "\[mgpu\]" gpurun_out/r2j_mgpu_p2p.log
SC_TRANSPORT=nccl timeout 600 $TR tests/mgpu_check.py > gpurun_out/r2j_mgpu_nccl.log 2>&1; echo "mgpu nccl rc=$?"; grep "\[mgpu\]" gpurun_out/r2j_mgpu_nccl.log
SC_CHECK_SCALE=1 SC_TRANSPORT=p2p timeout 900 $TR tests/mgpu_check.py > gpurun_out/r2j_mgpu_scale.log 2>&1; echo "mgpu scale rc=$?"; grep "\[mgpu\]" gpurun_out/r2j_mgpu_scale.log
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench2 rc=$?"
SC_DIST_DEFER=0 timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-weak-baseline > gpurun_out/r2j_bench_2gpu_nodefer.json 2> gpurun_out/r2j_bench_2gpu_nodefer.err; echo "bench2 nodefer rc=$?"
