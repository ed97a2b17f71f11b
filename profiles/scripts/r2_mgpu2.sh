set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log; tail -3 gpurun_out/r2f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SC_TRANSPORT=p2p timeout 600 $TR tests/mgpu_check.py > gpurun_out/r2f_mgpu_p2p.log 2>&1; echo "mgpu p2p rc=$?"; grep "\[mgpu\]" gpurun_out/r2f_mgpu_p2p.log
SC_TRANSPORT=nccl timeout 600 $TR tests/mgpu_check.py > gpurun_out/r2f_mgpu_nccl.log 2>&1; echo "mgpu nccl rc=$?"; grep "\[mgpu\]" gpurun_out/r2f_mgpu_nccl.log
SC_CHECK_SCALE=1 SC_TRANSPORT=p2p timeout 900 $TR tests/mgpu_check.py > gpurun_out/r2f_mgpu_scale.log 2>&1; echo "mgpu scale rc=$?"; grep "\[mgpu\]" gpurun_out/r2f_mgpu_scale.log
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?"
