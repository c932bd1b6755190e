mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -x -q -s > gpurun_out/s6_dp_test.log 2>&1; echo "dp pytest rc=$?"; grep -E "DP2|passed|failed|Error" gpurun_out/s6_dp_test.log | cut -c1-300 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/s6_bench_ce_2gpu.json 2> gpurun_out/s6_bench_ce_2gpu.err; echo "bench2 rc=$?"; tail -1 gpurun_out/s6_bench_ce_2gpu.json | cut -c1-260; tail -3 gpurun_out/s6_bench_ce_2gpu.err | cut -c1-200
