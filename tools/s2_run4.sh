mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_wavenet_gpu.py tests/test_gemm_engine.py -x -q -m gpu > gpurun_out/s2_wn_test.log 2>&1; echo "wn pytest rc=$?"
tail -15 gpurun_out/s2_wn_test.log
timeout 200 python bench.py --workload wavenet_ce --no-cpu-baseline > gpurun_out/s2_bench_ce_tma.json 2>gpurun_out/s2_bench_ce_tma.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/s2_bench_ce_tma.json; python -c "
import json;d=json.loads(open('gpurun_out/s2_bench_ce_tma.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
