mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/s2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_gputest.log
for W in wavenet_ce wavenet_default wavenet_mol tacotron; do
  timeout 400 python bench.py --workload $W > gpurun_out/s2_bench_$W.json 2> gpurun_out/s2_bench_$W.err; echo "$W rc=$?"
done
tail -5 gpurun_out/s2_gputest.log; cat gpurun_out/s2_bench_*.json | cut -c1-600
