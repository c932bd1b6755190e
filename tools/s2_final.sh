mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s7_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_gputest.log
tail -4 gpurun_out/s7_gputest.log
for W in wavenet_ce wavenet_default wavenet_mol tacotron; do
  timeout 400 python bench.py --workload $W > gpurun_out/s7_bench_$W.json 2> gpurun_out/s7_bench_$W.err; echo "$W rc=$?"
done
full() {  # name regex skip cmd...
  n=$1; r=$2; sk=$3; shift 3
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$r" -s $sk -c 1 -o gpurun_out/r2_full_$n -f "$@" > gpurun_out/r2_ncu_full_$n.log 2>&1
  echo "$n rc=$?"
}
full gate 'act_gemm2_kernel<\(int\)0' 40 python bench.py --workload wavenet_ce --steps 2 --warmup 3 --no-graph --no-cpu-baseline
full dx   'act_gemm2_kernel<\(int\)7' 40 python bench.py --workload wavenet_ce --steps 2 --warmup 3 --no-graph --no-cpu-baseline
full stft 'stft_mel_kernel_v2' 2 python tools/bench_audio.py
full gru  'gru_fwd_kernel' 1 python tools/bench_taco.py 1 --linear
cat gpurun_out/s7_bench_*.json | cut -c1-170
