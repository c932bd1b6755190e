#!/bin/bash
# Round-2 ncu passes (run under gpurun on ONE GPU). Outputs land in gpurun_out/; tools/summarize_ncu.py condenses them into profiles/.
# 1. launch lists with DRAM bytes for one eager step of each workload (cold-cache, serialised: compare SHARES, not absolutes)
# 2. --set full captures (tensor pipe, L2, DRAM counters, source-level stalls) of every hot kernel
set -u
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for W in wavenet_ce wavenet_default; do
  timeout 600 ncu --metrics $M --clock-control none --cache-control none -s 700 -c 700 --csv --log-file gpurun_out/r2_launches_$W.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_$W.log 2>&1
done
timeout 900 ncu --metrics $M --clock-control none --cache-control none -s 20000 -c 9000 --csv --log-file gpurun_out/r2_launches_tacotron.csv \
  python bench.py --workload tacotron --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_tacotron.log 2>&1
# full captures: 2 launches of each hot kernel from the middle of the run
full() {  # name regex skip workload
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 2 -o gpurun_out/r2_full_$1 -f \
    python bench.py --workload $4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_full_$1.log 2>&1
}
full gate  'act_gemm2_kernel<\(int\)0'  40 wavenet_ce
full out   'act_gemm2_kernel<\(int\)1'  40 wavenet_ce
full dz    'act_gemm2_kernel<\(int\)6'  40 wavenet_ce
full dx    'act_gemm2_kernel<\(int\)7'  40 wavenet_ce
full wgrad 'wgrad_gemm_kernel'          3  wavenet_ce
full gate_default 'act_gemm2_kernel<\(int\)0'  30 wavenet_default
full lstm  'act_gemm_kernel<\(int\)8'   2000 tacotron
full tout  'act_gemm_kernel<\(int\)9'   2000 tacotron
full attf  'att_fwd_kernel'             1000 tacotron
full attb  'att_bwd_kernel'             1000 tacotron
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wn_ar_kernel -c 1 -o gpurun_out/r2_full_ar -f \
  python tools/bench_ar.py 256 > gpurun_out/r2_ncu_full_ar.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel -s 2 -c 1 -o gpurun_out/r2_full_stft -f \
  python tools/bench_audio.py > gpurun_out/r2_ncu_full_stft.log 2>&1
ls -la gpurun_out/r2_full_*.ncu-rep gpurun_out/r2_launches_*.csv
