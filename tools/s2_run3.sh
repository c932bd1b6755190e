mkdir -p gpurun_out
timeout 200 python tools/kernel_scaling.py > gpurun_out/s2_kernel_scaling.txt 2>&1; echo "scaling rc=$?"; cat gpurun_out/s2_kernel_scaling.txt | tail -5
full() {  # name regex skip workload
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/r2_full_$1 -f \
    python bench.py --workload $4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_full_$1.log 2>&1
  echo "$1 rc=$?"
}
full gate  'act_gemm2_kernel<\(int\)0'  40 wavenet_ce
full out   'act_gemm2_kernel<\(int\)1'  40 wavenet_ce
if ! ls gpurun_out/r2_full_gate.ncu-rep >/dev/null 2>&1; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:act_gemm2_kernel -s 313 -c 2 -o gpurun_out/r2_full_gateout -f \
    python bench.py --workload wavenet_ce --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_full_gateout.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
