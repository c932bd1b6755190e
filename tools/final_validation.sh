set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-graph > gpurun_out/final_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:act_gemm_kernel -s 30 -c 2 -o gpurun_out/final_gate_full python bench.py --steps 1 --warmup 1 --no-graph > gpurun_out/final_ncu_b.log 2>&1
timeout 300 python tools/bench_ar.py > gpurun_out/final_ar.jsonl 2> gpurun_out/final_ar.err
timeout 120 python tools/bench_audio.py > gpurun_out/final_audio.json 2> gpurun_out/final_audio.err
timeout 200 python tools/bench_taco.py --graph > gpurun_out/final_taco.json 2> gpurun_out/final_taco.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/final_taco_launches.csv python tools/taco_one_step.py 100 > gpurun_out/final_taco_ncu.log 2>&1
cat gpurun_out/final_pytest.log; tail -1 gpurun_out/final_smoke.log; cut -c1-300 gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench_ref.json; cut -c1-200 gpurun_out/final_taco.json; ls -la gpurun_out/final_*
