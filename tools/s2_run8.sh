mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_cbhg_gpu.py tests/test_audio_gpu.py -x -q -s > gpurun_out/s4_test_a.log 2>&1; echo "pytest A rc=$?"; tail -4 gpurun_out/s4_test_a.log
timeout 700 python -m pytest tests/test_entrypoints_gpu.py -x -q -k linear_head > gpurun_out/s4_test_b.log 2>&1; echo "pytest B rc=$?"; tail -25 gpurun_out/s4_test_b.log | cut -c1-300
grep -h "MEASURED\|griffin" gpurun_out/s4_test_a.log | cut -c1-400
