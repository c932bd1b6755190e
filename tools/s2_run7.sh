mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_cbhg_gpu.py -x -q -s > gpurun_out/s3_cbhg_test.log 2>&1; echo "cbhg pytest rc=$?"
grep -E "MEASURED|passed|failed|Error|error|assert" gpurun_out/s3_cbhg_test.log | cut -c1-600 | tail -30
