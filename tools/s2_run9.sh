mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_audio_gpu.py tests/test_wavenet_gpu.py tests/test_tacotron_gpu.py -x -q > gpurun_out/s5_test.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s5_test.log | cut -c1-300
