mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_audio_gpu.py -x -q > gpurun_out/s2_audio_test.log 2>&1; echo "audio pytest rc=$?"
tail -3 gpurun_out/s2_audio_test.log
timeout 200 python tools/bench_audio.py > gpurun_out/s2_audio_bench_v2.json 2>gpurun_out/s2_audio_bench_v2.err; echo rc=$?
T2_STFT_V1=1 timeout 200 python tools/bench_audio.py > gpurun_out/s2_audio_bench_v1.json 2>/dev/null
cat gpurun_out/s2_audio_bench_v2.json gpurun_out/s2_audio_bench_v1.json | cut -c1-400
timeout 300 python tools/gap_probe.py > gpurun_out/s2_gap_probe.txt 2>&1; echo "gap rc=$?"
head -120 gpurun_out/s2_gap_probe.txt
