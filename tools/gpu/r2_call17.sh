mkdir -p gpurun_out /tmp/ix
python -m pytest tests -m gpu -x -q > gpurun_out/r2_c17_pytest.log 2>&1; tail -3 gpurun_out/r2_c17_pytest.log
python tools/stream_bench.py --files > gpurun_out/r2_stream_v1.json 2> gpurun_out/r2_stream_v1.err; tail -2 gpurun_out/r2_stream_v1.err; cat gpurun_out/r2_stream_v1.json | cut -c1-1500
bash tools/gpu/r2_call16.sh
