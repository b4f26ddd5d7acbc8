mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_c12_pytest.log 2>&1; tail -3 gpurun_out/r2_c12_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_v1.json 2> gpurun_out/r2_bench_ref_v1.err; cat gpurun_out/r2_bench_ref_v1.json | cut -c1-400
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err ) 2>&1 | grep real; tail -2 gpurun_out/r2_bench_v1.err; cat gpurun_out/r2_bench_v1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-scale --no-cpu-baseline --no-ncu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_cfg2_lookup python tools/ncu_target.py --index tests/golden/se_k31_m13.sshash --mode mix > /dev/null 2>&1
ls -la gpurun_out
