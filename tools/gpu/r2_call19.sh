mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_c19_pytest.log 2>&1; tail -2 gpurun_out/r2_c18_pytest.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_v3.json 2> gpurun_out/r2_bench_ref_v3.err
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err ) 2>&1 | grep real; tail -2 gpurun_out/r2_bench_v3.err; cat gpurun_out/r2_bench_v3.json | cut -c1-1500
ncu --set full --clock-control none --import-source on -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_cfg2_lookup_v3 python tools/ncu_target.py --index tests/golden/se_k31_m13.sshash --mode mix > /dev/null 2>&1
ls -la gpurun_out
