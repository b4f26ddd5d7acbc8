mkdir -p gpurun_out
free -g | head -2; df -h /tmp | tail -1; nproc
python tools/bench_configs.py --configs k63_3e9 --workdir /tmp/ix --keep > gpurun_out/r2_cfg4_k63_3e9.jsonl 2> gpurun_out/r2_cfg4_k63_3e9.err; tail -3 gpurun_out/r2_cfg4_k63_3e9.err; cat gpurun_out/r2_cfg4_k63_3e9.jsonl
K=/tmp/ix/synth_3000000_1062_k63_m25.sshash
ncu --set full --clock-control none -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_cfg4_k63_3e9_mix python tools/ncu_target.py --index $K --mode mix --max-k 63 > /dev/null 2>&1
rm -f $K
ls -la gpurun_out/
