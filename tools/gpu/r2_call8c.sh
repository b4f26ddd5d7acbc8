mkdir -p gpurun_out /tmp/ix
python tools/bench_configs.py --configs k63_3e9 --workdir /tmp/ix --keep > gpurun_out/r2_cfg4_k63_3e9.jsonl 2> gpurun_out/r2_cfg4_k63_3e9.err; tail -3 gpurun_out/r2_cfg4_k63_3e9.err; cat gpurun_out/r2_cfg4_k63_3e9.jsonl
K=/tmp/ix/synth_3000000_1062_k63_m25.sshash
SSHASH_GPU_BINNED=1 python tools/exp_locality.py --strings 3000000 --length 1062 -k 63 -m 25 --workdir /tmp/ix --no-sorted --variants binned --queries 50000000 > gpurun_out/r2_cfg4_binned.jsonl 2> gpurun_out/r2_cfg4_binned.err; tail -2 gpurun_out/r2_cfg4_binned.err; cat gpurun_out/r2_cfg4_binned.jsonl
for m in fwd mix; do
ncu --set full --clock-control none -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_cfg4_k63_3e9_$m python tools/ncu_target.py --index $K --mode $m --max-k 63 > gpurun_out/r2_cfg4_ncu_$m.log 2>&1; tail -2 gpurun_out/r2_cfg4_ncu_$m.log
done
ls -la gpurun_out/
