mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "binned or lookup_device or multi_partition or large_batch" > gpurun_out/r2_c2_pytest.log 2>&1; tail -5 gpurun_out/r2_c2_pytest.log
python tools/exp_locality.py --strings 500000 --length 1030 -k 31 -m 17 --workdir /tmp/ix --no-sorted > gpurun_out/r2_ab_t5e8.jsonl 2> gpurun_out/r2_ab_t5e8.err; tail -3 gpurun_out/r2_ab_t5e8.err
python tools/exp_locality.py --workdir /tmp/ix --no-sorted > gpurun_out/r2_ab_human.jsonl 2> gpurun_out/r2_ab_human.err; tail -3 gpurun_out/r2_ab_human.err
cat gpurun_out/r2_ab_t5e8.jsonl gpurun_out/r2_ab_human.jsonl
H=/tmp/ix/synth_2500000_1030_k31_m21.sshash
SSHASH_GPU_BINNED=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_human_binned_mix_launches.csv python tools/ncu_target.py --index $H --mode mix --launches 2 > /dev/null 2>&1
SSHASH_GPU_BINNED=0 ncu --set full --clock-control none --import-source on -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_human_direct_mix python tools/ncu_target.py --index $H --mode mix > /dev/null 2>&1
SSHASH_GPU_BINNED=1 ncu --set full --clock-control none --import-source on -k regex:lookup_binned_kernel -s 2 -c 1 -f -o gpurun_out/r2_human_binned_mix python tools/ncu_target.py --index $H --mode mix --launches 2 > /dev/null 2>&1
ls -la gpurun_out/
