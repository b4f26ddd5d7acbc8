mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "binned or u32 or replay or multi_gpu" > gpurun_out/r2_c7_pytest.log 2>&1; tail -5 gpurun_out/r2_c6_pytest.log
python tools/exp_locality.py --strings 500000 --length 1030 -k 31 -m 17 --workdir /tmp/ix --no-sorted --variants direct,binned > gpurun_out/r2_ab5_t5e8.jsonl 2> gpurun_out/r2_ab5_t5e8.err; tail -3 gpurun_out/r2_ab5_t5e8.err
python tools/exp_locality.py --workdir /tmp/ix --no-sorted --variants direct,binned,binned_noprefetch > gpurun_out/r2_ab5_human.jsonl 2> gpurun_out/r2_ab5_human.err; tail -3 gpurun_out/r2_ab5_human.err
cat gpurun_out/r2_ab5_t5e8.jsonl gpurun_out/r2_ab5_human.jsonl
H=/tmp/ix/synth_2500000_1030_k31_m21.sshash
SSHASH_GPU_BINNED=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_human_binned4_mix_launches.csv python tools/ncu_target.py --index $H --mode mix --launches 2 > /dev/null 2>&1
ls -la gpurun_out/
