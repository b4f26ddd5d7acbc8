mkdir -p gpurun_out
python tools/exp_locality.py --workdir /tmp/ix --no-sorted --variants direct,direct+prefix_window > gpurun_out/r2_ab3_human.jsonl 2> gpurun_out/r2_ab3_human.err; tail -3 gpurun_out/r2_ab3_human.err
cat gpurun_out/r2_ab3_human.jsonl
H=/tmp/ix/synth_2500000_1030_k31_m21.sshash
SSHASH_GPU_BINNED=1 ncu --set full --clock-control none -k regex:"unpermute_kernel|bin_scatter_kernel" -c 4 -f -o gpurun_out/r2_human_binned2_aux python tools/ncu_target.py --index $H --mode mix --launches 1 > /dev/null 2>&1
ls -la gpurun_out/
