mkdir -p gpurun_out /tmp/ix
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_v1_n8.json 2> gpurun_out/r2_bench_v1_n8.err; tail -2 gpurun_out/r2_bench_v1_n8.err; cat gpurun_out/r2_bench_v1_n8.json | cut -c1-5000
$T --master-port 29517 tools/cfg5_sharded.py --workdir /tmp/ix --mode copy --ids32 > gpurun_out/r2_cfg5_v2_n8_copy_u32.json 2> gpurun_out/r2_cfg5_v2_n8_copy_u32.err; tail -2 gpurun_out/r2_cfg5_v2_n8_copy_u32.err; cat gpurun_out/r2_cfg5_v2_n8_copy_u32.json
$T --master-port 29518 tools/cfg5_sharded.py --workdir /tmp/ix --mode copy --ids32 --chunk 4194304 --oracle-sample 0 > gpurun_out/r2_cfg5_v2_n8_copy_u32_c22.json 2> gpurun_out/r2_cfg5_v2_n8_copy_u32_c22.err; cat gpurun_out/r2_cfg5_v2_n8_copy_u32_c22.json
$T --master-port 29519 tools/cfg5_sharded.py --workdir /tmp/ix --mode copy --oracle-sample 0 > gpurun_out/r2_cfg5_v2_n8_copy_u64.json 2> gpurun_out/r2_cfg5_v2_n8_copy_u64.err; cat gpurun_out/r2_cfg5_v2_n8_copy_u64.json
ls -la gpurun_out
