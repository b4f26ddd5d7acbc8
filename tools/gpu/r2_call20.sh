mkdir -p gpurun_out /tmp/ix
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/cfg5_sharded.py --workdir /tmp/ix --mode staged --ids32 > gpurun_out/r2_cfg5_v3_n8_staged_u32.json 2> gpurun_out/r2_cfg5_v3_n8_staged_u32.err; tail -2 gpurun_out/r2_cfg5_v3_n8_staged_u32.err; cat gpurun_out/r2_cfg5_v3_n8_staged_u32.json
