mkdir -p gpurun_out /tmp/ix
for v in default minb5; do
  if [ $v = default ]; then unset SSHASH_GPU_LIB; else export SSHASH_GPU_LIB=$PWD/gpurun_ab/libsshash_gpu_$v.so; fi
  echo "== $v"
  SSHASH_GPU_BINNED=0 python tools/bench_configs.py --configs k63_3e9 --workdir /tmp/ix --keep 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('cfg4', {k: round(v['lookups_per_s']/1e9,2) for k,v in r['gpu'].items()})"
done
