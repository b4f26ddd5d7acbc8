mkdir -p gpurun_out /tmp/ix
for v in default minb5 minb8; do
  if [ $v = default ]; then unset SSHASH_GPU_LIB; else export SSHASH_GPU_LIB=$PWD/gpurun_ab/libsshash_gpu_$v.so; fi
  echo "== $v"
  python tools/scale_bench.py --index tests/golden/se_k31_m13.sshash 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('cfg2', {k: round(r[k]['lookups_per_s']/1e9,2) for k in ('positive_forward','positive_50rc','negative')})"
  python tools/exp_locality.py --strings 500000 --length 1030 -k 31 -m 17 --workdir /tmp/ix --no-sorted --variants direct 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    if 'variant' in r: print('t5e8', {q: round(r[q]['G_lookups_per_s'],2) for q in ('fwd','mix','neg')})"
  python tools/exp_locality.py --workdir /tmp/ix --no-sorted --variants direct 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    if 'variant' in r: print('human', {q: round(r[q]['G_lookups_per_s'],2) for q in ('fwd','mix','neg')})"
done
