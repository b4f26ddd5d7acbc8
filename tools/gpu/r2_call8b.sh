mkdir -p gpurun_out /tmp/ix
python - <<'PY' 2> gpurun_out/r2_cfg4_build.err
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from bench_configs import build_index
print(build_index("/tmp/ix", 3000000, 1062, 63, 25))
PY
K=/tmp/ix/synth_3000000_1062_k63_m25.sshash
ls -la $K
for m in fwd mix neg; do
ncu --set full --clock-control none -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_cfg4_k63_3e9_$m python tools/ncu_target.py --index $K --mode $m --max-k 63 > gpurun_out/r2_cfg4_ncu_$m.log 2>&1; tail -2 gpurun_out/r2_cfg4_ncu_$m.log
done
ls -la gpurun_out/
