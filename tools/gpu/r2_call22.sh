mkdir -p gpurun_out /tmp/ix
python - <<'PY'
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from bench_configs import build_index
print(build_index("/tmp/ix", 500000, 1030, 31, 17))
PY
ncu --set full --clock-control none -k regex:lookup_kernel -s 2 -c 1 -f -o gpurun_out/r2_t5e8_mix python tools/ncu_target.py --index /tmp/ix/synth_500000_1030_k31_m17.sshash --mode mix > /dev/null 2>&1
ls -la gpurun_out | tail -3
