timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in base mb7 mb6 mb5; do
  case $v in
    base) export RT_FUSE_SHADE=0; unset RT_LIB_PATH;;
    mb7) unset RT_FUSE_SHADE; unset RT_LIB_PATH;;
    mb6) export RT_LIB_PATH=$PWD/tests/_variant_mb6.so;;
    mb5) export RT_LIB_PATH=$PWD/tests/_variant_mb5.so;;
  esac
  echo "== $v"
  python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('synth1m', round(d['value']), round(d['ms_per_step'],4), d['roofline']['kernel_ms'])"
  python bench.py --workload blub4k --steps 20 --warmup 3 --no-cpu --no-others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('blub4k', round(d['value']), round(d['ms_per_step'],4))"
  python tests/gpu_share_probe.py synth1m 8 2>&1 | grep -v "128, 32\|32, 32" | cut -c1-120
done
