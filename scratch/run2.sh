set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wide_view or feedback or render_push or sharding" ) > gpurun_out/r2d_pytest_subset.log 2>&1; tail -8 gpurun_out/r2d_pytest_subset.log
PROBE_HEAD=1 timeout 600 python tests/gpu_wide_heavy_probe.py > gpurun_out/r2d_wide_heavy_probe.log 2>&1; cat gpurun_out/r2d_wide_heavy_probe.log | cut -c1-1500
