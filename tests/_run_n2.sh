timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "render_push or handshake or packed_tiles or sharding" 2>&1 | tail -3
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --no-others "${@:2}" 2>gpurun_out/n2_$1.err | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value']), round(d['ms_per_step'],4), 'p50', round(d['ms_per_step_p50'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"; }
echo "== inline push, tile 64x32"; run 29602 --tile 64 32
echo "== inline push, tile 32x32"; run 29601 --tile 32 32
echo "== separate push (RT_FUSE_SHADE=0), tile 64x32"; RT_FUSE_SHADE=0 run 29604 --tile 64 32
echo "== separate push (RT_FUSE_SHADE=0), tile 32x32"; RT_FUSE_SHADE=0 run 29603 --tile 32 32
