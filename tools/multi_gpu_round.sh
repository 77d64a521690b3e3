#!/bin/bash
# Everything that needs several GPUs, in one gpurun --gpus N call.   Usage: tools/multi_gpu_round.sh N
N=$1
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -9
python tools/h2d_peak.py
for k in 2 4 8; do [ $k -le $N ] && $TR --nproc-per-node $k --master-port $((29500 + k)) tools/h2d_peak.py 2>/dev/null | tail -1; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical_across_devices or host_multi_device" 2>&1 | tail -2
for k in 1 2 4 8; do
  [ $k -le $N ] || continue
  if [ $k -eq 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2/scale_$k.json 2> gpurun_out/r2/scale_$k.err
  else $TR --nproc-per-node $k --master-port $((29600 + k)) bench.py --gpus $k --steps 5 --warmup 3 > gpurun_out/r2/scale_$k.json 2> gpurun_out/r2/scale_$k.err; fi
  python - <<PY
import json
d = json.load(open("gpurun_out/r2/scale_$k.json"))
print("N=$k value %.4e ms/step %.3f e2e %.4e checksum %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["shard_checksum"]["sum_u64"], d["shard_checksum"]["xor_u64"]))
PY
done
python - <<'PY'
# one host call over all devices (single process): pgm_cuda_random_polyagamma_fill2_host_multi
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import polyagamma_b200 as pg
n = 1 << 28
hh = torch.empty(n, dtype=torch.float64, pin_memory=True); zz = torch.empty(n, dtype=torch.float64, pin_memory=True); oo = torch.empty(n, dtype=torch.float64, pin_memory=True)
rng = np.random.default_rng(1)
hh.numpy()[:] = rng.uniform(0.1, 200, n); zz.numpy()[:] = rng.uniform(-50, 50, n)
for nd in (1, 2, 4, 8):
    if nd > torch.cuda.device_count(): break
    devs = list(range(nd))
    pg.random_polyagamma(hh.numpy(), zz.numpy(), out=oo.numpy(), random_state=(5, 5), disable_checks=True, devices=devs)
    t0 = time.perf_counter()
    for rep in range(3):
        pg.random_polyagamma(hh.numpy(), zz.numpy(), out=oo.numpy(), random_state=(5, 5), disable_checks=True, devices=devs)
    dt = (time.perf_counter() - t0) / 3
    print("host_multi devices=%d: %.3e samples/s (%.1f GB/s over the host link), checksum %016x" % (nd, n / dt, 24 * n / dt / 1e9,
          int(np.bitwise_xor.reduce(oo.numpy().view(np.uint64)))))
PY
