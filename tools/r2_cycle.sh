#!/bin/bash
# One measure cycle on the GPU box: parity subset -> bench (no ncu) -> one ncu capture of the tile kernel.
# Usage: tools/r2_cycle.sh <tag> [ncu-kernel-regex] [case]
set -u
TAG=$1; REGEX=${2:-tile_kernel}; CASE=${3:-cfg4}
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_parity or launch_shape or chunk_boundary or invalid_h or hybrid_equals or small_fill" > gpurun_out/r2/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r2/${TAG}_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/${TAG}_bench.json 2> gpurun_out/r2/${TAG}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2/${TAG}_bench.json"))
print("value %.4e  ms/step %.3f" % (d["value"], d["ms_per_step"]))
print({k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
PY
if [ "$REGEX" != "none" ]; then
  bash tools/ncu_case.sh r2_${TAG} $CASE "$REGEX" 1 1; mv gpurun_out/r2_${TAG}.ncu-rep gpurun_out/r2/${TAG}.ncu-rep
  python tools/ncu_summary.py gpurun_out/r2/${TAG}.ncu-rep > gpurun_out/r2/${TAG}_ncu.txt
  grep -E "duration|fp64_pipe|issue_active|lanes_per|no_instruction|stall_barrier|stall_wait|dram_read|dram_write|warps_active|fp64_tflops" gpurun_out/r2/${TAG}_ncu.txt
  python tools/ncu_hot.py gpurun_out/r2/${TAG}.ncu-rep "$REGEX" --top 30 > gpurun_out/r2/${TAG}_hot.txt 2>&1; python tools/ncu_opmix.py gpurun_out/r2/${TAG}.ncu-rep "$REGEX" > gpurun_out/r2/${TAG}_opmix.txt 2>&1
  head -12 gpurun_out/r2/${TAG}_hot.txt
fi
