#!/bin/bash
# compute-sanitizer passes over tools/sanitize_case.py (run under gpurun, one GPU).
#   1. plain run (must exit 0 before any tool run counts)
#   2. memcheck, default launcher (one-launch small-fill kernel for the 4e4-element fills)
#   3. memcheck, pipeline forced: no small-fill kernel, 2^13-element passes, buckets on the helper streams
#   4. racecheck (shared-memory hazards: bucket_kernel, small_kernel, Gram kernels), pipeline forced
#   5. synccheck, 6. initcheck (library kernels only matter; torch's own kernels run too)
# Logs: gpurun_out/sanitize_*.log; summary lines at the end of stdout.
# NOTE: the round-1 GPU pool refuses compute-sanitizer (rc 86, "closed on this pool"); only step 1 ran there.
# tests/test_gpu_guard_bands.py looks for out-of-bounds stores directly instead.  The script is kept for boxes
# where the tool is allowed.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
export PYTORCH_NO_CUDA_MEMORY_CACHING=1   # one cudaMalloc per tensor: out-of-bounds accesses are not hidden by the pool
run() {  # name, time limit, env..., -- tool args
    local name=$1 limit=$2; shift 2
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout "$limit" "$CS" "$@" --error-exitcode 9 --log-file "gpurun_out/sanitize_${name}.log" \
        python tools/sanitize_case.py > "gpurun_out/sanitize_${name}.out" 2>&1
    local rc=$?
    echo "== ${name}: rc=${rc} $(tail -1 gpurun_out/sanitize_${name}.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${name}.log | tail -1)"
}
python tools/sanitize_case.py > gpurun_out/sanitize_plain.out 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.out; exit 1; }
echo "== plain: $(tail -1 gpurun_out/sanitize_plain.out)"
FORCE=(PGM_CUDA_SMALL_MAX=0 PGM_CUDA_CHUNK_LOG2=13 PGM_CUDA_MULTI_MIN_LOG2=10)
run memcheck_small 110 SANITIZE_N=40000 -- --tool memcheck --leak-check no
run memcheck_pipeline 110 SANITIZE_N=40000 "${FORCE[@]}" -- --tool memcheck --leak-check no
run racecheck 110 SANITIZE_N=12000 "${FORCE[@]}" -- --tool racecheck --racecheck-report all
run synccheck 70 SANITIZE_N=12000 "${FORCE[@]}" -- --tool synccheck
run initcheck 90 SANITIZE_N=12000 "${FORCE[@]}" -- --tool initcheck
