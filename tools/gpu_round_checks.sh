#!/bin/bash
# One gpurun call that re-establishes every measured artefact of a round on a fresh B200 (run from the repo root):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_checks.sh r02'
#
# Writes under gpurun_out/<tag>_*: the GPU test log, smoke log, bench lines (N = 1: train / render / train_lpips, reference
# arm), the ncu launch list of the bench command, one `--set full` capture of the three MLP kernels, and compute-sanitizer
# memcheck + racecheck over smoke().  Each stage has its own timeout so that one hang cannot eat the whole call; numbers
# printed under ncu / sanitizer are never bench values.  Reduce the ncu outputs with tools/ncu_summary.py and copy the
# summaries to profiles/.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
run() { # name, timeout_s, command...
  local name="$1" t="$2"; shift 2
  echo "=== $name ($(date +%T))"
  timeout "$t" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
  echo "    rc=$? -> $OUT/${TAG}_${name}.log"
}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/${TAG}_gpu.txt" 2>&1
run tests_gpu      600 python -m pytest tests -q -m gpu -x
run smoke          300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1       420 python bench.py --gpus 1
run bench_ref      420 python bench.py --impl reference --steps 3 --warmup 1
run bench_render   300 python bench.py --workload render --steps 6 --warmup 3
run bench_lpips    300 python bench.py --workload train_lpips --steps 30 --warmup 5
run bench_sustained 300 python bench.py --gpus 1 --steps 2000 --warmup 5 --no_cpu_baseline
run bench_strong1  300 python bench.py --gpus 1 --scaling strong --n_rand_global 8192 --steps 20 --warmup 5 --no_cpu_baseline
run ncu_launches   420 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 250 --csv \
                       --log-file "$OUT/${TAG}_launches.csv" python bench.py --steps 8 --warmup 3 --no_cpu_baseline --deadline 0
run ncu_full       600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_(fwd_ts|dgrad|wgrad)_kernel' -s 18 -c 6 \
                       -o "$OUT/${TAG}_mlp" -f python bench.py --steps 8 --warmup 3 --no_cpu_baseline --deadline 0
run san_memcheck   600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()"
run san_racecheck  900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()"
# multi-GPU only (gpurun --gpus 2): the opt-in peer-memory gradient exchange against NCCL
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  run peer_allreduce 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_peer_allreduce.py
  run bench_n2       420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2
fi
grep -h '^{' "$OUT/${TAG}"_bench_*.log > "$OUT/${TAG}_bench_lines.jsonl" 2>/dev/null
tail -n 3 "$OUT/${TAG}"_tests_gpu.log "$OUT/${TAG}"_smoke.log "$OUT/${TAG}"_san_*.log
