#!/bin/bash
# One multi-GPU gpurun call (run from the repo root):  gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_multi_checks.sh r02 8'
# Every stage has a short own timeout AND bench.py's --deadline watchdog: a hung collective can never hold N GPUs for long.
set -u
TAG="${1:-rXX}"; N="${2:-8}"
OUT=gpurun_out
mkdir -p "$OUT"
PORT=29520
run() { # name, timeout_s, bench args...
  local name="$1" t="$2"; shift 2
  PORT=$((PORT + 1))
  echo "=== $name ($(date +%T))"
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" \
      bench.py --gpus "$N" --deadline $((t - 15)) "$@" > "$OUT/${TAG}_n${N}_${name}.log" 2>&1
  echo "    rc=$? -> $OUT/${TAG}_n${N}_${name}.log"
}
# N GPUs are charged N times: STAGES picks what a call re-measures (default: everything)
STAGES="${STAGES:-strong weak weak_peer render lpips step_check}"
has() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
has strong    && run strong  100 --scaling strong --n_rand_global 8192 --steps 20 --warmup 5 --no_cpu_baseline
has weak      && run weak    100 --steps 20 --warmup 5 --no_cpu_baseline
has weak_peer && SPN_P2P_ALLREDUCE=1 run weak_peer 100 --steps 20 --warmup 5 --no_cpu_baseline
has render    && run render  100 --workload render --steps 3 --warmup 3
has lpips     && run lpips   100 --workload train_lpips --steps 10 --warmup 3
if has step_check; then
echo "=== check_multi_gpu_step"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29539 \
    tools/check_multi_gpu_step.py > "$OUT/${TAG}_n${N}_step_check.log" 2>&1
echo "    rc=$?"; tail -n 5 "$OUT/${TAG}_n${N}_step_check.log"
fi
if [ -f tools/check_peer_allreduce.py ] && [ "${SPN_CHECK_PEER:-0}" = "1" ]; then
  echo "=== peer_allreduce"
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29540 \
      tools/check_peer_allreduce.py > "$OUT/${TAG}_n${N}_peer.log" 2>&1
  echo "    rc=$?"
fi
grep -h '^{' "$OUT/${TAG}_n${N}"_*.log > "$OUT/${TAG}_n${N}_bench_lines.jsonl" 2>/dev/null
tail -n 2 "$OUT/${TAG}_n${N}"_*.log
