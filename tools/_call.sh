timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --deadline 120 --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r02e_n2.log 2> gpurun_out/r02e_n2.err; echo "rc=$?"; tail -4 gpurun_out/r02e_n2.err | cut -c1-200
python - <<'PY'
import json
ls=[l for l in open('gpurun_out/r02e_n2.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d.get('aborted'))
PY
sleep 3
SPN_GRAPH_E2E=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --deadline 120 --steps 20 --warmup 5 --no_cpu_baseline 2>gpurun_out/r02e_n2b.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('eager e2e', d['value'], d['ms_per_step'], d['e2e'])"
