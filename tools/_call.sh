timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 400 python bench.py --no_cpu_baseline > gpurun_out/r02g_bench_n1.log 2> gpurun_out/r02g_bench_n1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02g_bench_n1.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['roofline']['kernels'].items(): print(k, round(v['ms_per_step'],3), v['bound'], round(v['frac'],3))
print(d['psnr_vs_ref']['ert_skippable_fine_sample_frac'], d['psnr_vs_ref']['trained_psnr_db'])
print({k:(v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d['other_workloads'].items()})
PY
