mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-300
timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-torch-eager > gpurun_out/bench_pool.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_pool.log') if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], {k:(round(v.get('ms') or v.get('ms_per_launch') or 0,4), round(v.get('frac') or 0,3)) for k,v in d['kernels'].items()})
PY
