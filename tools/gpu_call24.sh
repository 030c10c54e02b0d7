python -m pytest tests -m gpu -q > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c24_pytest.log | tail -5
DRY_LIST=1 timeout 600 python tools/rowshard_dry.py 8 64 0,5 > gpurun_out/c24_dry.log 2>&1; echo "dry rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c24_dry.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c24_bench1.json 2> gpurun_out/c24_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c24_bench1.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v['ms']/d['steps'],2) for k,v in d['roofline']['per_op'].items()})
PY
DRY_KINETO=1 timeout 600 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c24_kineto.log 2>&1; echo "kineto rc=$?"; grep -v Warn gpurun_out/c24_kineto.log | tail -48
