python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c28_pytest.log | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c28_bench1.json 2> gpurun_out/c28_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c28_bench1.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v['ms']/d['steps'],2) for k,v in d['roofline']['per_op'].items()})
PY
timeout 300 python tools/step_ops.py rmat 64 > gpurun_out/c28_ops.log 2>&1; grep "pair_conv\|seg_reduce\|pair_dw\|total" gpurun_out/c28_ops.log
