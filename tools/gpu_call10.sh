python tools/one_step.py > gpurun_out/c10_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/one_step.py > gpurun_out/c10_ncu1.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/c10_plain.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_pair_conv|k_dw_tc|k_seg_' -c 40 -o gpurun_out/r2_full python tools/one_step.py > gpurun_out/c10_ncu2.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/c10_ncu2.log; ls -la gpurun_out/r2_full.ncu-rep
python tools/step_ops.py rmat > gpurun_out/c10_step_ops.txt 2>&1; tail -45 gpurun_out/c10_step_ops.txt
python bench.py --hidden 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_h128.json 2>gpurun_out/c10_b128.err; echo "h128 rc=$?"
python bench.py --hidden 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_h32.json 2>gpurun_out/c10_b32.err; echo "h32 rc=$?"
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
for f in ('c10_bench_h128','c10_bench_h32'):
    try:
        d=load('gpurun_out/%s.json'%f); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:(round(v['ms']/d['steps'],2), v['GBps']) for k,v in list(po.items())[:6]})
    except Exception as e: print(f,'ERR',e)
PY
