python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2>gpurun_out/c13_b.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/c13_bench.json').read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1]); po=d['roofline']['per_op']
print(round(d['ms_per_step'],2), {k:(round(v['ms']/d['steps'],2), v['GBps'], v.get('frac_dram')) for k,v in list(po.items())[:6]})
PY
python tools/one_step.py > gpurun_out/c13_plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:'k_pair_conv<[12]' -c 3 -o gpurun_out/r2_full_pc python tools/one_step.py > gpurun_out/c13_ncu1.log 2>&1
echo "full pc rc=$?"; tail -2 gpurun_out/c13_ncu1.log
ncu --set full --clock-control none --profile-from-start off -k regex:'k_seg_rows|k_seg_chunks|k_seg_long' -c 24 -o gpurun_out/r2_full_seg python tools/one_step.py > gpurun_out/c13_ncu2.log 2>&1
echo "full seg rc=$?"; ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
