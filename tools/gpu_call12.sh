python tools/one_step.py > gpurun_out/c12_plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k regex:'k_pair_conv<[12]|k_dw_tc' -c 4 -o gpurun_out/r2_full_tc2 python tools/one_step.py > gpurun_out/c12_ncu.log 2>&1
echo "full tc2 rc=$?"; ls -la gpurun_out/*.ncu-rep
python -m pytest tests -m gpu -q -x > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/c12_pytest.log | tail -6
for h in 32 64 128; do python bench.py --hidden $h --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench_h$h.json 2>gpurun_out/c12_b$h.err; echo "h$h rc=$?"; done
python tools/sweep_table.py gpurun_out/r2_model_sweep_rmat20.txt gpurun_out/c12_bench_h32.json gpurun_out/c12_bench_h64.json gpurun_out/c12_bench_h128.json | cut -c1-250
