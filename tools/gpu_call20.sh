# balanced two-range blocks: parity tests of the row-sharded path, dry per-rank compute, ncu of rank 0's seg / readout launches
python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/c20_pytest.log | tail -5
DRY_LIST=1 timeout 600 python tools/rowshard_dry.py 8 64 0,7 > gpurun_out/c20_dry.log 2>&1; echo "dry rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c20_dry.log
NCU=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_seg|k_gn2' -o gpurun_out/c20_seg_rank0 -f python tools/rowshard_dry.py 8 64 0 > gpurun_out/c20_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c20_ncu.log
ls -la gpurun_out/*.ncu-rep
