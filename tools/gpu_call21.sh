python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/c21_pytest.log | tail -5
DRY_LIST=1 timeout 600 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c21_dry32.log 2>&1; echo "dry rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c21_dry32.log
TWOWL_SEG_ROWS_CTAS=8 timeout 600 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c21_dry8.log 2>&1; echo "dry8 rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c21_dry8.log
timeout 600 python tools/step_ops.py rmat 64 > gpurun_out/c21_ops32.log 2>&1; grep "seg_reduce\|total" gpurun_out/c21_ops32.log
TWOWL_SEG_ROWS_CTAS=8 timeout 600 python tools/step_ops.py rmat 64 > gpurun_out/c21_ops8.log 2>&1; grep "seg_reduce\|total" gpurun_out/c21_ops8.log
timeout 600 python tools/diag_stages.py rmat 64 > gpurun_out/c21_stages.log 2>&1; echo "stages rc=$?"; grep -v Warn gpurun_out/c21_stages.log | tail -32
