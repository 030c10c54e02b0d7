python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x > gpurun_out/c22_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/c22_pytest.log | tail -5
DRY_LIST=1 timeout 600 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c22_dry.log 2>&1; echo "dry rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c22_dry.log
DRY_LIST=1 TWOWL_SEG_NARROW=1 timeout 600 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c22_dry_narrow.log 2>&1; echo "dry narrow rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c22_dry_narrow.log
timeout 600 python tools/diag_node.py rmat 64 > gpurun_out/c22_node.log 2>&1; echo "node rc=$?"; grep -v Warn gpurun_out/c22_node.log | tail -12
