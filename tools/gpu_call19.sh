DIAG_PATHS=1 timeout 600 python tools/diag_fullsize.py rmat 64 > gpurun_out/c19_diag.log 2>&1; echo "diag rc=$?"; grep -v Warning gpurun_out/c19_diag.log | tail -8
DRY_LIST=1 timeout 600 python tools/rowshard_dry.py 8 64 0,7 > gpurun_out/c19_dry.log 2>&1; echo "dry rc=$?"; tail -4 gpurun_out/c19_dry.log
