# (1) where the full-size error comes from  (2) one rank's compute of an 8-way row-sharded step, collectives skipped
timeout 600 python tools/diag_fullsize.py rmat 64 > gpurun_out/c18_diag.log 2>&1; echo "diag rc=$?"; grep -v Warning gpurun_out/c18_diag.log | tail -12
timeout 600 python tools/rowshard_dry.py 8 64 0,3,7 > gpurun_out/c18_dry.log 2>&1; echo "dry rc=$?"; tail -8 gpurun_out/c18_dry.log
