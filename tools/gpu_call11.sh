python tools/one_step.py > gpurun_out/c11_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/one_step.py > gpurun_out/c11_ncu1.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/c11_plain.log
ncu --set full --clock-control none --profile-from-start off -k regex:'k_pair_conv|k_dw_tc' -c 4 -o gpurun_out/r2_full_tc python tools/one_step.py > gpurun_out/c11_ncu2.log 2>&1
echo "full tc rc=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:'k_seg_rows|k_seg_chunks|k_seg_long|k_pair_init' -c 14 -o gpurun_out/r2_full_seg python tools/one_step.py > gpurun_out/c11_ncu3.log 2>&1
echo "full seg rc=$?"; ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
