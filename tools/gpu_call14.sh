python tools/one_step.py > gpurun_out/c14_plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:'k_pair_conv<\(int\)[12]' -c 3 -o gpurun_out/r2_full_pc python tools/one_step.py > gpurun_out/c14_ncu1.log 2>&1
echo "full pc rc=$?"; tail -2 gpurun_out/c14_ncu1.log | cut -c1-200
ncu --set full --clock-control none --profile-from-start off -k regex:'k_seg_rows|k_seg_chunks|k_seg_long' -c 24 -o gpurun_out/r2_full_seg python tools/one_step.py > gpurun_out/c14_ncu2.log 2>&1
echo "full seg rc=$?"; ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
