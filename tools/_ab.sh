cd /root/repo
python tools/profile_hbm.py assign 64 > gpurun_out/profile_hbm_assign.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_assign -c 2 -o gpurun_out/assign_r2 -f python tools/profile_hbm.py assign 64 > gpurun_out/ncu_assign.log 2>&1
tail -3 gpurun_out/ncu_assign.log; ls -la gpurun_out/assign_r2.ncu-rep
