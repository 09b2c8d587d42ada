set -x
python tools/time_train.py 8 f32 > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_bwd_kernel|pw_bwd_tc|dwpw_tc_kernel' -s 60 -c 12 -o gpurun_out/r02f_prof_train python tools/time_train.py 8 f32 > gpurun_out/ncu_train.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
