# Round-end measurement session (one gpurun call, one GPU).  Outputs under gpurun_out/, copied to profiles/ by hand.
set -x
L3D_DEBUG_POISON=1 python -m pytest tests -m gpu -q > gpurun_out/r02v_gpu_tests_poisoned.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_gpu_tests_poisoned.log
tail -3 gpurun_out/r02v_gpu_tests_poisoned.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v_bench_reference.json 2>> gpurun_out/r02v.err
python bench.py --steps 5 --dtype f32 --skip-cpu --quick > gpurun_out/r02v_bench_f32.json 2>> gpurun_out/r02v.err
python bench.py --steps 5 --skip-cpu --skip-train --sweep --eager-gpu > gpurun_out/r02v_bench_sweep.json 2>> gpurun_out/r02v.err
tail -3 gpurun_out/r02v.err
export L3D_INFER_GRAPH=0
python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 90 --csv --log-file gpurun_out/r02v_launches_ncu.csv python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3_tc -s 21 -c 7 -o gpurun_out/r02v_prof_conv3_tc python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r02v_prof_conv3_tc.ncu-rep.tmp
python tools/ncu_summary.py gpurun_out/r02v_prof_conv3_tc.ncu-rep > gpurun_out/r02v_prof_conv3_tc_summary.txt 2>&1
python tools/time_merge.py > gpurun_out/r02v_time_merge.txt 2>&1
python bench.py --steps 5 --variant dense --skip-cpu --skip-train > gpurun_out/r02v_bench_dense.json 2>> gpurun_out/r02v.err
python bench.py --steps 5 --variant grouped --skip-cpu --skip-train > gpurun_out/r02v_bench_grouped.json 2>> gpurun_out/r02v.err
tail -2 gpurun_out/r02v.err
