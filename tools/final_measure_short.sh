# Round-end measurement session without the --set full capture (one gpurun call, one GPU).  Outputs under gpurun_out/.
set -x
T=${1:-r02z}
L3D_DEBUG_POISON=1 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests_poisoned.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_gpu_tests_poisoned.log
tail -3 gpurun_out/${T}_gpu_tests_poisoned.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --dtype f32 --skip-cpu --quick > gpurun_out/${T}_bench_f32.json 2>> gpurun_out/${T}.err
tail -3 gpurun_out/${T}.err
export L3D_INFER_GRAPH=0
python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 90 --csv --log-file gpurun_out/${T}_launches_ncu.csv python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
