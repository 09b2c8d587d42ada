tool=$1; shift
compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/r02_sanitizer_${tool}.txt 2>&1
echo "rc=$?" >> gpurun_out/r02_sanitizer_${tool}.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|Error|error" gpurun_out/r02_sanitizer_${tool}.txt | head -20
