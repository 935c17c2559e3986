# Dev helper: compute-sanitizer passes over tools/sanitize_target.py (results summarised in profiles/r2_sanitizer.txt)
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck_r2.log 2>&1; echo memcheck rc=$?
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/sanitizer_synccheck_r2.log 2>&1; tail -1 gpurun_out/sanitizer_synccheck_r2.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck_r2.log 2>&1
grep "RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck_r2.log; grep -o "in [a-z_]*\.cuh:[0-9]*" gpurun_out/sanitizer_racecheck_r2.log | sort | uniq -c | sort -rn
# initcheck: every report must be one of the speculative lattice loads of row_grad_setup_spec (count by function)
timeout 1500 compute-sanitizer --tool initcheck --print-limit 2000000 python tools/sanitize_target.py 2>&1 | grep "=========     at " | sed 's/(.*//; s/.* at //' | sort | uniq -c | sort -rn > gpurun_out/sanitizer_initcheck_r2_by_function.txt
cat gpurun_out/sanitizer_initcheck_r2_by_function.txt
