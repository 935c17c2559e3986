set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 ./tools/probe/umma_probe > gpurun_out/umma_probe_r2.log 2>&1; grep -c OK gpurun_out/umma_probe_r2.log; grep -c MISMATCH gpurun_out/umma_probe_r2.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo bench rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
echo ref rc=$?; tail -c 400 gpurun_out/bench_r2_ref.json
