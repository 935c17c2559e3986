set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "== default groups"; timeout 300 python tools/quick_time.py c2 c3 c4
echo "== GROUPS=1"; RNNT_B200_GROUPS=1 timeout 300 python tools/quick_time.py c4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json | head -c 6000
