set -x
timeout 600 python -m pytest tests/test_gpu_add_joint.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/joint_time.py 2>&1 | tail -4
timeout 200 ./tools/probe/umma_probe > gpurun_out/umma_probe_r2.log 2>&1; tail -12 gpurun_out/umma_probe_r2.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file /tmp/lj.csv python tools/joint_profile_target.py > /dev/null 2>&1
grep -i "grad_fused\|gemm_kernel" /tmp/lj.csv | cut -d, -f5,12- | cut -c1-200 | tail -4
