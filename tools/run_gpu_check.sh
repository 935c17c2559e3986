set -x
timeout 120 ./tools/probe/umma_probe
timeout 600 python -m pytest tests/test_gpu_add_joint.py -q 2>&1 | tail -4
timeout 300 python tools/joint_time.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_joint_r2.csv python tools/joint_profile_target.py > /dev/null 2>&1
