set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/quick_time.py --bf16 c3 c5 2>&1 | grep "loss"
timeout 300 python tools/quick_time.py c3 2>&1 | grep "loss"
timeout 300 python tools/joint_time.py 2>&1 | tail -6
