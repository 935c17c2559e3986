set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
timeout 300 python tools/quick_time.py c2 c3 c4
timeout 300 python tools/joint_time.py 2>&1 | head -3
