set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/unprofiled_time.py
timeout 300 python tools/quick_time.py c2 c4 2>&1 | grep "loss"
