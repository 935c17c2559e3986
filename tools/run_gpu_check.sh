set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 300 python tools/unprofiled_time.py
RNNT_B200_GROUPS=1 timeout 300 python tools/quick_time.py c4 c2 c3 2>&1 | grep "loss+grad"
