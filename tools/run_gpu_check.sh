set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python tools/unprofiled_time.py
RNNT_B200_LATTICE2=0 timeout 300 python tools/unprofiled_time.py 2>&1 | grep c2
timeout 300 python tools/quick_time.py c2 2>&1 | grep "loss+grad"
