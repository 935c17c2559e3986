set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/unprofiled_time.py
for R in 8 32; do RNNT_B200_LAT_RING=$R timeout 200 python tools/unprofiled_time.py c4; done
for G in 1 2 3 6 8; do RNNT_B200_GROUPS=$G timeout 200 python tools/unprofiled_time.py c4; done
RNNT_B200_GROUPS=1 timeout 300 python tools/quick_time.py c4 2>&1 | grep "loss+grad"
RNNT_B200_TIMELINE=1 timeout 200 python tools/profile_target.py c4 3 2>&1 | tail -22
