set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
for nt in 256 128 64; do RNNT_B200_CHUNK_NT=$nt timeout 300 python tools/unprofiled_time.py 2>&1 | grep -v "c3"; done
RNNT_B200_CHUNK_NT=128 RNNT_B200_GROUPS=1 timeout 300 python tools/quick_time.py c4 c2 2>&1 | grep "loss+grad"
RNNT_B200_CHUNK_NT=64 RNNT_B200_GROUPS=1 timeout 300 python tools/quick_time.py c4 c2 2>&1 | grep "loss+grad"
