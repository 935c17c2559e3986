set -x
for P in 0 60000; do RNNT_B200_FUSED_PAD_SMEM=$P timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file /tmp/lj.csv python tools/joint_profile_target.py > /dev/null 2>&1; grep -i "grad_fused" /tmp/lj.csv | cut -d, -f5,12- | cut -c1-160 | tail -1; done
