set -x
mkdir -p /tmp/prof
P="--set full --import-source on --clock-control none -f"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file /tmp/prof/launches.csv python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline --no-c5 > gpurun_out/bench_under_ncu_r2.log 2>&1
python tools/summarize_profiles.py launches /tmp/prof/launches.csv gpurun_out/r2_launches_bench_c3.md > /dev/null
timeout 600 ncu $P -k regex:"rowstats|lattice|grad_" -s 3 -c 3 -o /tmp/prof/c3 python tools/profile_target.py c3 2 > /tmp/prof/c3.log 2>&1
python tools/summarize_profiles.py full /tmp/prof/c3.ncu-rep gpurun_out/r2_c3_full.md > /dev/null
RNNT_B200_GROUPS=1 timeout 600 ncu $P -k regex:"rowstats|lattice|grad_" -s 3 -c 3 -o /tmp/prof/c4 python tools/profile_target.py c4 2 > /tmp/prof/c4.log 2>&1
python tools/summarize_profiles.py full /tmp/prof/c4.ncu-rep gpurun_out/r2_c4_full.md > /dev/null
timeout 600 ncu $P -k regex:"rowstats|lattice|grad_" -s 3 -c 3 -o /tmp/prof/c2 python tools/profile_target.py c2 2 > /tmp/prof/c2.log 2>&1
python tools/summarize_profiles.py full /tmp/prof/c2.ncu-rep gpurun_out/r2_c2_full.md > /dev/null
timeout 600 ncu $P -k regex:"rowstats|lattice|grad_" -s 3 -c 3 -o /tmp/prof/c3bf python tools/profile_target.py c3 2 bf16 > /tmp/prof/c3bf.log 2>&1
python tools/summarize_profiles.py full /tmp/prof/c3bf.ncu-rep gpurun_out/r2_c3_bf16_full.md > /dev/null
timeout 600 ncu $P -k regex:"gemm_kernel|grad_fused|joint_prep" -s 4 -c 4 -o /tmp/prof/joint python tools/joint_profile_target.py > /tmp/prof/joint.log 2>&1
python tools/summarize_profiles.py full /tmp/prof/joint.ncu-rep gpurun_out/r2_joint_full.md > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file /tmp/prof/lj.csv python tools/joint_profile_target.py > /dev/null 2>&1
python tools/summarize_profiles.py launches /tmp/prof/lj.csv gpurun_out/r2_launches_add_joint_c3.md > /dev/null
ls -la gpurun_out/*.md
