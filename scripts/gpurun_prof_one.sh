# usage: gpurun -- 'bash scripts/gpurun_prof_one.sh <driver-case> <kernel-regex> <tag> [ENV=VAL ...]'
mkdir -p gpurun_out
w=$1; k=$2; tag=$3; shift 3
for kv in "$@"; do export "$kv"; done
D=tests/gpu_profile_driver.py
timeout 120 python $D $w > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_$tag python $D $w > gpurun_out/ncu_$tag.log 2>&1
cat gpurun_out/plain_$tag.log; tail -1 gpurun_out/ncu_$tag.log
