mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs"
timeout 300 $B > gpurun_out/launches_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_ncu.log 2>&1
tail -2 gpurun_out/launches_plain.log | cut -c1-200; wc -l gpurun_out/launches.csv
bash scripts/gpurun_prof_one.sh logistic logistic_fused2_kernel r1f_logistic
