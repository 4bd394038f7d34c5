mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for spec in gram:gram_pair_kernel logistic:logistic_fused_kernel colproj:colproj_kernel weighted:weighted_pairs_kernel logits:mixture_logits_kernel rowproj:rowproj_kernel suffstats:suffstats_tc_kernel; do
  w=${spec%%:*}; k=${spec##*:}
  timeout 120 python $D $w > gpurun_out/plain_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_r1f_$w python $D $w > gpurun_out/ncu_$w.log 2>&1
  cat gpurun_out/plain_$w.log; tail -1 gpurun_out/ncu_$w.log
done
