D=tests/gpu_profile_driver.py
mkdir -p gpurun_out
for b in 0 1; do
  BB_LIB_PATH=$PWD/bayesic_b200/lib/libbb_gram_timeline.so BB_GRAM_BALANCE=$b timeout 120 python $D gram > gpurun_out/gram_timeline_b$b.log 2>&1
  tail -1 gpurun_out/gram_timeline_b$b.log
done
for b in 0 1; do echo -n "BB_GRAM_BALANCE=$b  "; BB_GRAM_BALANCE=$b timeout 120 python $D gram 2>&1 | tail -1; done
