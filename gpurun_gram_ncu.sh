mkdir -p gpurun_out
T=tests/cuda/gram_test
timeout 60 $T 1048576 1024 1 > gpurun_out/gram_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_pair -s 3 -c 1 -o gpurun_out/prof_gram_r1b $T 1048576 1024 1 > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
