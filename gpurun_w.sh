mkdir -p gpurun_out
BB_SUFFSTATS_IMPL=bf16 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "gaussian or cfg1 or cfg2 or host_streamed or compensated" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_w.log
for impl in tf32 bf16; do echo $impl; BB_SUFFSTATS_IMPL=$impl timeout 100 python tests/gpu_profile_driver.py suffstats; done
for pf in 0 1 4; do echo "bf16 prefetch $pf"; BB_SUFFSTATS_IMPL=bf16 BB_SUFFSTATS_PREFETCH=$pf timeout 100 python tests/gpu_profile_driver.py suffstats; done
