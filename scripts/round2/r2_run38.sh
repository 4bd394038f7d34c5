mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for lib in "" "$PWD/bayesic_b200/lib/libbayesic_b200_chain4.so"; do
  export BB_LIB_PATH=$lib; [ -z "$lib" ] && unset BB_LIB_PATH
  echo "== lib: ${lib:-default}"
  for w in gram weighted_split logistic; do timeout 120 python $D $w 2>&1 | tail -1; done
done
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_chain4.so
timeout 900 python tests/gpu_parity_report.py 2>&1 | grep "cfg[345]" | grep -v global | cut -c1-150
