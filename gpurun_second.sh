mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
cat > /tmp/prof.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bayesic_b200.stats as S
n, d = 1 << 24, 64
X = torch.randn(n, d, device='cuda')
out = (torch.empty(d, dtype=torch.float64, device='cuda'), torch.empty((d, d), dtype=torch.float64, device='cuda'))
for _ in range(3):
    S.gaussian_suffstats(X, out=out)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/prof.py && timeout 600 ncu --set full --clock-control none --import-source on -k regex:suffstats_tc_kernel -s 1 -c 1 -o gpurun_out/prof_suffstats_r1a python /tmp/prof.py > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"; tail -5 gpurun_out/ncu.log
