set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
mkdir -p gpurun_out
timeout 180 python - <<'PY' > gpurun_out/tiny.log 2>&1
import torch, numpy as np, sys
sys.path.insert(0, '.')
import bayesic_b200.stats as S
X = (np.random.RandomState(0).randn(256, 64)).astype(np.float32)
n, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
torch.cuda.synchronize()
r2 = X.astype('f8').T @ X.astype('f8')
print('tiny ok, relerr', np.abs(s2.cpu().numpy() - r2).max() / np.abs(r2).max())
print('s1 relerr', np.abs(s1.cpu().numpy() - X.astype('f8').sum(0)).max())
PY
echo "tiny exit $?"; cat gpurun_out/tiny.log
timeout 600 python tests/gpu_first_light.py > gpurun_out/first_light.log 2>&1; echo "first_light exit $?"; cat gpurun_out/first_light.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
