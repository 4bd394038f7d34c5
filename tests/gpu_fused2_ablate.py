"""Developer script: times logistic_fused2_kernel at the cfg5 size (X[4 Mi, 512], S = 64) under the ablation
switches of BB_FUSED2_ABLATE (results are wrong when set; timing only).  One process per setting because the
library reads the variable once.   python tests/gpu_fused2_ablate.py [rows]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import bayesic_b200.stats as S
n, d, s = int(sys.argv[1]), 512, 64
g = torch.Generator(device='cuda').manual_seed(1)
X = torch.randn(n, d, device='cuda', generator=g)
W = torch.randn(s, d, device='cuda', generator=g) / d ** 0.5
y = (torch.rand(n, device='cuda', generator=g) < 0.5).float()
for _ in range(3): S.logistic_reparam_stats(X, y, W)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): S.logistic_reparam_stats(X, y, W)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('%%.3f ms  %%.0f GB/s' %% (ms, n * (4.0 * d + 4) / ms / 1e6))
''' % ROOT

NAMES = {1: 'no global loads', 2: 'no converter stores', 4: 'no W stream', 8: 'no Z MMAs', 16: 'no G MMAs',
         32: 'no epilogue math'}


def main():
    rows = sys.argv[1] if len(sys.argv) > 1 else str(1 << 22)
    for setting in (0, 1, 2, 3, 4, 7, 8, 16, 24, 32, 56, 63, 31):
        env = dict(os.environ, BB_FUSED2_ABLATE=str(setting))
        out = subprocess.run([sys.executable, '-c', CHILD, rows], env=env, capture_output=True, text=True, timeout=300)
        what = ' + '.join(NAMES[b] for b in sorted(NAMES) if setting & b) or 'full kernel'
        print('ablate %2d  %-70s %s' % (setting, what, (out.stdout.strip() or out.stderr.strip()[-300:])), flush=True)


if __name__ == '__main__':
    main()
