"""Developer driver: one weighted-statistics call at cfg3 extents (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.stats as S  # noqa: E402

n, d, k = 1 << 19, 64, 256
X = torch.randn(n, d, device='cuda')
R = torch.softmax(torch.randn(n, k, device='cuda') * 2, 1)
for _ in range(3):
    S.weighted_suffstats(X, R)
torch.cuda.synchronize()
print('ok')
