"""Developer script: phase timeline of the one-launch Gaussian pass (variant build with
-DBB_SUFFSTATS_TIMELINE, selected through BB_LIB_PATH)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesic_b200.parallel import GaussianPass
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
dev = torch.device('cuda')
X = torch.randn(rows, 64, device=dev)
e1 = torch.eye(64, dtype=torch.float64, device=dev); e2 = torch.ones(64, dtype=torch.float64, device=dev)
p = GaussianPass(64, dev)
for _ in range(6):
    p.run(X, e1, e2, 0.1, 0.2)
torch.cuda.synchronize()
