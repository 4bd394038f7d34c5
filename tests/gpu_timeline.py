"""Developer script: phase timeline of the one-launch Gaussian pass (variant build with
-DBB_SUFFSTATS_TIMELINE, selected through BB_LIB_PATH).  Under torchrun every rank runs its shard and rank 0's
device printf shows the exchange phases.   python tests/gpu_timeline.py [rows per rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bayesic_b200.parallel import GaussianPass
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
X = torch.randn(rows, 64, device=dev)
e1 = torch.eye(64, dtype=torch.float64, device=dev); e2 = torch.ones(64, dtype=torch.float64, device=dev)
p = GaussianPass(64, dev)
if world > 1:
    dist.barrier()
for _ in range(6):
    p.run(X, e1, e2, 0.1, 0.2)
torch.cuda.synchronize()
p.check()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
