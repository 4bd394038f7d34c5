"""Developer script (torchrun): the box's host -> device ingress ceiling.  Every rank copies a pinned buffer to its
GPU with plain cudaMemcpyAsync (torch .copy_(non_blocking=True)); reports the per-rank and aggregate GB/s with all
ranks copying at once and with rank 0 copying alone -- the ceiling `bb_suffstats_gaussian_host` (and bench.py's
`e2e`) can reach at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tests/gpu_h2d_probe.py [MiB per rank]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    host = torch.empty(mib << 20, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    devbuf = torch.empty_like(host, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(active, reps=10, chunk=None):
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                if chunk is None:
                    devbuf.copy_(host, non_blocking=True)
                else:
                    for lo in range(0, host.numel(), chunk):
                        devbuf[lo:lo + chunk].copy_(host[lo:lo + chunk], non_blocking=True)
            torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        barrier()
        total = time.perf_counter() - t0
        gbs = (mib << 20) * reps / mine / 1e9 if active else 0.0
        if world > 1:
            t = torch.tensor([gbs, total], dtype=torch.float64, device=dev)
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            return [float(a[0]) for a in allv], max(float(a[1]) for a in allv)
        return [gbs], total

    out = {'world': world, 'mib_per_rank': mib}
    run(True, reps=2)
    per_rank, total = run(True)
    out['all_ranks'] = {'per_rank_gbs': [round(v, 1) for v in per_rank],
                        'aggregate_gbs': round(world * (mib << 20) * 10 / total / 1e9, 1)}
    per_rank, total = run(True, chunk=1 << 28)
    out['all_ranks_256MiB_chunks'] = {'aggregate_gbs': round(world * (mib << 20) * 10 / total / 1e9, 1)}
    per_rank, total = run(rank == 0)
    out['rank0_alone_gbs'] = round(per_rank[0], 1)
    try:
        import psutil
        out['host'] = {'cpus': os.cpu_count(), 'mem_gib': round(psutil.virtual_memory().total / 2 ** 30)}
        numa = [d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')]
        out['host']['numa_nodes'] = len(numa)
    except Exception:
        pass
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
