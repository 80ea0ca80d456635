#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, per rank and in aggregate: what bounds the `e2e` number of bench.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py   (or plain python: N = 1)

Every rank copies pinned host buffers to its GPU and back with cudaMemcpyAsync on two streams -- uploads alone, downloads
alone, and both at once in the 3:1 byte ratio of an embed + extract step (75 MB up, 25 MB down per 4K frame) -- all ranks
at the same time, timed on the device, max over ranks.  Prints one JSON line (rank 0): GB/s per rank and aggregate, and the
e2e Mpix/s those rates allow for the C2 step (3 x 24.9 MB up, 1 x 24.9 MB down per 8.29 Mpix frame)."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    fb = 3840 * 2160 * 3
    nup, ndn = 3 * 8, 8                       # 8 steps' worth of frames per measurement
    h_up = torch.empty((nup, fb), dtype=torch.uint8).pin_memory()
    h_dn = torch.empty((ndn, fb), dtype=torch.uint8).pin_memory()
    d_up = torch.empty((nup, fb), dtype=torch.uint8, device='cuda')
    d_dn = torch.empty((ndn, fb), dtype=torch.uint8, device='cuda')
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up, dn, reps=5):
        best = 1e30
        for _ in range(reps):
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            s_up.wait_event(e0); s_dn.wait_event(e0)
            if up:
                with torch.cuda.stream(s_up):
                    for i in range(nup):
                        d_up[i].copy_(h_up[i], non_blocking=True)
                    e1.record(s_up)
            if dn:
                with torch.cuda.stream(s_dn):
                    for i in range(ndn):
                        h_dn[i].copy_(d_dn[i], non_blocking=True)
                    e2.record(s_dn)
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e1) if up else 0.0, e0.elapsed_time(e2) if dn else 0.0)
            if world > 1:
                t = torch.tensor([ms], device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = min(best, ms)
        return best

    ms_up, ms_dn, ms_both = run(True, False), run(False, True), run(True, True)
    gb = lambda n, ms: n * fb / (ms * 1e-3) / 1e9
    if rank == 0:
        steps = 8
        line = {
            'n_gpus': world,
            'h2d_gbs_per_rank': round(gb(nup, ms_up), 1), 'd2h_gbs_per_rank': round(gb(ndn, ms_dn), 1),
            'duplex_gbs_per_rank': {'h2d': round(gb(nup, ms_both), 1), 'd2h': round(gb(ndn, ms_both), 1)},
            'aggregate_duplex_gbs': round(world * gb(nup + ndn, ms_both), 1),
            'c2_step_floor_ms': round(ms_both / steps, 3),
            'c2_e2e_ceiling_mpix_s': round(world * steps * 3840 * 2160 / (ms_both * 1e-3) / 1e6, 0),
            'note': 'all ranks copy at the same time; pinned buffers; times are max over ranks, best of 5; the ceiling is what the '
                    'copies alone allow for the embed + extract step of one 4K frame per rank (75 MB up + 25 MB down)',
            'cpu_affinity': sorted(os.sched_getaffinity(0))[:4] + ['...', len(os.sched_getaffinity(0))],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
