"""Multi-GPU check of the sharded frame: torchrun --nproc-per-node N tests/dist/run_sharded_nccl.py W H K
Every rank shards a synthetic frame by rows, runs embed + extract through NCCL, and rank 0 compares the
gathered result with the unsharded path on its own GPU (RGB8 within 1 LSB, same ordered indices up to
near-ties, extracted vector close, mark detected)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)

import spread_spectrum_watermarking_b200 as wm  # noqa: E402
from spread_spectrum_watermarking_b200 import sharded  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    w, h, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    ops = sharded.CudaOps(local)
    ctx = ops.ctx
    plan = sharded.ShardPlan(w, h, world, rank)
    rows = torch.empty((plan.hb, w, 3), dtype=torch.uint8, device='cuda')
    # rows of the synthetic frame: generate the whole frame's rows on this GPU (cheap) and slice
    full = torch.empty((h, w, 3), dtype=torch.uint8, device='cuda')
    with ops.scope():
        wm._lib.check(wm.lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 4, 0, 1, full.data_ptr()))
        rows.copy_(full[plan.row0:plan.row0 + plan.hb])
    mark = np.random.default_rng(7).standard_normal(k).astype(np.float32)
    cfg = wm._lib.ssw_config(2, 0.1, 0)
    wr = sharded.ShardedWriter(rows, w, h, cfg, ops)
    wr.embed([mark])
    out_rows = wr.result_rgb8()
    with ops.scope():
        idx = wr.indices.cpu().numpy().astype(np.int64)
        parts = [torch.empty_like(out_rows) for _ in range(world)]
        dist.all_gather(parts, out_rows)
        out = torch.cat(parts)
    rd = sharded.ShardedReader(rows, w, h, cfg, ops)
    ext_t = rd.extract(out_rows, k)
    with ops.scope():
        ext = ext_t.cpu().numpy()
        frame_host = full.cpu().numpy()
        out_host = out.cpu().numpy()
    ops.synchronize()
    sim = float(wm.Tester.new(ext, ctx=ctx).similarity(mark).similarity)
    ok = True
    if rank == 0:
        frame = frame_host
        if w * h <= 16384 * 16384 and max(w, h) <= 16384:
            plain = wm.Writer.new(frame, ctx=ctx)
            ref_idx = plain.indices(k).astype(np.int64)
            ref = plain.mark_rgb8([mark])
            d = np.abs(out_host.astype(int) - ref.astype(int))
            same = float((idx == ref_idx).mean())
            print('sharded vs unsharded: rgb8 max |d| %d, differing %.5f%%, identical ranks %.4f, set equal %s'
                  % (d.max(), 100.0 * (d > 0).mean(), same, set(idx.tolist()) == set(ref_idx.tolist())))
            ok = d.max() <= 1 and (d > 0).mean() < 0.02 and same > 0.95
        print('similarity %.3f (world %d, %dx%d)' % (sim, world, w, h))
        ok = ok and sim > 6
        print('SHARDED_NCCL_OK' if ok else 'SHARDED_NCCL_FAILED')
    dist.barrier()
    del wr, rd
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
