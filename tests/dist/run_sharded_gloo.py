"""world_size-N gloo worker: the sharded orchestration (spread_spectrum_watermarking_b200.sharded) with the
numpy stand-in ops, checked against the single-process oracle on the same frame.  Launched by
tests/test_sharded_gloo.py through torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.dirname(os.path.abspath(__file__))):
    sys.path.insert(0, p)

import ssw_oracle as so  # noqa: E402
from oracle_ops import OracleOps  # noqa: E402
from spread_spectrum_watermarking_b200 import sharded  # noqa: E402
from spread_spectrum_watermarking_b200._lib import ssw_config  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    w, h, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    ordering = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    frame = so.synth_frame(w, h, seed=21)
    rng = np.random.default_rng(3)
    marks = [rng.standard_normal(k).astype(np.float32) for _ in range(2)]
    plan = sharded.ShardPlan(w, h, world, rank)
    rows = torch.from_numpy(frame[plan.row0:plan.row0 + plan.hb].copy())
    cfg = ssw_config(2, 0.1, ordering)
    ops = OracleOps()

    # forward: every rank's transposed columns against the oracle's coefficient plane
    wr = sharded.ShardedWriter(rows, w, h, cfg, ops)
    ref_c, _, _ = so.forward(frame)
    mine = wr.frame.coeff.numpy()                                   # [wb][H]
    ref_cols = ref_c[:, plan.col0:plan.col0 + plan.wb].T
    assert np.abs(mine - ref_cols).max() <= 2e-6 * np.abs(ref_c).max(), 'coefficients'
    # ordered indices identical on every rank and equal to the oracle's full stable sort
    wr.embed(marks)
    idx = wr.indices.numpy().astype(np.int64)
    full = np.zeros((h, w), np.float32)
    gathered = [torch.empty_like(torch.from_numpy(mine)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(np.ascontiguousarray(mine)))
    ref_idx = so.obtain_indices(ref_c.astype(np.float32).ravel(), ordering, w, h, k=k)
    assert (idx == ref_idx).all(), 'ordered top-k'
    # owners: every index is modified on exactly one rank
    owners = np.array([plan.owner_of(int(p)) for p in idx])
    assert ((owners >= 0) & (owners < world)).all()
    # embed + inverse: this rank's rows of the watermarked image
    out = wr.result_rgb8().numpy()
    ref_img, _, _ = so.embed(frame, marks, ordering=ordering)
    d = np.abs(out.astype(int) - ref_img[plan.row0:plan.row0 + plan.hb].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 5e-3, 'watermarked rows (%d, %g)' % (d.max(), (d > 0).mean())
    # extract from the oracle's watermarked image: same vector on every rank, equal to the oracle's
    rd = sharded.ShardedReader(rows, w, h, cfg, ops)
    ext = rd.extract(torch.from_numpy(ref_img[plan.row0:plan.row0 + plan.hb].copy()), k).numpy()
    ref_ext, _ = so.extract(frame, ref_img, k, ordering=ordering)
    assert np.abs(ext - ref_ext).max() < 2e-3, 'extracted'
    both = [torch.empty(k) for _ in range(world)]
    dist.all_gather(both, torch.from_numpy(ext))
    assert all((b.numpy() == ext).all() for b in both), 'extracted vector differs between ranks'
    assert so.similarity(ext, marks[0]) > 6
    dist.barrier()
    if rank == 0:
        print('SHARDED_GLOO_OK world=%d %dx%d k=%d ordering=%d' % (world, w, h, k, ordering))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
