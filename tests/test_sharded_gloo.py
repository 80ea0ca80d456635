"""The N>1 path of the sharded frame (host orchestration: partition, all-to-all transposes, distributed
top-k merge, owner-side embed / extract) on CPU: world_size 2 and 4 over gloo, numpy stand-in kernels,
checked against the single-process oracle.  The CUDA kernels of the same steps are covered by
tests/test_gpu_sharded.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, 'tests', 'dist', 'run_sharded_gloo.py')


def _run(world, args, port, min_chunk_lines=None):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port), WORKER] + [str(a) for a in args]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    if min_chunk_lines:
        env['SSW_SHARD_MIN_CHUNK_LINES'] = str(min_chunk_lines)   # force the chunked (overlapped) exchange
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and 'SHARDED_GLOO_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_shard_plan_partition():
    sys.path.insert(0, ROOT)
    from spread_spectrum_watermarking_b200.sharded import ShardPlan
    from spread_spectrum_watermarking_b200 import SswError
    plans = [ShardPlan(96, 64, 4, r) for r in range(4)]
    assert [p.row0 for p in plans] == [0, 16, 32, 48] and [p.col0 for p in plans] == [0, 24, 48, 72]
    p = plans[2]
    assert p.owner_of(5 * 96 + 50) == 2 and p.local_position(5 * 96 + 50) == (50 - 48) * 64 + 5
    with pytest.raises(SswError):
        ShardPlan(97, 64, 4, 0)


@pytest.mark.parametrize('world,w,h,k,ordering,port', [(2, 96, 64, 200, 0, 29621), (4, 128, 96, 300, 0, 29622),
                                                        (2, 64, 48, 100, 1, 29623)])
def test_sharded_frame_over_gloo(world, w, h, k, ordering, port):
    _run(world, [w, h, k, ordering], port)


def test_sharded_frame_chunked_exchange_over_gloo():
    """4 slices per pass, each with its own asynchronous all-to-all"""
    _run(2, [128, 64, 200, 0], 29624, min_chunk_lines=4)
