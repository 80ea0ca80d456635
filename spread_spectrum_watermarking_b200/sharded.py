"""Row-sharded single frames across the GPUs of one node (BASELINE.json configs[3]).

The reference holds one whole frame on one core (`Writer::new`, /root/reference/src/algorithm.rs:295-316);
a gigapixel frame does not need to fit one GPU here.  One process per GPU (`torch.distributed`, NCCL over
NVLink / NVSwitch); rank g owns the pixel rows [g*H/G, (g+1)*H/G) and, after the first exchange, the
coefficient columns [g*W/G, (g+1)*W/G), kept TRANSPOSED (local plane [W/G][H]) so that both passes of
the separable DCT run over contiguous lines:

    forward   rows: RGB8 -> Y -> DCT-II along x          (local, ssw_lines_forward_dev)
              all-to-all of (H/G x W/G) blocks, each transposed on the way  <- the one exchange per transform
              cols: DCT-II along y over contiguous lines  (local)
    top-k     local lower bound of the k-th key -> max over ranks -> local candidates -> all-gather -> merge
    embed     every rank modulates the coefficients it owns (index lists hold the reference's flat indices)
    inverse   cols: DCT-III -> all-to-all back -> rows: DCT-III, x4/(W*H), YIQ->RGB8
    extract   base + derived forward, top-k on the base, owners gather their values, sum over ranks

This module is host plumbing: buffers are torch tensors, the exchanges are torch.distributed collectives,
every arithmetic step is a libssw kernel (`CudaOps`).  The orchestration is written against a small `ops`
interface so that the partition / index / merge logic is covered on CPU by world_size-2 gloo tests with a
numpy stand-in (tests/test_sharded_gloo.py) -- that stand-in lives in tests/, never in this package.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import SswError, check, lib, ssw_config, ssw_shard

TOPK_CAP = 8192
PIX_RGB8, PIX_RGB32F, PIX_PLANE = 0, 1, 2


class ShardPlan:
    """which rows / columns a rank owns (equal blocks; W and H must be divisible by the world size)"""

    def __init__(self, width, height, world, rank):
        if width % world or height % world:
            raise SswError(_lib.SSW_ERR_UNSUPPORTED, 'sharded frames need width and height divisible by the number of ranks')
        if not 0 <= rank < world:
            raise SswError(_lib.SSW_ERR_INVALID, 'rank out of range')
        self.width, self.height, self.world, self.rank = int(width), int(height), int(world), int(rank)
        self.hb, self.wb = height // world, width // world
        self.row0, self.col0 = rank * self.hb, rank * self.wb

    def owner_of(self, p):
        """rank that owns flat coefficient index p = r*W + c"""
        return (p % self.width) // self.wb

    def local_position(self, p):
        """position of flat index p inside the owner's transposed plane [wb][H]"""
        r, c = divmod(p, self.width)
        return (c % self.wb) * self.height + r


def _pick_chunks(n_lines, world):
    """slices of the local pass whose all-to-all overlaps the next slice's kernels (1 when nothing is exchanged)"""
    import os
    if world == 1:
        return 1
    floor = int(os.environ.get('SSW_SHARD_MIN_CHUNK_LINES', '64'))
    for c in (4, 2):
        if n_lines % (2 * c) == 0 and n_lines // c >= floor:
            return c
    return 1


def _scope(ops):
    """stream scope of the ops object (CudaOps: its torch stream; CPU stand-ins: nothing)"""
    import contextlib
    return ops.scope() if hasattr(ops, 'scope') else contextlib.nullcontext()


def _all_to_all(send, group):
    """send[j] goes to rank j; returns recv with recv[g] = block sent by rank g"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return send
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def _all_to_all_async(recv, send, group):
    """start the exchange of one chunk; returns a handle whose wait() orders the current stream after it
    (None when there is nothing to exchange).  NCCL runs it on its own stream, so the kernels launched next
    on the ops stream overlap with the transfer."""
    return dist.all_to_all_single(recv, send, group=group, async_op=True)


def _all_gather(t, group):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t.unsqueeze(0)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, t.contiguous(), group=group)
    return torch.stack(parts)


def _all_reduce(t, op, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


class CudaOps:
    """the per-rank arithmetic steps: libssw kernels on torch CUDA tensors (no CPU path).

    libssw launches on ITS context's stream, torch (copies, NCCL collectives) on torch's current stream:
    both must be the same stream.  `CudaOps(device)` therefore owns a torch stream, builds the libssw
    context on it, and every step of the orchestration runs inside `scope()` (= torch.cuda.stream(...))."""

    def __init__(self, device=0, ctx=None, stream=None):
        from . import Context
        self.device = torch.device('cuda', int(device))
        self.stream = stream if stream is not None else torch.cuda.Stream(self.device)
        if ctx is not None and int(ctx.stream or 0) != int(self.stream.cuda_stream):
            raise SswError(_lib.SSW_ERR_INVALID, 'the libssw context must be bound to the CudaOps stream')
        self.ctx = ctx if ctx is not None else Context(int(device), stream=self.stream.cuda_stream)

    def scope(self):
        return torch.cuda.stream(self.stream)

    def synchronize(self):
        self.stream.synchronize()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def lines_forward(self, src, n, n_lines, src_type, out=None):
        plane = out if out is not None else self.empty((n_lines, n), torch.float32)
        check(lib.ssw_lines_forward_dev(self.ctx.handle, src_type, src.data_ptr(), n, n_lines, plane.data_ptr()))
        return plane

    def lines_inverse(self, plane, n, n_lines, scale, dst_type=PIX_PLANE, dst=None, src_type=PIX_PLANE, src=None):
        dst = plane if dst is None else dst
        check(lib.ssw_lines_inverse_dev(self.ctx.handle, plane.data_ptr(), n, n_lines, ctypes.c_float(scale), dst_type,
                                        dst.data_ptr(), src_type, src.data_ptr() if src is not None else None))
        return dst

    def transpose_blocks(self, src, rows, cols, ld, nblocks):
        """src: [rows][nblocks*cols] with leading dimension ld -> [nblocks][cols][rows], block j = src[:, j*cols:(j+1)*cols].T"""
        dst = self.empty((nblocks, cols, rows), torch.float32)
        check(lib.ssw_transpose_dev(self.ctx.handle, src.data_ptr(), rows, cols, ld, cols, dst.data_ptr(), rows, cols * rows, nblocks))
        return dst

    def interleave_blocks(self, recv):
        """recv [C][G][lines][seg] -> [lines][G*C*seg]: line l = concatenation over ranks g, chunks c of its segments"""
        c, g, lines, seg = recv.shape
        if g == 1 and c == 1:
            return recv.view(lines, seg)
        out = self.empty((lines, g, c, seg), torch.float32)
        out.copy_(recv.permute(2, 1, 0, 3))   # plain strided copy (data movement only)
        return out.view(lines, g * c * seg)

    def lines_forward_segmented(self, recv, n):
        """DCT-II of the lines held as all-to-all blocks [C][G][lines][seg], read in place when the kernels
        support it (power-of-two segments, planned length); otherwise interleave first"""
        c, g, lines, seg = recv.shape
        if c * g > 1:
            plane = self.empty((lines, n), torch.float32)
            rc = lib.ssw_lines_forward_seg_dev(self.ctx.handle, recv.data_ptr(), n, lines, seg, c, g, plane.data_ptr())
            if rc == _lib.SSW_OK:
                return plane
            if rc != _lib.SSW_ERR_UNSUPPORTED:
                check(rc)
        t = self.interleave_blocks(recv)
        return self.lines_forward(t, n, lines, PIX_PLANE, out=t)

    def lines_inverse_segmented(self, recv, n, scale, dst, pixels):
        """DCT-III + colour of the lines held as all-to-all blocks [C][G][lines][seg] -> RGB8 rows `dst`"""
        c, g, lines, seg = recv.shape
        if c * g > 1:
            rc = lib.ssw_lines_inverse_seg_dev(self.ctx.handle, recv.data_ptr(), n, lines, seg, c, g, ctypes.c_float(scale),
                                               PIX_RGB8, dst.data_ptr(), PIX_RGB8, pixels.data_ptr())
            if rc == _lib.SSW_OK:
                return dst
            if rc != _lib.SSW_ERR_UNSUPPORTED:
                check(rc)
        a = self.interleave_blocks(recv)
        return self.lines_inverse(a, n, lines, scale, PIX_RGB8, dst, PIX_RGB8, pixels)

    def topk_bin(self, plane, shard, ordering, k):
        b = self.empty((1,), torch.int32)
        check(lib.ssw_shard_topk_bin_dev(self.ctx.handle, plane.data_ptr(), ctypes.byref(shard), ordering, k, b.data_ptr()))
        return b

    def topk_collect(self, plane, shard, ordering, bin_t):
        cand = self.empty((TOPK_CAP,), torch.int64)
        cnt = self.empty((1,), torch.int32)
        check(lib.ssw_shard_topk_collect_dev(self.ctx.handle, plane.data_ptr(), ctypes.byref(shard), ordering,
                                             bin_t.data_ptr(), cand.data_ptr(), cnt.data_ptr()))
        return cand, cnt

    def topk_merge(self, lists, counts, k):
        idx = self.empty((k,), torch.int32)   # u32 flat indices (< 2^31 for every supported frame)
        ov = self.empty((1,), torch.int32)
        check(lib.ssw_shard_topk_merge_dev(self.ctx.handle, lists.data_ptr(), counts.data_ptr(), lists.shape[0], k,
                                           idx.data_ptr(), ov.data_ptr()))
        return idx, ov

    def embed(self, plane, shard, idx, marks, cfg):
        """marks: [n_marks][k] f32 tensor (zero padded), all of length k"""
        check(lib.ssw_shard_embed_dev(self.ctx.handle, plane.data_ptr(), ctypes.byref(shard), idx.data_ptr(), idx.numel(),
                                      marks.data_ptr(), marks.shape[1], marks.shape[0], None, ctypes.byref(cfg)))

    def extract(self, base, derived, shard, idx, n, cfg):
        out = self.empty((n,), torch.float32)
        check(lib.ssw_shard_extract_dev(self.ctx.handle, base.data_ptr(), derived.data_ptr(), ctypes.byref(shard),
                                        idx.data_ptr(), n, ctypes.byref(cfg), out.data_ptr()))
        return out

    def to_device(self, array, dtype):
        return torch.as_tensor(array, dtype=dtype).to(self.device)


class Sharded:
    """`ssw_sharded_*` (include/ssw.h): the whole sharded path behind the C ABI -- orchestration in libssw, the exchange
    between the two passes of the transform through peer-mapped memory (no all-to-all collective), NCCL (inside the
    library) only for the bootstrap, the barriers and the few hundred bytes of the distributed top-k.  This class is the
    thin binding the tests and bench.py use; the only thing it adds is the hand-over of rank 0's NCCL id to the other
    ranks (here: one torch.distributed broadcast -- a Rust host would use its own channel)."""

    def __init__(self, ctx, width, height, rank=0, world=1, group=None):
        self.ctx, self.width, self.height, self.rank, self.world = ctx, int(width), int(height), int(rank), int(world)
        ident = torch.zeros(128, dtype=torch.uint8)
        if world > 1:
            if rank == 0:
                check(lib.ssw_sharded_unique_id(ident.data_ptr()))
            dev = ident.cuda() if dist.get_backend(group) == 'nccl' else ident
            dist.broadcast(dev, 0, group=group)
            ident = dev.cpu()
        h = ctypes.c_void_p()
        check(lib.ssw_sharded_create(ctx.handle, ident.data_ptr(), self.rank, self.world, self.width, self.height, ctypes.byref(h)))
        self.handle = h
        self.hb, self.wb = self.height // self.world, self.width // self.world

    def embed_rgb8(self, rows, cfg, mark_dev, out=None):
        """rows: [H/G][W][3] uint8 CUDA tensor (this rank's rows); mark_dev: float32 CUDA tensor; stream-ordered on the
        context's stream, no host synchronisation"""
        out = torch.empty_like(rows) if out is None else out
        check(lib.ssw_sharded_embed_rgb8_dev(self.handle, rows.data_ptr(), ctypes.byref(cfg), mark_dev.data_ptr(), mark_dev.numel(),
                                             out.data_ptr()))
        return out

    def extract(self, base_rows, derived_rows, cfg, n, out=None):
        out = torch.empty((n,), dtype=torch.float32, device=base_rows.device) if out is None else out
        check(lib.ssw_sharded_extract_rgb8_dev(self.handle, base_rows.data_ptr(), derived_rows.data_ptr(), ctypes.byref(cfg), n,
                                               out.data_ptr()))
        return out

    def indices(self, n):
        import numpy as np
        idx = np.empty(n, np.uint32)
        check(lib.ssw_sharded_indices(self.handle, idx.ctypes.data, n))
        return idx

    def coefficients(self, which=0):
        import numpy as np
        c = np.empty((self.wb, self.height), np.float32)
        check(lib.ssw_sharded_coefficients(self.handle, which, c.ctypes.data))
        return c

    def overflow(self):
        v = ctypes.c_int(0)
        check(lib.ssw_sharded_overflow(self.handle, ctypes.byref(v)))
        return bool(v.value)

    def close(self):
        if self.handle:
            check(lib.ssw_sharded_destroy(self.handle))
            self.handle = None


class ShardedFrame:
    """forward-transformed frame: this rank's coefficient columns, transposed ([wb][H] f32)"""

    def __init__(self, rgb_rows, width, height, ops, group=None, rank=None, world=None):
        world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.plan = ShardPlan(width, height, world, rank)
        self.ops, self.group = ops, group
        p = self.plan
        if tuple(rgb_rows.shape) != (p.hb, p.width, 3):
            raise SswError(_lib.SSW_ERR_INVALID, 'expected this rank\'s rows as [%d][%d][3] uint8' % (p.hb, p.width))
        self.rgb_rows = rgb_rows
        self.shard = ssw_shard(p.width, p.height, p.col0, p.wb)
        # the row pass runs in `chunks` slices; the exchange of slice i overlaps the kernels of slice i+1
        chunks = _pick_chunks(p.hb, world)
        hc = p.hb // chunks
        with _scope(ops):
            recv = ops.empty((chunks, world, p.wb, hc), torch.float32) if world > 1 else None
            pending = []
            for ci in range(chunks):
                a = ops.lines_forward(rgb_rows[ci * hc:(ci + 1) * hc], p.width, hc, PIX_RGB8)   # rows: [hc][W]
                send = ops.transpose_blocks(a, hc, p.wb, p.width, world)                       # [G][wb][hc]
                if world == 1:
                    recv = send.view(1, 1, p.wb, hc)          # nothing to exchange: the transpose is the result
                else:
                    pending.append((_all_to_all_async(recv[ci], send, group), send))
            for work, _keep in pending:
                if work is not None:
                    work.wait()
            # line l of my columns = segments [rank g][chunk c] of recv[c][g][l]
            self.coeff = ops.lines_forward_segmented(recv, p.height)                  # cols: [wb][H]

    def ordered_indices(self, k, ordering=0):
        """first k entries of obtain_indices_by_function (src/algorithm.rs:200-210), identical on every rank"""
        p, ops = self.plan, self.ops
        k = min(int(k), p.width * p.height - 1)
        with _scope(ops):
            b = _all_reduce(ops.topk_bin(self.coeff, self.shard, ordering, k), dist.ReduceOp.MAX, self.group)
            cand, cnt = ops.topk_collect(self.coeff, self.shard, ordering, b)
            lists, counts = _all_gather(cand, self.group), _all_gather(cnt, self.group).reshape(-1).contiguous()
            idx, overflow = ops.topk_merge(lists, counts, k)
        self._overflow = overflow     # checked once per step (check_overflow): no host round trip on the critical path
        return idx

    def check_overflow(self):
        ov = getattr(self, '_overflow', None)
        self._overflow = None
        if ov is None:
            return
        with _scope(self.ops):          # read on the stream that wrote it
            overflowed = int(ov.item())
        if overflowed:
            raise SswError(_lib.SSW_ERR_UNSUPPORTED, 'sharded top-k: candidate overflow (flat spectrum); '
                           'the low-frequency bound was too loose for this frame')

    def inverse_rgb8(self):
        """DCT-III of the (possibly modified) coefficients back to this rank's RGB8 rows; consumes them"""
        p, ops = self.plan, self.ops
        chunks = _pick_chunks(p.wb, p.world)
        wc = p.wb // chunks
        with _scope(ops):
            recv = ops.empty((chunks, p.world, p.hb, wc), torch.float32) if p.world > 1 else None
            pending = []
            for ci in range(chunks):
                t = ops.lines_inverse(self.coeff[ci * wc:(ci + 1) * wc], p.height, wc, 1.0)     # cols: [wc][H]
                send = ops.transpose_blocks(t, wc, p.hb, p.height, p.world)                    # [G][hb][wc]
                if p.world == 1:
                    recv = send.view(1, 1, p.hb, wc)
                else:
                    pending.append((_all_to_all_async(recv[ci], send, self.group), send))
            for work, _keep in pending:
                if work is not None:
                    work.wait()
            out = ops.empty((p.hb, p.width, 3), torch.uint8)                         # rows: read the blocks in place
            ops.lines_inverse_segmented(recv, p.width, 4.0 / float(p.width * p.height), out, self.rgb_rows)
        self.coeff = None
        return out


class ShardedWriter:
    """`Writer::new(img, cfg).mark(&[marks])` (src/algorithm.rs:295-358) for a frame sharded by rows"""

    def __init__(self, rgb_rows, width, height, config, ops, group=None, rank=None, world=None):
        self.cfg = config if isinstance(config, ssw_config) else ssw_config(*config)
        self.frame = ShardedFrame(rgb_rows, width, height, ops, group, rank, world)
        self.ops = ops

    def embed(self, marks):
        """marks: list of 1-D float arrays of equal length (several marks: deltas against the original
        coefficients are summed, src/algorithm.rs:399-408)"""
        lens = {len(m) for m in marks}
        if len(lens) != 1:
            raise SswError(_lib.SSW_ERR_UNSUPPORTED, 'sharded embed takes marks of equal length')
        k = min(lens.pop(), self.frame.plan.width * self.frame.plan.height - 1)
        self.indices = self.frame.ordered_indices(k, self.cfg.ordering)
        import numpy as np
        with _scope(self.ops):
            m = self.ops.to_device(np.stack([np.asarray(x, dtype=np.float32)[:k] for x in marks]), torch.float32)
            self.ops.embed(self.frame.coeff, self.frame.shard, self.indices, m, self.cfg)

    def result_rgb8(self):
        out = self.frame.inverse_rgb8()
        self.frame.check_overflow()
        return out

    def mark_rgb8(self, marks):
        self.embed(marks)
        return self.result_rgb8()


class ShardedReader:
    """`Reader::base` + `Reader::derived` + `extract` (src/algorithm.rs:462-562) for sharded frames"""

    def __init__(self, base_rows, width, height, config, ops, group=None, rank=None, world=None):
        self.cfg = config if isinstance(config, ssw_config) else ssw_config(*config)
        self.args = (width, height, ops, group, rank, world)
        self.base = ShardedFrame(base_rows, *self.args)
        self.ops, self.group = ops, group

    def extract(self, derived_rows, n):
        p = self.base.plan
        if n >= p.width * p.height:
            raise SswError(_lib.SSW_ERR_INVALID, 'Desired extraction length exceeds available coefficients.')
        derived = ShardedFrame(derived_rows, *self.args)
        idx = self.base.ordered_indices(n, self.cfg.ordering)
        with _scope(self.ops):
            part = self.ops.extract(self.base.coeff, derived.coeff, self.base.shard, idx, n, self.cfg)
            out = _all_reduce(part, dist.ReduceOp.SUM, self.group)   # every index has exactly one owner
        self.base.check_overflow()
        return out
