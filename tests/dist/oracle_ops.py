"""TEST INFRASTRUCTURE: numpy stand-in for the per-rank arithmetic steps of
spread_spectrum_watermarking_b200.sharded (same interface as CudaOps), so that the partition / exchange /
index / merge logic of the sharded path runs on CPU under gloo.  Arithmetic comes from the oracle
(oracle/ssw_oracle.py); tensors are CPU torch tensors.  Never imported by the product."""
import numpy as np
import scipy.fft
import torch

import ssw_oracle as so

TOPK_CAP = 8192
F32 = np.float32


class OracleOps:
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def to_device(self, array, dtype):
        return torch.as_tensor(np.asarray(array), dtype=dtype)

    def lines_forward(self, src, n, n_lines, src_type, out=None):
        a = src.numpy()
        if src_type == 0:
            y, _, _ = so.rgb32f_to_yiq(so.rgb8_to_rgb32f(a))
        else:
            y = a.reshape(n_lines, n)
        r = scipy.fft.dct(y.astype(np.float64), type=2, axis=1).astype(F32)
        res = torch.from_numpy(r)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def lines_inverse(self, plane, n, n_lines, scale, dst_type=2, dst=None, src_type=2, src=None):
        y = (0.25 * scipy.fft.dct(plane.numpy().astype(np.float64), type=3, axis=1) * np.float64(F32(scale))).astype(F32)
        if dst_type == 2:
            (plane if dst is None else dst).copy_(torch.from_numpy(y))
            return plane if dst is None else dst
        _, i, q = so.rgb32f_to_yiq(so.rgb8_to_rgb32f(src.numpy()))
        dst.copy_(torch.from_numpy(so.rgb32f_to_rgb8(so.yiq_to_rgb32f(y, i, q))))
        return dst

    def transpose_blocks(self, src, rows, cols, ld, nblocks):
        a = src.numpy().reshape(rows, ld)
        return torch.from_numpy(np.stack([a[:, j * cols:(j + 1) * cols].T.copy() for j in range(nblocks)]))

    def interleave_blocks(self, recv):
        c, g, lines, seg = recv.shape
        return recv.permute(2, 1, 0, 3).contiguous().view(lines, g * c * seg)

    def lines_forward_segmented(self, recv, n):
        t = self.interleave_blocks(recv)
        return self.lines_forward(t, n, t.shape[0], 2, out=t)

    def lines_inverse_segmented(self, recv, n, scale, dst, pixels):
        a = self.interleave_blocks(recv)
        return self.lines_inverse(a, n, a.shape[0], scale, 0, dst, 0, pixels)

    @staticmethod
    def _keys(plane, shard, ordering):
        ncols, h = plane.shape
        c_local, r = np.meshgrid(np.arange(ncols), np.arange(h), indexing='ij')
        p = (r * shard.width + shard.col0 + c_local).astype(np.int64)
        v = plane.numpy().astype(F32)
        if ordering == 0:
            val = (v * v).astype(F32)
        else:
            # ordering_values works on a whole row-major plane; evaluate the scaling per element instead
            full = np.zeros(shard.width * shard.height, F32)
            full[p.ravel()] = v.ravel()
            val = so.ordering_values(full, ordering, shard.width, shard.height)[p.ravel()].reshape(v.shape)
        return so._total_cmp_key(val.ravel()).reshape(v.shape), p

    def topk_bin(self, plane, shard, ordering, k):
        key, p = self._keys(plane, shard, ordering)
        blk = key[:128, :256][p[:128, :256] != 0]
        b = int(np.sort(blk)[::-1][k - 1] >> 20) if blk.size >= k else 0
        return torch.tensor([b], dtype=torch.int32)

    def topk_collect(self, plane, shard, ordering, bin_t):
        key, p = self._keys(plane, shard, ordering)
        sel = ((key >> 20) >= int(bin_t.item())) & (p != 0)
        comp = (key[sel].astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - p[sel].astype(np.uint64))
        cand = np.zeros(TOPK_CAP, np.uint64)
        n = min(comp.size, TOPK_CAP)
        cand[:n] = comp[:n]
        return torch.from_numpy(cand.view(np.int64)), torch.tensor([comp.size], dtype=torch.int32)

    def topk_merge(self, lists, counts, k):
        parts = [lists[i].numpy().view(np.uint64)[:min(int(counts[i]), TOPK_CAP)] for i in range(lists.shape[0])]
        allc = np.concatenate(parts)
        overflow = int(sum(int(c) for c in counts) > TOPK_CAP or allc.size < k)
        top = np.sort(allc)[::-1][:k]
        idx = (np.uint64(0xFFFFFFFF) - (top & np.uint64(0xFFFFFFFF))).astype(np.int64)
        out = np.zeros(k, np.int32)
        out[:idx.size] = idx
        return torch.from_numpy(out), torch.tensor([overflow], dtype=torch.int32)

    @staticmethod
    def _owned(shard, idx):
        p = idx.numpy().astype(np.int64)
        r, c = p // shard.width, p % shard.width
        own = (c >= shard.col0) & (c < shard.col0 + shard.ncols)
        return own, (c - shard.col0) * shard.height + r

    def embed(self, plane, shard, idx, marks, cfg):
        own, q = self._owned(shard, idx)
        flat = plane.view(-1).numpy()
        mk = [m.numpy() for m in marks]
        pos = q[own]
        sub = so.embed_watermark(flat[pos].copy(), np.arange(pos.size), [m[own] for m in mk], cfg.method, cfg.alpha)
        flat[pos] = sub

    def extract(self, base, derived, shard, idx, n, cfg):
        own, q = self._owned(shard, idx)
        out = np.zeros(n, F32)
        pos = q[own]
        b, d = base.view(-1).numpy()[pos], derived.view(-1).numpy()[pos]
        out[own] = so.extract_watermark(np.concatenate([b, [1.0]]), np.arange(pos.size), np.concatenate([d, [1.0]]),
                                        pos.size, cfg.method, cfg.alpha)
        return torch.from_numpy(out)
