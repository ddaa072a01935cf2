"""Brute-force cosine retrieval on the B200.

The reference's retrieval step is `faiss.normalize_L2(arr); index = faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT);
index.add(arr); D, I = index.search(arr, k)` (daodian_infer.py:225-230, 295-302: catalogue x catalogue cosine top-13 /
26 / 100).  `CosineIndex` keeps that interface -- `add`, `ntotal`, `search(x, k) -> (D, I)` -- over the fused
top-k kernel (`ops.cosine_topk`: two cosine-GEMM passes and a shared-memory sort; the B x C score matrix is
never built).  Rows are L2-normalised on the way in (K1), so inputs need not be normalised; scores are cosines
computed from bf16 operands with fp32 accumulation.
"""
from __future__ import annotations

import torch

from . import ops

MAX_K = 128


class CosineIndex:
    """`faiss.IndexFlat(d, METRIC_INNER_PRODUCT)` over L2-normalised rows, resident on one B200."""

    def __init__(self, d: int, device="cuda"):
        if d % 8 != 0:
            raise ValueError("d must be a multiple of 8")
        self.d = d
        self.device = torch.device(device)
        self._chunks = []       # bf16 [n_i, d] normalised rows
        self._what = None       # concatenation, built lazily

    @property
    def ntotal(self) -> int:
        return sum(c.shape[0] for c in self._chunks)

    def add(self, x) -> None:
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device).contiguous()
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("expected [n, %d] rows" % self.d)
        xhat, _, _ = ops.normalize_cast(x)
        self._chunks.append(xhat)
        self._what = None

    def reset(self) -> None:
        self._chunks, self._what = [], None

    def _catalogue(self) -> torch.Tensor:
        if self._what is None:
            if not self._chunks:
                raise RuntimeError("the index is empty")
            self._what = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks).contiguous()
            self._chunks = [self._what]
        return self._what

    @torch.no_grad()
    def search(self, x, k: int):
        """(D, I): cosine scores fp32 [n, k] (descending) and catalogue row ids int64 [n, k]; (-inf, -1) past the end
        of a catalogue smaller than k, like faiss."""
        if not 1 <= k <= MAX_K:
            raise ValueError("k must be in [1, %d]" % MAX_K)
        what = self._catalogue()
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device).contiguous()
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("expected [n, %d] queries" % self.d)
        out_v, out_i = [], []
        for lo in range(0, x.shape[0], ops.MAX_BATCH):          # the kernels take up to MAX_BATCH query rows per launch
            xhat, _, _ = ops.normalize_cast(x[lo:lo + ops.MAX_BATCH].contiguous())
            v, i = ops.cosine_topk(xhat, what, k)
            out_v.append(v)
            out_i.append(i)
        return (out_v[0], out_i[0]) if len(out_v) == 1 else (torch.cat(out_v), torch.cat(out_i))


def cosine_topk(queries, catalogue, k: int):
    """One-shot `CosineIndex(d).add(catalogue).search(queries, k)`."""
    catalogue = torch.as_tensor(catalogue)
    index = CosineIndex(catalogue.shape[1], device=catalogue.device if catalogue.is_cuda else "cuda")
    index.add(catalogue)
    return index.search(queries, k)
