"""Brute-force cosine retrieval on the B200.

The reference's retrieval step is `faiss.normalize_L2(arr); index = faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT);
index.add(arr); D, I = index.search(arr, k)` (daodian_infer.py:225-230, 295-302: catalogue x catalogue cosine top-13 /
26 / 100).  `CosineIndex` keeps that interface -- `add`, `ntotal`, `search(x, k) -> (D, I)` -- over the fused
top-k kernel (`ops.cosine_topk`: two cosine-GEMM passes and a shared-memory sort; the B x C score matrix is
never built, k <= 128; larger k falls back to materialised cosines).  Rows are L2-normalised on the way in (K1), so
inputs need not be normalised; any width d (zero-padded to a multiple of 8); scores are cosines computed from bf16
operands with fp32 accumulation (~1e-3 of fp32) or, with precision='bf16x3', from hi/lo bf16 pairs (~1e-5).
"""
from __future__ import annotations

import torch

from . import ops

MAX_K = 128


class CosineIndex:
    """`faiss.IndexFlat(d, METRIC_INNER_PRODUCT)` over L2-normalised rows, resident on one B200."""

    def __init__(self, d: int, device="cuda", precision="bf16"):
        """d: any width (rows are zero-padded to a multiple of 8, which leaves cosines unchanged: the reference builds
        IndexFlat(100) for its fastText vectors, daodian_infer.py:227).  precision: 'bf16' (scores within ~1e-3 of fp32
        cosines) or 'bf16x3' (within 1e-5) -- use the latter where scores are compared against tuned thresholds
        (nlp_score_th / cv_score_th in the reference's inference scripts)."""
        if d < 1:
            raise ValueError("d must be positive")
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.d = d
        self.dp = (d + 7) // 8 * 8
        self.precision = precision
        self.device = torch.device(device)
        self._chunks = []       # bf16 [n_i, dp] (or [n_i, 3 dp]) normalised rows
        self._what = None       # concatenation, built lazily

    def _rows(self, x, order: int) -> torch.Tensor:
        """fp32 [n, d] -> normalised bf16 operand rows (padded; three-part rows in the bf16x3 mode)."""
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("expected [n, %d] rows" % self.d)
        if self.dp != self.d:
            x = torch.nn.functional.pad(x, (0, self.dp - self.d))
        x = x.contiguous()
        if self.precision == "bf16x3":
            return ops.normalize_cast3(x, order)[0]
        return ops.normalize_cast(x)[0]

    @property
    def ntotal(self) -> int:
        return sum(c.shape[0] for c in self._chunks)

    def add(self, x) -> None:
        self._chunks.append(self._rows(x, 1))
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
        if k < 1:
            raise ValueError("k must be positive")
        what = self._catalogue()
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("expected [n, %d] queries" % self.d)
        out_v, out_i = [], []
        if k <= MAX_K:
            for lo in range(0, x.shape[0], ops.MAX_BATCH):      # the kernels take up to MAX_BATCH query rows per launch
                v, i = ops.cosine_topk(self._rows(x[lo:lo + ops.MAX_BATCH], 0), what, k)
                out_v.append(v)
                out_i.append(i)
        else:
            # k beyond the fused kernel's candidate lists (the reference searches with k = len(catalogue slice) in
            # daodian_infer.py:230): materialise the cosines with the logits kernel for a bounded block of queries and
            # let the library sort them -- a convenience path, O(n x ntotal) memory traffic like faiss's own
            C = what.shape[0]
            kk = min(k, C)
            step = max(1, min(ops.MAX_BATCH, (1 << 28) // max(C, 1)))
            for lo in range(0, x.shape[0], step):
                cos = ops.logits(self._rows(x[lo:lo + step], 0), what, None, None, 1.0)
                v, i = torch.topk(cos, kk, dim=1)
                if kk < k:   # like faiss: (-inf, -1) past the end of the catalogue
                    v = torch.nn.functional.pad(v, (0, k - kk), value=float("-inf"))
                    i = torch.nn.functional.pad(i, (0, k - kk), value=-1)
                out_v.append(v)
                out_i.append(i)
        return (out_v[0], out_i[0]) if len(out_v) == 1 else (torch.cat(out_v), torch.cat(out_i))


def cosine_topk(queries, catalogue, k: int):
    """One-shot `CosineIndex(d).add(catalogue).search(queries, k)`."""
    catalogue = torch.as_tensor(catalogue)
    index = CosineIndex(catalogue.shape[1], device=catalogue.device if catalogue.is_cuda else "cuda")
    index.add(catalogue)
    return index.search(queries, k)
