"""Peer-memory exchanges for the class-sharded head: the host side of csrc/p2p.cu.

One symmetric allocation per (head, shapes) -- `torch.distributed._symmetric_memory` maps every rank's buffer
into every process over NVLink -- laid out as [flags 1 KB | channel 0 | channel 1 | channel 2]:
  channel 0  all-gather of the packed local (x | labels)        slot = b_loc * (4 D + 8) bytes per rank
  channel 1  all-gather of the packed per-row statistics         slot = 20 B bytes per rank
  channel 2  reduce-scatter of the embedding-gradient partials   slot = b_loc * D * 4 bytes per rank (summed by
             normalize_bwd_x_sum in fixed rank order)
The kernels only store to peers and spin on local flags; allocation and rendezvous happen once, outside any
CUDA-graph capture.  NCCL (`engine.py`) remains the path for gloo / non-NVLink groups and whenever symmetric memory
is unavailable.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib

FLAG_BYTES = 1024


def _up(v: int, k: int = 256) -> int:
    return (v + k - 1) // k * k


class PeerExchange:
    def __init__(self, group, device, b_cap: int, D: int):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.b_cap, self.D = b_cap, D      # slots are sized for up to b_cap local rows; messages may be shorter
        B = b_cap * self.world
        self.slot = (_up(b_cap * (4 * D + 8), 16), _up(20 * B, 16), _up(b_cap * D * 4, 16))
        self.off = []
        o = FLAG_BYTES
        for s in self.slot:
            self.off.append(o)
            o = _up(o + s * self.world)
        self.buf = symm_mem.empty(o, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._flags = (ctypes.c_uint64 * 16)(*(ptrs + [0] * (16 - len(ptrs))))
        self._bufs = [(ctypes.c_uint64 * 16)(*([p + off for p in ptrs] + [0] * (16 - len(ptrs)))) for off in self.off]
        self.sync = torch.zeros(16, dtype=torch.int32, device=device)   # [call number, arrival counter] per channel
        # raised by the exchange kernel when a peer never shows up (csrc/p2p.cu); pinned host memory the device can
        # write, examined before every exchange: a lost rank becomes a Python exception, not a dead CUDA context
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        # Host-side record of the last channel ISSUED (eagerly or by a graph replay).  Channels 0 and 1 are always
        # separated by the other one, which is what makes their buffers safe to reuse (csrc/p2p.cu); two scatters in
        # a row (two backwards of one head without a forward between them) are not, so the second one is declined
        # and the caller takes the NCCL reduce-scatter instead.  Every rank sees the same call sequence.
        self.last_channel = -1
        torch.cuda.synchronize(device)
        dist.barrier(group=group)   # every rank's flags are zero before anyone stores to them

    def matches(self, b_loc: int, D: int) -> bool:
        return b_loc <= self.b_cap and D == self.D

    def check(self) -> None:
        """Raise if an earlier exchange timed out waiting for a peer (the results since then are meaningless)."""
        code = int(self.err[0])
        if code != 0:
            raise RuntimeError("multimodalsimilar_b200: peer-memory exchange on channel %d timed out waiting for a rank "
                               "(ARCFACE_B200_E_CUDA): a peer process died or the ranks issued different call sequences"
                               % (code - 1))

    def _exchange(self, channel: int, src: torch.Tensor, bytes_per_peer: int, src_stride: int) -> torch.Tensor:
        self.check()
        _lib.call("arcface_b200_p2p_exchange", ctypes.c_void_p(src.data_ptr()), bytes_per_peer, src_stride,
                  self._bufs[channel], self._flags, self.rank, self.world, self.slot[channel], channel,
                  ctypes.c_void_p(self.sync.data_ptr()), ctypes.c_void_p(self.err.data_ptr()),
                  torch.cuda.current_stream().cuda_stream)
        self.last_channel = channel
        lo = self.off[channel]
        return self.buf[lo: lo + self.slot[channel] * self.world].view(self.world, self.slot[channel])

    def all_gather_split(self, channel: int, packed: torch.Tensor, split: int):
        """All-gather of `packed` uint8 [n] whose first `split` bytes and last n - split bytes arrive as two contiguous
        rank-ordered regions: returns (uint8 [R * split], uint8 [R * (n - split)]) views of the receive buffer."""
        n = packed.numel()
        if n % 16 != 0 or split % 16 != 0 or not (0 < split < n) or n > self.slot[channel]:
            raise ValueError("message of %d bytes (split at %d) does not fit channel %d" % (n, split, channel))
        self.check()
        _lib.call("arcface_b200_p2p_gather_split", ctypes.c_void_p(packed.data_ptr()), n, 0, split, self._bufs[channel],
                  self._flags, self.rank, self.world, self.slot[channel], channel, ctypes.c_void_p(self.sync.data_ptr()),
                  ctypes.c_void_p(self.err.data_ptr()), torch.cuda.current_stream().cuda_stream)
        self.last_channel = channel
        lo = self.off[channel]
        R = self.world
        return self.buf[lo: lo + R * split], self.buf[lo + R * split: lo + R * n]

    def all_gather_bytes(self, channel: int, packed: torch.Tensor) -> torch.Tensor:
        """packed uint8 [n] (n a multiple of 16, <= the channel's slot) -> uint8 [R, n] view of the receive buffer."""
        n = packed.numel()
        if n % 16 != 0 or n > self.slot[channel]:
            raise ValueError("message of %d bytes does not fit channel %d" % (n, channel))
        return self._exchange(channel, packed, n, 0)[:, :n]

    def scatter_rows(self, full: torch.Tensor) -> torch.Tensor:
        """full fp32 [R * b_loc, D] (this rank's partial for every rank's rows) -> fp32 [R, b_loc, D]: the partials
        every rank sent for THIS rank's rows (to be summed by normalize_bwd_x_sum); None when the previous exchange
        was a scatter too (see `last_channel`)."""
        if self.last_channel == 2:
            return None
        b_loc = full.shape[0] // self.world
        n = b_loc * self.D * 4
        if n > self.slot[2]:
            raise ValueError("%d rows per rank exceed the exchange's capacity of %d" % (b_loc, self.b_cap))
        out = self._exchange(2, full, n, n)
        return out[:, :n].view(torch.float32).view(self.world, b_loc, self.D)
