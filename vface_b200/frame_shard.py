"""Frame-chunk data parallelism for one video clip.

The reference runs single-GPU (REFace/VFace_video_swap_batch.sh:30).  Every op of the hot path is
per-frame except align_by_flow (scripts/temporal_flow.py:231-235), where frame i+1 reads the
post-FSAI, pre-alignment q/k of frame i.  So a clip shards by contiguous frame chunk with exactly one
exchange: per DDIM step and per flow-active attn1 module, rank r sends the q and k rows of its LAST
frame (2 x 4096 x 320 elements) to rank r+1.  No other collective exists on the path.

One process per GPU; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries
the halo as a single packed point-to-point message.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced, order-preserving: rank r owns [r*F // G, (r+1)*F // G)."""
    if world_size < 1 or total_frames < world_size:
        raise ValueError(f"cannot shard {total_frames} frames over {world_size} ranks")
    return [(r * total_frames // world_size, (r + 1) * total_frames // world_size) for r in range(world_size)]


class FrameShard:
    def __init__(self, rank: int, world_size: int, total_frames: int, group=None):
        self.rank = rank
        self.world_size = world_size
        self.total_frames = total_frames
        self.group = group
        self.lo, self.hi = shard_bounds(total_frames, world_size)[rank]
        self.halo_messages = 0
        self.halo_bytes = 0
        self._side = None                    # CUDA side stream of exchange_halo_async

    @property
    def frames(self) -> int:
        return self.hi - self.lo

    def take(self, t: torch.Tensor) -> torch.Tensor:
        """Rows [lo, hi) of a per-frame tensor."""
        return t[self.lo:self.hi]

    def local_flow(self, flow):
        """The flow fields this rank's warp needs, from the clip's F-1 fields (flow[i]: frame i -> i+1).
        Rank 0 gets frames-1 fields; rank r > 0 gets `frames` fields, the first one mapping the halo
        frame (last frame of rank r-1) onto its first frame."""
        if flow is None:
            return None
        first = max(self.lo - 1, 0)
        if isinstance(flow, (list, tuple)):
            return list(flow[first:self.hi - 1])
        return flow[first:self.hi - 1]

    def exchange_halo(self, q_last: torch.Tensor, k_last: torch.Tensor) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Send (q_last, k_last) of this rank's last frame to rank+1, receive the previous rank's.
        Returns (halo_q, halo_k), or (None, None) on rank 0."""
        if self.world_size == 1:
            return None, None
        ops_ = []
        recv = None
        if self.rank + 1 < self.world_size:
            send = torch.stack([q_last, k_last]).contiguous()
            ops_.append(dist.P2POp(dist.isend, send, self._peer(self.rank + 1), group=self.group))
            self.halo_messages += 1
            self.halo_bytes += send.numel() * send.element_size()
        if self.rank > 0:
            recv = torch.empty((2,) + tuple(q_last.shape), dtype=q_last.dtype, device=q_last.device)
            ops_.append(dist.P2POp(dist.irecv, recv, self._peer(self.rank - 1), group=self.group))
        for w in dist.batch_isend_irecv(ops_):
            w.wait()
        if recv is None:
            return None, None
        return recv[0], recv[1]

    def exchange_halo_async(self, q_last: torch.Tensor, k_last: torch.Tensor) -> "_PendingHalo":
        """exchange_halo without stalling the caller's stream: on CUDA the send/receive is issued on a side stream that
        waits for what has been enqueued so far (the FSAI of the last frame); .wait() makes the caller's stream wait for
        the received rows and returns (halo_q, halo_k) -- (None, None) on rank 0.  On CPU (gloo tests) it is synchronous."""
        if self.world_size == 1 or not q_last.is_cuda:
            return _PendingHalo(None, self.exchange_halo(q_last, k_last), None)
        main = torch.cuda.current_stream(q_last.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=q_last.device)
        side = self._side
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            q_last.record_stream(side)
            k_last.record_stream(side)
            halo = self.exchange_halo(q_last, k_last)
            done = torch.cuda.Event()
            done.record(side)
        return _PendingHalo(main, halo, done)

    def _peer(self, group_rank: int) -> int:
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)

    def gather_frames(self, local: torch.Tensor) -> torch.Tensor:
        """All ranks' per-frame results concatenated in frame order (used once, after sampling)."""
        if self.world_size == 1:
            return local
        bounds = shard_bounds(self.total_frames, self.world_size)
        parts = [torch.empty((hi - lo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for lo, hi in bounds]
        dist.all_gather(parts, local.contiguous(), group=self.group) if len({hi - lo for lo, hi in bounds}) == 1 \
            else _all_gather_uneven(parts, local.contiguous(), self)
        return torch.cat(parts, dim=0)


def _all_gather_uneven(parts, local, shard: "FrameShard"):
    for r, buf in enumerate(parts):
        if r == shard.rank:
            buf.copy_(local)
        dist.broadcast(buf, shard._peer(r), group=shard.group)


class _PendingHalo:
    def __init__(self, main, halo, done):
        self._main, self._halo, self._done = main, halo, done

    def wait(self):
        if self._done is not None:
            self._main.wait_event(self._done)
            for t in self._halo:
                if t is not None:
                    t.record_stream(self._main)       # allocated on the side stream, consumed on the caller's
        return self._halo


_current: Optional[FrameShard] = None


def activate(shard: Optional[FrameShard]):
    global _current
    _current = shard


def current() -> Optional[FrameShard]:
    return _current
