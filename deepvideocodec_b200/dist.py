"""Multi-GPU plumbing: the hot path shards by independent units, nothing else.

The reference is single-GPU (``CUDA_VISIBLE_DEVICES`` hard-coded, train.py:43,
test.py:14; ``nn.DataParallel`` inert, train.py:598-600).  Its natural unit of
independence is the (sequence, GOP): the decoded-picture buffer is reset at
every I-frame (test.py:162-173) and sequences are processed in a serial loop
(test.py:275-281).  So: one process per GPU, ``units[rank::world]`` per rank,
NO collective on the data path, and one all-reduce (NCCL over NVLink on the GPU
box, gloo in the CPU tests) of a handful of fp64 scalars per report.
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["Unit", "make_units", "shard_units", "RateStats", "reduce_stats"]

GOP = 32    # I-frame period of the reference (literal 32, test.py:162)


@dataclass(frozen=True)
class Unit:
    """One independently decodable piece of work: frames [start, stop) of a
    sequence; frame ``start`` is the I-frame (not on the hot path)."""
    sequence: int
    start: int
    stop: int

    @property
    def p_frames(self) -> int:
        return max(0, self.stop - self.start - 1)


def make_units(frames_per_sequence: Sequence[int], gop: int = GOP) -> List[Unit]:
    """Split every sequence at its I-frames (every ``gop`` frames)."""
    units = []
    for s, n in enumerate(frames_per_sequence):
        for start in range(0, n, gop):
            units.append(Unit(s, start, min(n, start + gop)))
    return units


def shard_units(units: Sequence[Unit], rank: int, world: int) -> List[Unit]:
    """Greedy longest-first assignment (ties broken by order) so ranks get
    balanced P-frame counts; deterministic and identical on every rank."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    order = sorted(range(len(units)), key=lambda i: (-units[i].p_frames, i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += units[i].p_frames
        if r == rank:
            mine.append(i)
    return [units[i] for i in sorted(mine)]


@dataclass
class RateStats:
    """Per-rank accumulators of a report interval (all fp64)."""
    bits: float = 0.0
    sq_err: float = 0.0
    frames: float = 0.0
    pixels: float = 0.0
    extra: List[float] = field(default_factory=list)

    def as_tensor(self, device) -> torch.Tensor:
        return torch.tensor([self.bits, self.sq_err, self.frames, self.pixels, *self.extra],
                            dtype=torch.float64, device=device)

    @staticmethod
    def from_tensor(t: torch.Tensor) -> "RateStats":
        v = t.detach().cpu().tolist()
        return RateStats(v[0], v[1], v[2], v[3], list(v[4:]))

    @property
    def bpp(self) -> float:
        return self.bits / self.pixels if self.pixels else float("nan")

    @property
    def mse(self) -> float:
        return self.sq_err / self.pixels if self.pixels else float("nan")


def reduce_stats(stats: RateStats, device="cpu", group=None) -> RateStats:
    """Sum the accumulators over all ranks: ONE all-reduce of <= 8 fp64 words
    per report, off the per-frame path.  A world of one is a no-op."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    t = stats.as_tensor(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return RateStats.from_tensor(t)
