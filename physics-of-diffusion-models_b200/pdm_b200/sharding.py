"""2-D layout of one node's GPUs for the statistics path: dataset shards x query groups.

SURVEY.md section 8e names two ways the path shards: dataset rows (every GPU sees all queries, one log-sum-exp-aware merge
per launch) and queries (no collective at all).  Round 1 used the first alone and paid for it: every rank regenerated and
prepared ALL noised queries (17 ms of a 107 ms step on 8 GPUs, constant in the number of GPUs).  The grid does both:

    rank r  ->  (data index i = r % G_d,  query index j = r // G_d)

  * the G_d ranks with the same j form a DATA group: each holds rows [i N/G_d, (i+1) N/G_d) of the dataset and they merge
    partial records (32 bytes per query row, all-gather) exactly as before;
  * the G_q = world / G_d ranks with the same i form a QUERY group: the temperatures of a schedule are dealt round-robin over
    them (j, j + G_q, j + 2 G_q, ...: balanced for any monotone schedule, also with screening on), each rank draws only the
    noise of its own temperatures -- at the Philox offsets those draws have in the single stream a one-GPU run consumes --
    and the (n_T, B) results are all-gathered at the end of the call.

The dataset (0.6 - 9.8 GB in the named configurations) is replicated across query groups; 180 GB of HBM3e per GPU make that free.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass
class ShardGrid:
    data_group: Optional[object] = None      # torch.distributed group of the ranks that share this rank's queries
    query_group: Optional[object] = None     # group of the ranks that share this rank's dataset shard
    data_index: int = 0
    data_shards: int = 1
    query_index: int = 0
    query_groups: int = 1

    @property
    def world(self) -> int:
        return self.data_shards * self.query_groups

    def rows(self, n_total: int) -> tuple[int, int]:
        """[lo, hi) of the dataset rows this rank holds."""
        per = (n_total + self.data_shards - 1) // self.data_shards
        return min(n_total, self.data_index * per), min(n_total, (self.data_index + 1) * per)

    def describe(self) -> str:
        return f"dataset rows / {self.data_shards} x temperatures / {self.query_groups}"


def default_data_shards(world: int) -> int:
    """2 dataset shards whenever the world size allows it: the merge step of the north-star design stays live on every
    multi-GPU run while the replicated query preparation shrinks by world/2; an odd world size splits queries only."""
    return 2 if world % 2 == 0 else 1


def make_grid(data_shards: int = 0) -> ShardGrid:
    """Build the grid over torch.distributed's default group.  Collective: every rank must call it with the same argument
    (``new_group`` is).  ``data_shards`` = 0 picks ``default_data_shards``."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ShardGrid()
    world, rank = dist.get_world_size(), dist.get_rank()
    gd = data_shards if data_shards > 0 else default_data_shards(world)
    if world % gd != 0:
        raise ValueError(f"data_shards = {gd} does not divide the world size {world}")
    gq = world // gd
    i, j = rank % gd, rank // gd
    data_group = query_group = None
    if gd == world:
        data_group = dist.group.WORLD
    elif gd > 1:
        groups = [dist.new_group([jj * gd + ii for ii in range(gd)]) for jj in range(gq)]
        data_group = groups[j]
    if gq == world:
        query_group = dist.group.WORLD
    elif gq > 1:
        groups = [dist.new_group([jj * gd + ii for jj in range(gq)]) for ii in range(gd)]
        query_group = groups[i]
    return ShardGrid(data_group, query_group, i, gd, j, gq)
