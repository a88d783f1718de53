"""Data-parallel plumbing (new; the reference is single-GPU — SURVEY §2.3, §8e).

The path shards on batch: one process per GPU, full weight replica, rank r trains on its own images.  The
only collective is ONE all-reduce (sum) of the flat fp32 gradient buffer per step; the optimiser then applies
``1/world`` (mean of per-rank mean-losses == the global mean loss when shards are equal-sized).  BatchNorm
statistics stay per rank, like the reference's plain BatchNorm2d under DDP(broadcast_buffers=False).
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class GradSync:

    def __init__(self, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) and \
            dist.is_initialized() else 1
        if process_group is None and self.world > 1:
            self.pg = dist.group.WORLD

    @property
    def grad_scale(self):
        """Factor the optimiser applies to the summed gradient."""
        return 1.0 / self.world

    def all_reduce_(self, flat_grad):
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
        return flat_grad

    def broadcast_(self, tensors, src=0):
        if self.world > 1:
            for t in tensors:
                dist.broadcast(t, src, group=self.pg)

    def shard(self, global_batch):
        """[begin, end) of this rank's images in a global batch (equal shards required)."""
        if global_batch % self.world:
            raise ValueError(f'global batch {global_batch} is not divisible by world size {self.world}')
        per = global_batch // self.world
        rank = dist.get_rank(self.pg) if self.world > 1 else 0
        return rank * per, (rank + 1) * per
