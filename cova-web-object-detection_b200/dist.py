"""Multi-GPU plumbing for the hot path: pages are independent units, so the batch shards by page with no
data-path collective; training adds exactly one all-reduce(SUM) of a flat fp32 gradient bucket per step
(SURVEY.md section 8(e)).  SUM, not mean, because the reference loss is `CrossEntropyLoss(reduction="sum")`
(`/root/reference/main.py:139`).  One process per GPU (`torchrun`); NCCL on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def page_range(n_pages, rank, world):
    """Contiguous, balanced page interval [lo, hi) of `rank`."""
    base, rem = divmod(n_pages, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(images, bboxes, additional_feats, context_indices, labels=None, rank=0, world=1):
    """Slice a `custom_collate_fn` batch (`datasets.py:159-190`) down to this rank's pages and re-base the
    batch-index column and the batch-global context ids to be shard-local (-1 padding is kept)."""
    lo, hi = page_range(images.shape[0], rank, world)
    page = bboxes[:, 0].long()
    rows = ((page >= lo) & (page < hi)).nonzero(as_tuple=True)[0]
    if rows.numel():
        r0, r1 = int(rows[0]), int(rows[-1]) + 1   # boxes of a page are contiguous, pages in order
    else:
        r0 = r1 = 0
    bb = bboxes[r0:r1].clone()
    bb[:, 0] -= lo
    ci = context_indices[r0:r1].clone()
    ci[ci >= 0] -= r0
    out = [images[lo:hi], bb, additional_feats[r0:r1], ci]
    if labels is not None:
        out.append(labels[r0:r1])
    return tuple(out)


class FlatGradBucket:
    """All parameter gradients as views into ONE contiguous fp32 buffer, so the per-step exchange is a single
    all-reduce (6.5 MB for ResNet-18 / 37.8 MB for ResNet-50 - latency-bound on NVLink 5 / NVSwitch)."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_sum(self, group=None, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
