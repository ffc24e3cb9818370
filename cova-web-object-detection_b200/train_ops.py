"""Native versions of the callers either side of ``CoVA.forward``, each with the
interface of the object the reference builds in ``main.py`` / ``train.py`` / ``datasets.py``:

  ``CrossEntropyLossSum``  for ``nn.CrossEntropyLoss(reduction="sum")``      (`/root/reference/main.py:139`)
  ``FlatAdam``             for ``torch.optim.Adam(params, lr, weight_decay)`` (`main.py:133-135`)
  ``evaluate_model``       for ``train.evaluate_model``                       (`train.py:99-171`)
  ``assemble_batch``       for the context window + collate index work        (`datasets.py:117-128`, `:170-178`)

CUDA tensors only; every op is a call into ``libcova_b200.so`` (``train_tail.cu``)."""
from time import time

import numpy as np
import torch
import torch.nn as nn

from . import ops


class _CESumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index, holder):
        loss, dl, nc = ops.ce_sum_fwd_bwd(logits.detach().float(), labels, want_grad=True, want_correct=True,
                                          ignore_index=ignore_index)
        holder.n_correct = nc
        ctx.save_for_backward(dl)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None


class CrossEntropyLossSum(nn.Module):
    """`criterion(output, labels)` of `train.py:56`: sum-reduced cross entropy over all boxes of the batch.  The
    forward kernel also produces d loss / d logits (so `loss.backward()` starts from a ready gradient) and the number
    of arg-max hits (`train.py:53-54`), kept on the device in `.n_correct`."""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index
        self.n_correct = None

    def forward(self, output, labels):
        if not output.is_cuda:
            raise RuntimeError("cova_b200: CrossEntropyLossSum needs CUDA tensors (no CPU path)")
        return _CESumFn.apply(output, labels, self.ignore_index, self)


class FlatAdam(torch.optim.Optimizer):
    """`torch.optim.Adam` (amsgrad off, L2 `weight_decay` added to the gradient) whose parameters, gradients and
    moments live in ONE flat fp32 buffer each: `step()` is one kernel launch per parameter group, and `flat_grad` is
    the bucket the data-parallel all-reduce(SUM) runs on (`allreduce_grads`).

    Construct it after `model.to(device)` (as `main.py:122-135` does): every `p.data` / `p.grad` is re-pointed to a
    view of the flat buffers.  `zero_grad()` zeroes the bucket in place (the views stay).  `state_dict()` has the layout
    of torch's Adam (`step`, `exp_avg`, `exp_avg_sq` per parameter).  A parameter that never receives a gradient still
    sees a zero gradient here (torch skips it): every CoVA parameter is used by the forward, so the two coincide."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = []
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            if any(not p.is_cuda or p.dtype != torch.float32 for p in ps):
                raise RuntimeError("cova_b200: FlatAdam needs fp32 CUDA parameters (construct it after model.to(device))")
            n = sum(p.numel() for p in ps)
            n_pad = (n + 3) // 4 * 4
            dev = ps[0].device
            flat_p = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            flat_g, flat_m, flat_v = (torch.zeros_like(flat_p) for _ in range(3))
            step_t = torch.zeros((), dtype=torch.float32)
            off = 0
            with torch.no_grad():
                for p in ps:
                    k = p.numel()
                    flat_p[off:off + k].copy_(p.detach().reshape(-1))
                    p.data = flat_p[off:off + k].view(p.shape)
                    p.grad = flat_g[off:off + k].view(p.shape)
                    self.state[p] = dict(step=step_t, exp_avg=flat_m[off:off + k].view(p.shape),
                                         exp_avg_sq=flat_v[off:off + k].view(p.shape))
                    off += k
            self._flat.append(dict(p=flat_p, g=flat_g, m=flat_m, v=flat_v, step=step_t, n=n))
        ops.param_generation += 1

    @property
    def flat_grad(self):
        """The gradient bucket of the first parameter group (what the all-reduce runs on)."""
        return self._flat[0]["g"]

    def zero_grad(self, set_to_none=False):
        for f in self._flat:
            if f is not None:
                f["g"].zero_()

    def allreduce_grads(self, group=None, async_op=False):
        """all-reduce(SUM) of the flat gradient bucket - SUM because the loss is sum-reduced (`main.py:139`)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        works = [dist.all_reduce(f["g"], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
                 for f in self._flat if f is not None]
        return works if async_op else None

    def _adopt_stray_grads(self):
        """`model.zero_grad()` (set_to_none=True is torch's default) or `p.grad = None` makes autograd allocate a fresh
        gradient outside the bucket: copy it in and re-point `p.grad`, instead of stepping on a stale (zero) bucket."""
        for g, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            off = 0
            base = f["g"].data_ptr()
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                k = p.numel()
                view = f["g"][off:off + k].view(p.shape)
                if p.grad is None:
                    view.zero_()
                    p.grad = view
                elif p.grad.data_ptr() != base + 4 * off:
                    view.copy_(p.grad)
                    p.grad = view
                if p.data_ptr() != f["p"].data_ptr() + 4 * off:      # e.g. model.to(...) / p.data = ... after construction
                    f["p"][off:off + k].copy_(p.detach().reshape(-1))
                    p.data = f["p"][off:off + k].view(p.shape)
                off += k

    def load_state_dict(self, state_dict):
        """torch's Adam layout in (`step`, `exp_avg`, `exp_avg_sq` per parameter), copied INTO the flat moment buffers the
        kernel reads (the inherited method would replace `state[p]` with detached copies the kernel never sees)."""
        super().load_state_dict(state_dict)
        for g, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            off = 0
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                k = p.numel()
                st = self.state.get(p, {})
                if "exp_avg" in st:
                    f["m"][off:off + k].copy_(st["exp_avg"].reshape(-1).to(f["m"].device))
                    f["v"][off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(f["v"].device))
                    f["step"].fill_(float(st["step"]))
                self.state[p] = dict(step=f["step"], exp_avg=f["m"][off:off + k].view(p.shape),
                                     exp_avg_sq=f["v"][off:off + k].view(p.shape))
                off += k

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._adopt_stray_grads()
        for g, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            f["step"] += 1
            b1, b2 = g["betas"]
            ops.adam_step(f["p"], f["g"], f["m"], f["v"], g["lr"], b1, b2, g["eps"], g["weight_decay"],
                          int(f["step"].item()))
        return loss


def page_offsets_of(bboxes):
    """Row offsets int32 [B+1] of the pages of a collated batch (boxes of a page are contiguous and pages appear in
    batch order, `datasets.py:170-181`): one `torch.bincount` + cumsum on the device."""
    page = bboxes[:, 0].long()
    B = int(page.max().item()) + 1 if page.numel() else 0
    counts = torch.bincount(page, minlength=B)
    off = torch.zeros(B + 1, dtype=torch.int32, device=bboxes.device)
    off[1:] = counts.cumsum(0).to(torch.int32)
    return off


def assemble_batch(boxes_per_page, context_size, device, boxes_xywh=None):
    """N1: build `bboxes [T,5]` (from raw [x,y,w,h] rows, optional) and `context_indices [T, 2*context_size]` on the
    device from the per-page box counts alone - the Python loops of `datasets.py:117-128` and `:170-178` never run."""
    counts = torch.as_tensor(boxes_per_page, dtype=torch.int64)
    off = torch.zeros(counts.numel() + 1, dtype=torch.int32)
    off[1:] = counts.cumsum(0).to(torch.int32)
    off = off.to(device)
    if boxes_xywh is not None:
        boxes_xywh = boxes_xywh.to(device=device, dtype=torch.float32)
    return ops.build_batch(off, context_size, boxes_xywh) + (off,)


@torch.no_grad()
def evaluate_model(model, eval_loader, device, k=1, split_name="VAL", log_file="log.txt", print_and_log=None):
    """`train.evaluate_model` (`train.py:99-171`) with the per-image / per-class Python loop replaced by one segmented
    top-k launch per batch.  Same returns: `img_acc` int32 [n_imgs, n_classes] rows [img_id, acc_1, ...] and
    `class_acc` [n_classes] (percent, class 0 = 0).  `print_and_log` defaults to `print` + append to `log_file`
    like `utils.print_and_log`."""
    if print_and_log is None:
        def print_and_log(msg, lf):
            print(msg)
            with open(lf, "a") as f:
                f.write(msg + "\n")
    start = time()
    model.eval()
    n_classes = model.n_classes
    ids, hits = [], []
    for img_ids, images, bboxes, additional_feats, context_indices, labels in eval_loader:
        bboxes = bboxes.to(device)
        output = model(images.to(device), bboxes, additional_feats.to(device), context_indices.to(device))
        off = page_offsets_of(bboxes)
        hits.append(ops.topk_hits(output.float(), labels.to(device), off, k))
        ids.extend(list(img_ids)[: off.numel() - 1])
    hits = torch.cat(hits).cpu().numpy() if hits else np.zeros((0, n_classes), np.int32)
    if (hits[:, 1:] < 0).any():
        raise IndexError("evaluate_model: a page has no box of some non-background class (the reference indexes [0, 0] "
                         "of an empty tensor there, train.py:146)")
    img_acc = np.concatenate((np.asarray(ids).reshape(-1, 1), hits[:, 1:]), axis=1).astype(np.int32)
    class_acc = np.zeros(n_classes)
    class_acc[1:] = img_acc[:, 1:].mean(0) * 100
    print_and_log("[%s] Avg_class_Accuracy: %.2f%% (%.2fs)" % (split_name, class_acc[1:].mean(), time() - start), log_file)
    for c in range(1, n_classes):
        print_and_log("%s top-%d-Acc: %.2f%%" % (model.class_names[c], k, class_acc[c]), log_file)
    print_and_log("", log_file)
    return img_acc, class_acc
