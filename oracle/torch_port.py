"""ORACLE - TEST / BASELINE INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU port of the reference forward (`/root/reference/models.py:94-212`) written against the SAME arithmetic
library the reference calls (torch ATen on CPU, torchvision's compiled `roi_pool`), operator for operator and
in the reference's as-written form (the GAT materialises [T,K,F], [T,K,H] x2 and [T,K,2H], `models.py:180-208`).
It exists because `/root/reference` cannot travel to the GPU box: `bench.py` times it on the box's host
cores as the `cpu_baseline` ("kind": "port") and as the `--impl reference` arm.  Validated against the live
reference by `tests/test_oracle.py::test_torch_port_matches_golden`.
"""
import numpy as np
import torch
import torch.nn.functional as F

try:   # the op the reference calls (models.py:58); pure-numpy restatement otherwise
    from torchvision.ops import roi_pool as _tv_roi_pool
except Exception:   # pragma: no cover
    _tv_roi_pool = None


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def backbone(sd, images):
    x = F.relu(_bn(sd, "convnet.1", F.conv2d(images, sd["convnet.0.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    b = 0
    while f"convnet.4.{b}.conv1.weight" in sd:
        p = f"convnet.4.{b}"
        idt = x
        if p + ".conv3.weight" in sd:
            o = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
            o = F.relu(_bn(sd, p + ".bn2", F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1)))
            o = _bn(sd, p + ".bn3", F.conv2d(o, sd[p + ".conv3.weight"]))
            if p + ".downsample.0.weight" in sd:
                idt = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"]))
        else:
            o = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)))
            o = _bn(sd, p + ".bn2", F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1))
        x = F.relu(o + idt)
        b += 1
    return x


def roi_pool(fm, bboxes, P, scale):
    if _tv_roi_pool is not None:
        return _tv_roi_pool(fm, bboxes, P, scale)
    from . import cova_oracle as O
    return torch.from_numpy(O.roi_pool(fm.numpy(), bboxes.numpy(), P, scale))


def gat_as_written(sd, h_i, ci, prefix="gat"):
    N, K = ci.shape
    Wi, Wj = sd[prefix + ".W_i.weight"], sd[prefix + ".W_j.weight"]
    hp = torch.cat((h_i, torch.zeros((1, h_i.shape[1]))), 0)
    h_j = hp[ci.view(-1)].view(N, K, -1)
    Wh_i = F.linear(h_i, Wi)
    Wh_i_rep = Wh_i.repeat_interleave(K, dim=0).view(N, K, -1)
    Wh_j = F.linear(h_j, Wj)
    e = F.linear(torch.cat((Wh_i_rep, Wh_j), 2), sd[prefix + ".attention_layer.weight"],
                 sd[prefix + ".attention_layer.bias"]).squeeze(2)
    e = F.leaky_relu(e, 0.2)
    e = torch.where(ci >= 0, e, -9e15 * torch.ones_like(e))
    a = torch.softmax(e, dim=1)
    return (a.unsqueeze(-1) * Wh_j).sum(1)


@torch.no_grad()
def forward(sd, images, bboxes, additional_feats, context_indices, roi_output_size=(3, 3)):
    """Eval-mode `CoVA.forward`; `sd` = state_dict of CPU tensors."""
    fm = backbone(sd, images)
    scale = fm.shape[2] / images.shape[2]
    T = bboxes.shape[0]
    parts = [roi_pool(fm, bboxes, roi_output_size, scale).reshape(T, -1)]
    if "bbox_feat_encoder.0.weight" in sd:
        f = bboxes[:, 1:].clone()
        f[:, 2:] -= f[:, :2]
        f = torch.cat((f, (f[:, 2] / f[:, 3]).view(-1, 1)), 1)
        f = F.linear(f, sd["bbox_feat_encoder.0.weight"], sd["bbox_feat_encoder.0.bias"])
        parts.append(F.relu(_bn(sd, "bbox_feat_encoder.1", f)))
    if "bn_additional_feat.weight" in sd:
        additional_feats = _bn(sd, "bn_additional_feat", additional_feats)
    parts.append(additional_feats)
    own = torch.cat(parts, 1)
    if "gat.W_i.weight" in sd:
        ctx = gat_as_written(sd, own, context_indices)
    elif "gat.heads.0.W_i.weight" in sd:
        n = len([k for k in sd if k.startswith("gat.heads.") and k.endswith(".W_i.weight")])
        ctx = torch.cat([gat_as_written(sd, own, context_indices, f"gat.heads.{i}") for i in range(n)], 1)
    else:
        ctx = own[:, :0]
    x = torch.cat((own, ctx), 1)
    x = F.relu(_bn(sd, "decoder.2", F.linear(x, sd["decoder.1.weight"], sd["decoder.1.bias"])))
    return F.linear(x, sd["decoder.5.weight"], sd["decoder.5.bias"])
