"""Batched attention-weights export: what `/root/reference/extract_attn_wts_and_visualize.py`
lines 89-150 compute one page at a time - for every non-background box its [x, y, w, h], label, the [x, y, w, h] of
its K context boxes (zeros for padding) and its K attention weights - for whole collated batches, through the native
kernels (`_get_visual_features` / `_get_bbox_features` / fused GAT with `return_attn_wts=True`), written in the
reference's csv format (`np.savetxt(..., delimiter=",", fmt="%.3f")`, one file per page, `:140-145`)."""
import os

import numpy as np
import torch


@torch.no_grad()
def attention_rows(model, images, bboxes, additional_feats, context_indices, labels):
    """One collated batch on the device -> list (one per page) of float32 arrays [n_fg, 5 + 4K + K]."""
    N = bboxes.shape[0]
    bbox_coords = bboxes[:, 1:].clone()                    # :107-110  [x1,y1,x2,y2] -> [x,y,w,h]
    bbox_coords[:, 2:] -= bbox_coords[:, :2]
    padded = torch.cat((bbox_coords, torch.zeros((1, 4), device=bboxes.device)), dim=0)     # :112-113 (-1 -> zero row)
    context_bbox_coords = padded[context_indices.view(-1)].view(N, -1)                       # :114-116
    visual = model._get_visual_features(images, bboxes)                                      # :118
    own = torch.cat((visual, model._get_bbox_features(bboxes), model.bn_additional_feat(additional_feats)), dim=1)
    _, attn = model.gat(own, context_indices, return_attn_wts=True)                          # :123-125
    if attn.dim() == 3:                                    # multi-head variant: [N, heads, K] -> heads side by side
        attn = attn.reshape(N, -1)
    dump = torch.cat((bbox_coords, labels.float().view(-1, 1), context_bbox_coords, attn), dim=1)   # :132-141
    fg = labels > 0                                        # :127-130
    page = bboxes[:, 0].long()
    dump, page = dump[fg].cpu().numpy(), page[fg].cpu().numpy()
    return [dump[page == b] for b in range(images.shape[0])]


def export_attention(model, batches, device, out_dir):
    """`batches`: iterable of collated batches (img_ids, images, bboxes, additional_feats, context_indices, labels)
    of any batch size.  Writes `<out_dir>/<img_id>.csv` per page; returns the list of files."""
    os.makedirs(out_dir, exist_ok=True)
    model.eval()
    files = []
    for img_ids, images, bboxes, additional_feats, context_indices, labels in batches:
        rows = attention_rows(model, images.to(device), bboxes.to(device), additional_feats.to(device),
                              context_indices.to(device), labels.to(device))
        for img_id, r in zip(img_ids, rows):
            path = "%s/%s.csv" % (out_dir, img_id)
            np.savetxt(path, r, delimiter=",", fmt="%.3f")
            files.append(path)
    return files
