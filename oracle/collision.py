"""ORACLE (test infrastructure): disc collision metric, torch (autograd gives the backward).

discs_pairwise        torchdrivesim/infractions.py:503-545 collision_detection_with_discs
                      + infractions.py:378-409 bbox2discs
collision_allpairs    torchdrivesim/simulator.py:1064-1109, 1161-1194 (Simulator.compute_collision):
                      out[b,i] = sum_j o_ij m_j - max_j o_ij m_j
"""
import math

import torch


def bbox2discs(box, num_discs=5):
    """box [...,5] (x,y,length,width,psi) -> centres [...,num_discs,2], radius [...]."""
    n_side = (num_discs - 1) // 2
    xy, ln, wd, yaw = box[..., 0:2], box[..., 2], box[..., 3], box[..., 4]
    r = torch.minimum(ln, wd) / 2
    offs = torch.stack([i * ((torch.maximum(ln, wd) / 2) - r) / n_side for i in range(-n_side, n_side + 1)], dim=-1)
    yaw = yaw + (math.pi / 2) * (wd > ln)
    cx = offs * torch.cos(yaw).unsqueeze(-1)
    cy = offs * torch.sin(yaw).unsqueeze(-1)
    centres = torch.stack([cx, cy], dim=-1) + xy.unsqueeze(-2)
    return centres, r


def discs_pairwise(box1, box2, num_discs=5):
    """Element-wise API of the reference: box1, box2 [...,5] -> [...] overlap in [0,1]."""
    c1, r1 = bbox2discs(box1, num_discs)
    c2, r2 = bbox2discs(box2, num_discs)
    shape = r1.shape
    # torch.cdist as in infractions.py:527: its backward returns 0 (not NaN) at zero distance
    d = torch.cdist(c1.reshape(-1, num_discs, 2), c2.reshape(-1, num_discs, 2), p=2.0)
    d = d.reshape(-1, num_discs * num_discs).min(dim=-1)[0].reshape(shape)
    return torch.relu(1 - d / (r1 + r2))


def overlap_matrix(ego_box, all_box, metric_fn=discs_pairwise):
    """[B,A,5] x [B,N,5] -> [B,A,N]"""
    a, n = ego_box.shape[1], all_box.shape[1]
    e = ego_box.unsqueeze(2).expand(-1, -1, n, -1)
    o = all_box.unsqueeze(1).expand(-1, a, -1, -1)
    return metric_fn(e, o)


def collision_allpairs(ego_box, all_box, mask, metric_fn=discs_pairwise):
    """Simulator.compute_collision aggregate. mask [B,N] bool of column presence."""
    ego_box = torch.nan_to_num(ego_box, nan=0.0)
    all_box = torch.nan_to_num(all_box, nan=0.0)
    o = torch.nan_to_num(overlap_matrix(ego_box, all_box, metric_fn), nan=0.0)
    o = o * mask.to(o.dtype).unsqueeze(1)
    return o.sum(-1) - o.max(dim=-1)[0]
