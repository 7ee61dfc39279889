"""ORACLE (test infrastructure): rotated-box IoU, numpy restatement of the reference chain

torchdrivesim/_iou_utils.py:344-367  iou_differentiable_fast
  :270-299 box2corners_th   :42-84 box_intersection_th   :87-114 box1_in_box2
  :134-157 build_vertices   :160-227 sort_indices        :230-247 calculate_area

evaluated in float64 by default: the reference's own fp32 evaluation is chaotic on the
diagonal, so parity for IoU is defined against the reference run in
float64 (golden vectors in tests/golden/iou.npz), rtol 1e-5 + atol 2e-6.
"""
import numpy as np

EPS = 1e-8


def corners(box):
    """[P,5] (x,y,l,w,psi) -> [P,4,2], corner order (+,+), (-,+), (-,-), (+,-)."""
    box = np.asarray(box)
    x, y, l, w, a = (box[:, i:i + 1] for i in range(5))
    x4 = np.array([0.5, -0.5, -0.5, 0.5], box.dtype) * l
    y4 = np.array([0.5, 0.5, -0.5, -0.5], box.dtype) * w
    s, c = np.sin(a), np.cos(a)
    return np.stack([x4 * c - y4 * s + x, x4 * s + y4 * c + y], -1)


def _edge_intersections(c1, c2):
    l1a, l1b = c1, np.roll(c1, -1, axis=1)
    l2a, l2b = c2, np.roll(c2, -1, axis=1)
    x1, y1 = l1a[:, :, None, 0], l1a[:, :, None, 1]
    x2, y2 = l1b[:, :, None, 0], l1b[:, :, None, 1]
    x3, y3 = l2a[:, None, :, 0], l2a[:, None, :, 1]
    x4, y4 = l2b[:, None, :, 0], l2b[:, None, :, 1]
    num = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    den_t = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
    den_u = (x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = den_t / num
        u = -den_u / num
        small = np.abs(num) < 1e-4
        t = np.where(small, -1.0, t)
        u = np.where(small, -1.0, u)
        mask = (t > 0) & (t < 1) & (u > 0) & (u < 1)
        t2 = den_t / (num + EPS)
        pts = np.stack([x1 + t2 * (x2 - x1), y1 + t2 * (y2 - y1)], -1)
    pts = np.where(mask[..., None], pts, 0.0)
    return pts, mask


def _c1_in_c2(c1, c2):
    a, b, d = c2[:, 0:1], c2[:, 1:2], c2[:, 3:4]
    ab, am, ad = b - a, c1 - a, d - a
    with np.errstate(divide="ignore", invalid="ignore"):
        q1 = np.round((ab * am).sum(-1) / (ab * ab).sum(-1) * 1e6) / 1e6
        q2 = np.round((ad * am).sum(-1) / (ad * ad).sum(-1) * 1e6) / 1e6
    return (q1 > -1e-6) & (q1 < 1 + 1e-6) & (q2 > -1e-6) & (q2 < 1 + 1e-6)


def intersection_area(c1, c2):
    """c1, c2 [P,4,2] -> [P] intersection area following the reference's vertex-sort method."""
    p = c1.shape[0]
    inter, mask_inter = _edge_intersections(c1, c2)
    verts = np.concatenate([c1, c2, inter.reshape(p, 16, 2)], 1)                # [P,24,2]
    mask = np.concatenate([_c1_in_c2(c1, c2), _c1_in_c2(c2, c1), mask_inter.reshape(p, 16)], 1)

    def sort(mask):
        nv = mask.sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            center = (verts * mask[..., None]).sum(1, keepdims=True) / nv[:, None, None]
            rel = verts - center
            r = np.sqrt((rel ** 2).sum(-1))
            ac = np.arccos(rel[..., 0] / r)
        ang = np.where(rel[..., 1] > 0, ac, 2 * np.pi - ac)
        ang = np.where(mask, ang, np.inf)
        return np.argsort(ang, axis=1, kind="stable"), nv

    order, nv = sort(mask)
    while (nv > 8).any():
        mask = mask.copy()
        for i in np.nonzero(nv > 8)[0]:
            sv = verts[i, order[i]]
            dist = np.linalg.norm(sv[:-1] - sv[1:], axis=-1)
            dist[np.arange(23) >= nv[i] - 1] = np.inf
            mask[i, order[i, int(dist.argmin())]] = False
        order, nv = sort(mask)
    idx = order[:, :9].copy()
    pad = np.argmin(mask[:, 8:], axis=1) + 8
    pos = np.arange(9)[None, :]
    idx = np.where((pos >= nv[:, None]) | (nv[:, None] < 3), pad[:, None], idx)
    first = order[:, 0]
    first = np.where(nv < 3, pad, first)
    idx[np.arange(p), np.minimum(nv, 8)] = np.where(nv <= 8, first, idx[np.arange(p), np.minimum(nv, 8)])
    sel = np.take_along_axis(verts, idx[..., None].repeat(2, -1), axis=1)        # [P,9,2]
    tot = (sel[:, :-1, 0] * sel[:, 1:, 1] - sel[:, :-1, 1] * sel[:, 1:, 0]).sum(1)
    return np.abs(tot) / 2


def iou_pairwise(box1, box2, dtype=np.float64):
    """Element-wise API of the reference: [..,5] x [..,5] -> [..] IoU."""
    b1 = np.asarray(box1, dtype).reshape(-1, 5)
    b2 = np.asarray(box2, dtype).reshape(-1, 5)
    inter = intersection_area(corners(b1), corners(b2))
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (b1[:, 2] * b1[:, 3] + b2[:, 2] * b2[:, 3] - inter)
    return iou.reshape(np.asarray(box1).shape[:-1])


def iou_matrix(ego_box, all_box, dtype=np.float64):
    """[A,5] x [N,5] -> [A,N] for one environment."""
    a, n = len(ego_box), len(all_box)
    e = np.repeat(np.asarray(ego_box, dtype)[:, None], n, 1)
    o = np.repeat(np.asarray(all_box, dtype)[None], a, 0)
    return iou_pairwise(e, o, dtype)
