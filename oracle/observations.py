"""ORACLE (test infrastructure): non-visual observations.

Restates  Simulator.get_all_agents_absolute / get_all_agents_relative   torchdrivesim/simulator.py:730-781
          utils.relative, rotate, normalize_angle                        torchdrivesim/utils.py:31-79
For every agent i (the origin) the pose of every agent j in i's frame:
  rel_xy = R(-psi_i) (xy_j - xy_i),  rel_psi = (psi_j - psi_i + pi) mod 2 pi - pi,  then length, width, present of j.
With exclude_self the entry j = i is removed (the reference does it with a boolean-mask index, which syncs the GPU).
Pinned by tests/golden/relative.npz (unmodified reference).
"""
import numpy as np


def agents_relative(absolute, n_agents=None, exclude_self=True):
    """absolute [B,N,6] (x, y, psi, length, width, present) -> [B,A,N(-1),6]; the first n_agents are the origins."""
    a = np.asarray(absolute, np.float32)
    B, N = a.shape[:2]
    A = N if n_agents is None else n_agents
    o = a[:, :A]
    d = a[:, None, :, :2] - o[:, :, None, :2]                                    # [B,A,N,2]
    s, c = np.sin(-o[:, :, None, 2]), np.cos(-o[:, :, None, 2])
    rel_xy = np.stack([c * d[..., 0] - s * d[..., 1], s * d[..., 0] + c * d[..., 1]], -1)
    rel_psi = (a[:, None, :, 2] - o[:, :, None, 2] + np.float32(np.pi)) % np.float32(2 * np.pi) - np.float32(np.pi)
    out = np.concatenate([rel_xy, rel_psi[..., None], np.broadcast_to(a[:, None, :, 3:], (B, A, N, 3))], -1).astype(np.float32)
    if exclude_self:
        keep = ~np.eye(A, N, dtype=bool)
        out = out[:, keep].reshape(B, A, N - 1, 6)
    return out
