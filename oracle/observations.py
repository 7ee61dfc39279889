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


def noisy_state(all_state, n_agents, eps):
    """StandardSensingObservationNoise.get_noisy_state (observation_noise.py:74-89): all_state [B,N,4], eps [B,A,N,4]."""
    s = np.asarray(all_state, np.float32)
    A = n_agents
    d = s[:, :A, None, :2] - s[:, None, :, :2]
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1], dtype=np.float32)
    dev = np.max(np.stack([np.float32(0.19) * (dist > 0.5), np.float32(1.6) * (dist > 25), np.float32(3.2) * (dist > 50),
                           np.float32(3.83) * (dist > 100)], -1), -1).astype(np.float32)
    return (s[:, None] + np.asarray(eps, np.float32) * dev[..., None]).astype(np.float32)


def line_circle_intersection(p1, p2, centre, radius):
    """utils.py:139-187, float32 operation by operation."""
    f32 = np.float32
    d = (p2 - p1).astype(f32)
    f = (p1 - centre).astype(f32)
    a = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]
    b = f32(2) * (f[..., 0] * d[..., 0] + f[..., 1] * d[..., 1])
    c = (f[..., 0] * f[..., 0] + f[..., 1] * f[..., 1]) - radius * radius
    disc = b * b - f32(4) * a * c
    sq = np.sqrt(np.maximum(disc, f32(0)))
    a_safe = np.where(np.abs(a) < f32(1e-8), f32(1e-8), a)
    t1, t2 = (-b - sq) / (f32(2) * a_safe), (-b + sq) / (f32(2) * a_safe)
    return (disc >= 0) & (np.minimum(t1, t2) <= 1) & (np.maximum(t1, t2) >= 0)


def noisy_present_mask(all_state, all_size, base_mask, n_agents):
    """StandardSensingObservationNoise.get_noisy_present_mask (observation_noise.py:91-132) -> bool [B,A,N]."""
    s = np.asarray(all_state, np.float32)
    r = (np.asarray(all_size, np.float32)[..., 1] / np.float32(2))
    B, N = s.shape[:2]
    A = n_agents
    ego = s[:, :A, None, None, :2]
    target = s[:, None, :, None, :2]
    occ = s[:, None, None, :, :2]
    hit = line_circle_intersection(np.broadcast_to(ego, (B, A, N, N, 2)), np.broadcast_to(target, (B, A, N, N, 2)),
                                   np.broadcast_to(occ, (B, A, N, N, 2)), np.broadcast_to(r[:, None, None, :], (B, A, N, N)))
    hit &= ~np.eye(N, dtype=bool)[None, None]
    for a in range(A):
        hit[:, a, :, a] = False
    return np.asarray(base_mask, bool)[:, None] & ~hit.any(-1)
