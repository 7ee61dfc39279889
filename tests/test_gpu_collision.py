"""Collision kernels (discs fwd+bwd, IoU fwd, fused all-pairs) vs the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
IOU_RTOL, IOU_ATOL = 1e-5, 2e-6       # IoU is compared against the reference in float64


def _boxes(B, N, rng, spread=6.0, ped_every=3, zero_size_every=0):
    xy = rng.uniform(0, 300, (B, 1, 2)) + spread * rng.standard_normal((B, N, 2))
    lw = np.tile(np.array(util.VEH[:2], np.float32), (B, N, 1))
    if ped_every:
        lw[:, ped_every - 1::ped_every] = util.PED
    if zero_size_every:
        lw[:, zero_size_every - 1::zero_size_every] = 0.0
    psi = rng.uniform(-7, 7, (B, N, 1))
    return np.concatenate([xy, lw, psi], -1).astype(np.float32)


def test_golden_discs_and_iou():
    import torchdrivesim_b200 as tds
    g = util.golden("collision")
    box = torch.tensor(g["box"]).cuda()
    present = torch.tensor(g["present"]).cuda()
    B, A = box.shape[:2]
    e = box.unsqueeze(2).expand(-1, -1, A, -1).reshape(B, A * A, 5).contiguous()
    o = box.unsqueeze(1).expand(-1, A, -1, -1).reshape(B, A * A, 5).contiguous()
    pair = tds.collision_detection_with_discs(e, o).reshape(B, A, A).cpu().numpy()
    np.testing.assert_allclose(pair, g["discs_pair"], rtol=RTOL, atol=ATOL)
    st = torch.cat([box[..., :2], box[..., 4:5], torch.zeros(B, A, 1).cuda()], -1).requires_grad_(True)
    bx = torch.cat([st[..., :2], box[..., 2:4], st[..., 2:3]], -1)
    coll = tds.collision_allpairs(bx, bx, present, "discs")
    np.testing.assert_allclose(coll.detach().cpu().numpy(), g["discs_collision"], rtol=RTOL, atol=ATOL)
    coll.sum().backward()
    np.testing.assert_allclose(st.grad.cpu().numpy()[..., :3], g["discs_grad_state"][..., :3], rtol=1e-4, atol=1e-5)
    iou = tds.iou_differentiable(e, o).reshape(B, A, A).cpu().numpy()
    np.testing.assert_allclose(iou, g["iou64"], rtol=IOU_RTOL, atol=IOU_ATOL)


@pytest.mark.parametrize("B,A,N", [(4, 9, 9), (2, 64, 64), (1, 70, 300), (3, 1, 1)])
def test_discs_allpairs_vs_oracle(B, A, N):
    from oracle import collision as C
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(B * 100 + N)
    allb = _boxes(B, N, rng, zero_size_every=7 if N > 8 else 0)
    mask = rng.uniform(size=(B, N)) > 0.25
    w = rng.standard_normal((B, A)).astype(np.float32)
    a_o = torch.tensor(allb, requires_grad=True)
    out_o = C.collision_allpairs(a_o[:, :A], a_o, torch.tensor(mask))
    (out_o * torch.tensor(w)).sum().backward()
    a_g = torch.tensor(allb).cuda().requires_grad_(True)
    out_g = tds.collision_allpairs(a_g[:, :A], a_g, torch.tensor(mask).cuda(), "discs")
    (out_g * torch.tensor(w).cuda()).sum().backward()
    np.testing.assert_allclose(out_g.detach().cpu().numpy(), out_o.detach().numpy(), rtol=RTOL, atol=2e-6)
    go, gg = a_o.grad.numpy(), a_g.grad.cpu().numpy()
    # zero-size (padding) boxes: their 0/0 self pair makes the reference's autograd NaN for the whole box;
    # the kernel returns the finite gradient of the remaining pairs.  Compare where the oracle is finite.
    ok = np.isfinite(go).all(-1)
    assert ok.mean() > 0.8 and np.isfinite(gg).all()
    np.testing.assert_allclose(gg[ok], go[ok], rtol=2e-4, atol=2e-5)


def test_discs_pairwise_gradients_incl_size():
    from oracle import collision as C
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(3)
    b1 = _boxes(1, 400, rng, spread=2.5)[0]
    b2 = _boxes(1, 400, rng, spread=2.5)[0]
    b2[:, :2] = b1[:, :2] + rng.normal(0, 2.0, (400, 2)).astype(np.float32)
    t1o, t2o = torch.tensor(b1, requires_grad=True), torch.tensor(b2, requires_grad=True)
    C.discs_pairwise(t1o, t2o).sum().backward()
    t1g, t2g = torch.tensor(b1).cuda().requires_grad_(True), torch.tensor(b2).cuda().requires_grad_(True)
    out = tds.collision_detection_with_discs(t1g[None], t2g[None])
    out.sum().backward()
    assert int((out > 0).sum()) > 50
    np.testing.assert_allclose(t1g.grad.cpu().numpy(), t1o.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(t2g.grad.cpu().numpy(), t2o.grad.numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("B,N", [(3, 24), (1, 200)])
def test_iou_allpairs_vs_oracle_f64(B, N):
    from oracle import iou as I
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(N)
    allb = _boxes(B, N, rng, spread=5.0)
    allb[0, 1] = allb[0, 0]                                  # identical boxes
    allb[0, 2, :4] = allb[0, 3, :4]; allb[0, 2, 4] = allb[0, 3, 4]; allb[0, 2, 0] += 1.0   # parallel, shifted
    mask = rng.uniform(size=(B, N)) > 0.2
    g = torch.tensor(allb).cuda()
    e = g.unsqueeze(2).expand(-1, -1, N, -1).reshape(B, N * N, 5).contiguous()
    o = g.unsqueeze(1).expand(-1, N, -1, -1).reshape(B, N * N, 5).contiguous()
    pair = tds.iou_differentiable(e, o).reshape(B, N, N).cpu().numpy()
    ref = np.stack([I.iou_matrix(allb[b], allb[b]) for b in range(B)])
    np.testing.assert_allclose(pair, ref, rtol=IOU_RTOL, atol=IOU_ATOL)
    # fused aggregate with diag := 1
    refd = ref.copy()
    for b in range(B):
        np.fill_diagonal(refd[b], 1.0)
    om = refd * mask[:, None, :]
    agg = om.sum(-1) - om.max(-1)
    out = tds.collision_allpairs(g, g, torch.tensor(mask).cuda(), "iou").cpu().numpy()
    np.testing.assert_allclose(out, agg, rtol=1e-5, atol=N * IOU_ATOL)


def test_simulator_collision_properties_full_size():
    """Size-independent checks at the config-2 size (1024 x 64): permutation invariance of the row sums,
    zero for isolated agents, symmetry of the pair matrix."""
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(0)
    B, N = 1024, 64
    allb = torch.tensor(_boxes(B, N, rng, spread=15.0)).cuda()
    mask = torch.ones(B, N, dtype=torch.bool).cuda()
    out = tds.collision_allpairs(allb, allb, mask, "discs")
    perm = torch.randperm(N).cuda()
    out_p = tds.collision_allpairs(allb[:, perm], allb[:, perm], mask, "discs")
    np.testing.assert_allclose(out_p.cpu().numpy(), out[:, perm].cpu().numpy(), rtol=1e-5, atol=1e-5)
    far = allb.clone()
    far[..., 0] += torch.arange(N).cuda() * 100.0
    assert float(tds.collision_allpairs(far, far, mask, "discs").abs().max()) == 0.0
    e = allb[:8].unsqueeze(2).expand(-1, -1, N, -1).reshape(8, N * N, 5).contiguous()
    o = allb[:8].unsqueeze(1).expand(-1, N, -1, -1).reshape(8, N * N, 5).contiguous()
    pm = tds.collision_detection_with_discs(e, o).reshape(8, N, N)
    np.testing.assert_allclose(pm.cpu().numpy(), pm.transpose(1, 2).cpu().numpy(), rtol=0, atol=1e-6)
    assert float((torch.diagonal(pm, dim1=1, dim2=2) - 1).abs().max()) == 0.0


def test_iou_backward_matches_float64_finite_differences():
    """IoU backward (pairwise and fused all-pairs) against central differences of the float64 oracle."""
    from oracle import iou as I
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(17)
    P = 300
    b1 = _boxes(1, P, rng, spread=2.0)[0]
    b2 = _boxes(1, P, rng, spread=2.0)[0]
    b2[:, :2] = b1[:, :2] + rng.normal(0, 1.5, (P, 2)).astype(np.float32)
    t1 = torch.tensor(b1).cuda().requires_grad_(True)
    t2 = torch.tensor(b2).cuda().requires_grad_(True)
    out = tds.iou_differentiable(t1[None], t2[None])
    out.sum().backward()
    base = I.iou_pairwise(b1, b2)
    np.testing.assert_allclose(out[0].detach().cpu().numpy(), base, rtol=IOU_RTOL, atol=IOU_ATOL)
    eps = 1e-6
    for which, (grad, arr) in enumerate(((t1.grad, b1), (t2.grad, b2))):
        g = grad.cpu().numpy()
        for k in range(5):
            hi, lo = arr.astype(np.float64).copy(), arr.astype(np.float64).copy()
            hi[:, k] += eps; lo[:, k] -= eps
            args_hi = (hi, b2.astype(np.float64)) if which == 0 else (b1.astype(np.float64), hi)
            args_lo = (lo, b2.astype(np.float64)) if which == 0 else (b1.astype(np.float64), lo)
            fd = (I.iou_pairwise(*args_hi) - I.iou_pairwise(*args_lo)) / (2 * eps)
            fd2 = (I.iou_pairwise(*[a if i != which else arr.astype(np.float64) + (np.arange(5) == k) * 10 * eps
                                    for i, a in enumerate((b1.astype(np.float64), b2.astype(np.float64)))]) - base) / (10 * eps)
            smooth = np.abs(fd - fd2) < 1e-3 + 1e-2 * np.abs(fd)     # skip kinks (a vertex crossing an edge)
            ok = smooth & (base > 1e-4)
            assert ok.sum() > 100
            np.testing.assert_allclose(g[ok, k], fd[ok], rtol=2e-3, atol=2e-4)
    assert float(t1.grad[out[0] == 0].abs().max()) == 0.0
    # fused aggregate: gradient equals the sum of the pairwise gradients with the arg-max column removed
    B, N = 2, 20
    allb = _boxes(B, N, rng, spread=3.0)
    mask = rng.uniform(size=(B, N)) > 0.2
    a_g = torch.tensor(allb).cuda().requires_grad_(True)
    w = torch.tensor(rng.standard_normal((B, N)).astype(np.float32)).cuda()
    agg = tds.collision_allpairs(a_g, a_g, torch.tensor(mask).cuda(), "iou")
    (agg * w).sum().backward()
    a_p = torch.tensor(allb).cuda().requires_grad_(True)
    e = a_p.unsqueeze(2).expand(-1, -1, N, -1).reshape(B, N * N, 5)
    o = a_p.unsqueeze(1).expand(-1, N, -1, -1).reshape(B, N * N, 5)
    pm = tds.iou_differentiable(e, o).reshape(B, N, N)
    eye = torch.eye(N, dtype=torch.bool).cuda()
    pm = torch.where(eye, torch.ones_like(pm), pm) * torch.tensor(mask).cuda()[:, None, :]
    ref = pm.sum(-1) - pm.max(-1)[0]
    (ref * w).sum().backward()
    np.testing.assert_allclose(agg.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(a_g.grad.cpu().numpy(), a_p.grad.cpu().numpy(), rtol=1e-3, atol=1e-4)


def test_infraction_metrics_vector():
    """tds_infraction_metrics against the torch expression of distributed.infraction_metrics, accumulation included;
    there is no CPU implementation of the op."""
    import torchdrivesim_b200 as tds
    from torchdrivesim_b200 import distributed as D
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    for shape in [(7, 33), (1, 1), (1024, 64), (0, 5)]:
        coll = torch.tensor(rng.uniform(-0.5, 2, shape).clip(0).astype(np.float32), device=dev)
        off = torch.tensor(rng.uniform(-3, 5, shape).clip(0).astype(np.float32), device=dev)
        present = torch.tensor(rng.uniform(size=shape) > 0.25, device=dev)
        got = tds.ops.infraction_metrics(coll, off, present)
        ref = D.infraction_metrics(coll.cpu(), off.cpu(), present.cpu())
        assert got.dtype == torch.float64 and got.shape == (6,)
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=1e-6)
        assert np.array_equal(got[2:].cpu().numpy(), ref[2:].numpy())          # the counts are exact
        twice = tds.ops.infraction_metrics(coll, off, present, got.clone())
        np.testing.assert_allclose(twice.cpu().numpy(), 2 * ref.numpy(), rtol=1e-6)
        again = tds.ops.infraction_metrics(coll, off, present)
        assert torch.equal(again, got)                                           # fixed reduction order
    everyone = tds.ops.infraction_metrics(coll, off)
    assert float(everyone[4]) == float(everyone[5]) == coll.numel()
    with pytest.raises(tds._lib.TdsError):
        tds.ops.infraction_metrics(coll.cpu(), off.cpu())
