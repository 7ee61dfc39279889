"""Kinematic step kernels vs the oracle (forward: rtol 1e-5 as BASELINE.json states; backward vs autograd)."""
import math

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _rand(B, A, gen):
    state = torch.cat([torch.rand(B, A, 2, generator=gen) * 400, torch.rand(B, A, 1, generator=gen) * 12 - 6,
                       torch.rand(B, A, 1, generator=gen) * 8 - 2], -1)
    action = torch.rand(B, A, 2, generator=gen) * 2 - 1
    lr = 1 + 2 * torch.rand(B, A, generator=gen)
    return state, action, lr


def test_golden_trajectories():
    """12-step rollouts of the reference's KinematicBicycle / BicycleNoReversing (tests/golden/kinematic.npz)."""
    import torchdrivesim_b200 as tds
    g = util.golden("kinematic")
    for case in g["cases"]:
        q = lambda k: g[f"{case}/{k}"]
        cls = tds.BicycleNoReversing if case.startswith("noreverse") else tds.KinematicBicycle
        km = cls(left_handed=bool(q("left_handed")))
        km.set_params(lr=torch.tensor(q("lr")).cuda())
        km.set_state(torch.tensor(q("state0")).cuda())
        acts = torch.tensor(q("actions")).cuda()
        for t in range(acts.shape[0]):
            km.step(acts[t])
            np.testing.assert_allclose(km.get_state().cpu().numpy(), q("traj")[t + 1], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("left_handed", [False, True])
def test_compound_forward_backward(left_handed):
    from oracle import kinematic as K
    import torchdrivesim_b200 as tds
    gen = torch.Generator().manual_seed(7)
    B, A = 5, 33
    state, action, lr = _rand(B, A, gen)
    model = torch.randint(0, 3, (B, A), generator=gen)
    w = torch.randn(B, A, 4, generator=gen)
    s_o, a_o, l_o = (t.clone().requires_grad_(True) for t in (state, action, lr))
    out_o = K.compound_step(s_o, a_o, l_o, model, 0.1, left_handed)
    (out_o * w).sum().backward()
    km = tds.FusedCompoundKinematicModel(model.cuda(), left_handed=left_handed)
    s_g, a_g, l_g = (t.clone().cuda().requires_grad_(True) for t in (state, action, lr))
    km.set_params(lr=l_g)
    km.set_state(s_g)
    km.step(a_g)
    out_g = km.get_state()
    (out_g * w.cuda()).sum().backward()
    np.testing.assert_allclose(out_g.detach().cpu().numpy(), out_o.detach().numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(s_g.grad.cpu().numpy(), s_o.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a_g.grad.cpu().numpy(), a_o.grad.numpy(), rtol=1e-4, atol=1e-5)
    uni = (model == 2)
    np.testing.assert_allclose(l_g.grad.cpu().numpy()[~uni.numpy()], l_o.grad.numpy()[~uni.numpy()], rtol=1e-4, atol=1e-5)


def test_unicycle_and_rollout_gradient():
    """20-step differentiable rollout (config 5 shape): gradient w.r.t. the actions matches autograd."""
    from oracle import kinematic as K
    import torchdrivesim_b200 as tds
    gen = torch.Generator().manual_seed(11)
    B, A, T = 4, 16, 20
    state, _, lr = _rand(B, A, gen)
    acts = torch.rand(T, B, A, 2, generator=gen) * 2 - 1
    target = state[..., :2] + 3.0
    a_o = acts.clone().requires_grad_(True)
    s = state
    for t in range(T):
        s = K.bicycle_step(s, a_o[t], lr, 0.1, True)
    ((s[..., :2] - target) ** 2).sum().backward()
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=lr.cuda())
    km.set_state(state.cuda())
    a_g = acts.clone().cuda().requires_grad_(True)
    for t in range(T):
        km.step(a_g[t])
    ((km.get_state()[..., :2] - target.cuda()) ** 2).sum().backward()
    np.testing.assert_allclose(km.get_state().detach().cpu().numpy(), s.detach().numpy(), rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(a_g.grad.cpu().numpy(), a_o.grad.numpy(), rtol=2e-4, atol=1e-4)
    # unicycle forward
    u = tds.KinematicUnicycle(max_yaw_rate=1.0)
    u.set_state(state.cuda())
    u.step(acts[0].cuda())
    ref = K.unicycle_step(state, acts[0], 0.1, 5.0, 1.0)
    np.testing.assert_allclose(u.get_state().cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)


def test_edge_cases():
    import torchdrivesim_b200 as tds
    km = tds.KinematicBicycle()
    km.set_params(lr=torch.zeros(2, 0).cuda())
    km.set_state(torch.zeros(2, 0, 4).cuda())
    km.step(torch.zeros(2, 0, 2).cuda())          # empty batch of agents
    assert km.get_state().shape == (2, 0, 4)
    km = tds.KinematicBicycle()
    km.set_params(lr=torch.ones(1, 3))
    km.set_state(torch.zeros(1, 3, 4))
    with pytest.raises(tds._lib.TdsError):        # CPU tensors: no fallback
        km.step(torch.zeros(1, 3, 2))


def test_simple_and_oriented_models_and_padded_compound():
    """SimpleKinematicModel / OrientedKinematicModel (kinematic.py:328-397) and a compound batch whose action
    is padded to 4 values, forward and backward."""
    from oracle import kinematic as K
    import torchdrivesim_b200 as tds
    gen = torch.Generator().manual_seed(5)
    B, A = 4, 21
    state, _, lr = _rand(B, A, gen)
    act4 = torch.rand(B, A, 4, generator=gen) * 2 - 1
    for cls, oriented in ((tds.SimpleKinematicModel, False), (tds.OrientedKinematicModel, True)):
        km = cls()
        km.set_state(state.cuda())
        km.step(act4.cuda())
        ref = K.simple_step(state, act4, oriented=oriented)
        np.testing.assert_allclose(km.get_state().cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)
        nxt = km.get_state()
        km.set_state(state.cuda())
        np.testing.assert_allclose(km.fit_action(nxt).cpu().numpy(), act4.numpy(), atol=5e-4)
    model = torch.randint(0, 5, (B, A), generator=gen)
    w = torch.randn(B, A, 4, generator=gen)
    s_o, a_o = state.clone().requires_grad_(True), act4.clone().requires_grad_(True)
    out_o = K.compound_step(s_o, a_o, lr, model, 0.1, True)
    (out_o * w).sum().backward()
    km = tds.FusedCompoundKinematicModel(model.cuda(), left_handed=True, action_size=4)
    s_g, a_g = state.clone().cuda().requires_grad_(True), act4.clone().cuda().requires_grad_(True)
    km.set_params(lr=lr.cuda())
    km.set_state(s_g)
    km.step(a_g)
    (km.get_state() * w.cuda()).sum().backward()
    np.testing.assert_allclose(km.get_state().detach().cpu().numpy(), out_o.detach().numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(s_g.grad.cpu().numpy(), s_o.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a_g.grad.cpu().numpy(), a_o.grad.numpy(), rtol=1e-4, atol=1e-5)
