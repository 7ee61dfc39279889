"""FusedRollout (BASELINE config 5): the graph-captured forward + backward rollout against the eager autograd path of the
same kernels and against the oracle's torch-autograd restatement (oracle/kinematic.py, oracle/collision.py,
oracle/offroad.py) - loss and d loss / d actions."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _inputs(B, A, T, seed, absent=0.0):
    rng = np.random.default_rng(seed)
    m = util.load_map_np("carla_Town01")
    state, size, types, present = util.random_scene(m, B, A, rng, spread=6.0, absent_p=absent)
    actions = rng.uniform(-1, 1, (T, B, A, 2)).astype(np.float32)
    target = (state[..., :2] + rng.normal(0, 5, (B, A, 2))).astype(np.float32)
    return m, state, size, present, actions, target


def _eager(town, state, size, lr, present, actions, target, w, threshold=0.5):
    import torchdrivesim_b200 as tds
    a = actions.clone().requires_grad_(True)
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=lr)
    km.set_state(state)
    sim = tds.Simulator(town, km, size, present, tds.TorchDriveConfig(left_handed_coordinates=True, offroad_threshold=threshold))
    loss = 0.0
    for t in range(a.shape[0]):
        sim.step(a[t])
        loss = loss + w[0] * sim.compute_collision().sum() + w[1] * sim.compute_offroad().sum() \
            + w[2] * ((sim.get_state()[..., :2] - target) ** 2).mean()
    loss.backward()
    return loss.detach(), a.grad, sim.get_state().detach()


@pytest.mark.parametrize("B,A,T,absent,w", [(3, 9, 6, 0.0, (1.0, 1.0, 1.0)), (2, 17, 4, 0.3, (0.5, 2.0, 0.25)), (5, 64, 3, 0.0, (1.0, 0.0, 0.0))])
def test_fused_rollout_matches_eager_autograd(B, A, T, absent, w):
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    m, state, size, present, actions, target = _inputs(B, A, T, 11 + A, absent)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    st, sz, pr = torch.tensor(state, device=dev), torch.tensor(size, device=dev), torch.tensor(present, device=dev)
    lr = torch.full((B, A), util.VEH[2], device=dev)
    act, tgt = torch.tensor(actions, device=dev), torch.tensor(target, device=dev)
    ro = tds.FusedRollout(town, st, sz, lr, pr, T, offroad_threshold=0.5, left_handed=True, target_xy=tgt,
                          w_collision=w[0], w_offroad=w[1], w_target=w[2])
    loss, grad = ro.run(act)
    loss, grad = float(loss), grad.clone()          # run() returns static buffers that the next replay overwrites
    loss_e, grad_e, final = _eager(town, st, sz, lr, pr, act, tgt, w)
    torch.cuda.synchronize()
    assert torch.isfinite(grad).all() and float(grad.abs().max()) > 0
    np.testing.assert_allclose(loss, float(loss_e), rtol=1e-5)
    scale = float(grad_e.abs().max())
    np.testing.assert_allclose(grad.cpu().numpy(), grad_e.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)
    assert torch.equal(ro.trajectory[-1], final)
    # a second replay with other actions (static buffers, same graph) and back: same numbers
    loss2, _ = ro.run(act * 0.5)
    assert float(loss2) != loss
    loss3, grad3 = ro.run(act)
    assert float(loss3) == loss        # the forward pass is deterministic; the collision backward accumulates with float atomics
    np.testing.assert_allclose(grad3.cpu().numpy(), grad.cpu().numpy(), rtol=1e-5, atol=1e-6 * scale)


def test_fused_rollout_against_the_oracle():
    """Small case against plain torch autograd through the oracle's restatement of the reference functions."""
    import torchdrivesim_b200 as tds
    from oracle import collision as OC, kinematic as OK, offroad as OO
    dev = torch.device("cuda:0")
    B, A, T = 2, 6, 4
    m, state, size, present, actions, target = _inputs(B, A, T, 5)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    lr = torch.full((B, A), util.VEH[2])
    ro = tds.FusedRollout(town, torch.tensor(state, device=dev), torch.tensor(size, device=dev), lr.to(dev),
                          torch.tensor(present, device=dev), T, offroad_threshold=0.5, left_handed=True,
                          target_xy=torch.tensor(target, device=dev), w_offroad=0.0)
    loss, grad = ro.run(torch.tensor(actions, device=dev))
    a = torch.tensor(actions, requires_grad=True)
    s = torch.tensor(state)
    ref = 0.0
    for t in range(T):
        s = OK.bicycle_step(s, a[t], lr, 0.1, True)
        box = torch.cat([s[..., :2], torch.tensor(size), s[..., 2:3]], -1)
        ref = ref + OC.collision_allpairs(box, box, torch.tensor(present)).sum() + ((s[..., :2] - torch.tensor(target)) ** 2).mean()
    ref.backward()
    np.testing.assert_allclose(float(loss), float(ref), rtol=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), a.grad.numpy(), rtol=2e-4, atol=2e-5 * float(a.grad.abs().max()))
    off = OO.offroad_loss(ro.trajectory[-1][0].cpu().numpy(), size[0], m["verts"], m["faces"], 0.5, present[0])
    assert off.shape == (A,)


def test_agent_boxes_and_heading_ops():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    st = torch.tensor(rng.normal(0, 10, (3, 5, 4)).astype(np.float32), device=dev, requires_grad=True)
    sz = torch.tensor(rng.uniform(1, 5, (3, 5, 2)).astype(np.float32), device=dev, requires_grad=True)
    box = tds.ops.agent_boxes(st, sz)
    want = torch.cat([st[..., :2], sz, st[..., 2:3]], -1)
    assert torch.equal(box, want)
    g = torch.randn_like(box)
    box.backward(g)
    gs, gz = torch.autograd.grad(want, (st, sz), g)
    assert torch.equal(st.grad, gs) and torch.equal(sz.grad, gz)
    sc = tds.ops.heading_sincos(st.detach())
    ref = np.stack([np.sin(st.detach().cpu().numpy()[..., 2].astype(np.float64)), np.cos(st.detach().cpu().numpy()[..., 2].astype(np.float64))], -1)
    assert np.array_equal(sc.cpu().numpy(), ref.astype(np.float32))


def test_torch_library_ops_match_the_autograd_functions():
    """torch.ops.tds_b200.* (registered custom operators) against the autograd.Function wrappers of ops.py: values and
    gradients, and torch.library.opcheck's consistency tests of the registrations."""
    import torchdrivesim_b200 as tds
    from torchdrivesim_b200 import ops, _lib
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(2)
    B, A = 3, 7
    mk = lambda *shape, lo=-1.0, hi=1.0: torch.tensor(rng.uniform(lo, hi, shape).astype(np.float32), device=dev)
    state, action, lr = mk(B, A, 4, lo=-20, hi=20), mk(B, A, 2), mk(B, A, lo=1.5, hi=2.5)
    size = mk(B, A, 2, lo=1.5, hi=5.0)
    mask = torch.tensor(rng.uniform(size=(B, A)) > 0.2, device=dev)

    def run(kin, boxes, allpairs):
        s, a, z = state.clone().requires_grad_(True), action.clone().requires_grad_(True), size.clone().requires_grad_(True)
        s1 = kin(s, a)
        box = boxes(s1, z)
        loss = allpairs(box, mask).sum() + (s1 ** 2).sum() * 1e-3
        loss.backward()
        return loss.detach(), s.grad, a.grad, z.grad

    ref = run(lambda s, a: ops.kinematic_step(s, a, lr, None, _lib.MODEL_BICYCLE, ops.kinematic_params(left_handed=True)),
              ops.agent_boxes, lambda b, m: ops.collision_allpairs(b, b, m, _lib.METRIC_DISCS, True))
    new = run(lambda s, a: torch.ops.tds_b200.kinematic_step(s, a, lr, None, 0, 0.1, True),
              torch.ops.tds_b200.agent_boxes, lambda b, m: torch.ops.tds_b200.collision_allpairs(b, b, m, 0, True)[0])
    for r, n in zip(ref, new):
        assert torch.equal(r, n)
    b1, b2 = torch.cat([state[..., :2], size, state[..., 2:3]], -1), torch.cat([state[..., :2] + 1.0, size, state[..., 3:4]], -1)
    for metric in (0, 1):
        assert torch.equal(torch.ops.tds_b200.collision_pairwise(b1, b2, metric), ops.collision_pairwise(b1, b2, metric))
    torch.library.opcheck(torch.ops.tds_b200.agent_boxes.default, (state, size), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(torch.ops.tds_b200.kinematic_step.default, (state, action, lr, None, 0, 0.1, True),
                          test_utils=("test_schema", "test_faketensor"))
