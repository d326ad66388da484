"""The reference-named stage API, DrQV2Agent.update_critic / update_actor on encoded features
(drqv2.py:177-228), driven on its own for several steps in both modes: every optimiser keeps its own Adam step
count (torch.optim.Adam does), update_actor leaves the target critic alone, detached features leave the
encoder alone."""
import numpy as np
import pytest
import torch

from oracle import drq_oracle as O
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
SCHED = "linear(1.0,0.1,100000)"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stage_api_steps_match_oracle(dev, mode, capsys):
    from drqv2_b200 import DrQV2Agent, utils
    from drqv2_b200.drqv2 import SCAL_OFF
    torch.set_num_threads(max(8, torch.get_num_threads()))
    A, Fd, H, B, lr, steps = 6, 50, 256, 16, 1e-4, 3
    params = O.synthetic_params(9, A, Fd, H, seed=4)
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", lr, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True, use_cuda_graph=False,
                       seed=5, mode=mode)
    for net in ("encoder", "actor", "critic", "critic_target"):
        getattr(agent, net).load_state_dict(params[net])
    agent.refresh()
    o = O.OracleAgent(params, lr, 0.01, SCHED, 0.3, dtype=torch.float64, operands="bf16" if mode == "bf16" else "exact")
    enc0 = {k: v.detach().clone() for k, v in agent.encoder.state_dict().items()}
    tgt0 = {k: v.detach().clone() for k, v in agent.critic_target.state_dict().items()}
    p0 = {net: {k: v.detach().clone() for k, v in o.p[net].items()} for net in ("critic", "actor")}
    b = O.synthetic_batch(B, A, seed=30)
    with torch.no_grad():
        feat, feat_next = o.encode(b["obs"], b["shift_obs"]), o.encode(b["next_obs"], b["shift_next"])
    f32 = lambda t: t.float().cuda()
    tol0 = 1e-4 if mode == "fp32" else 2e-3
    for s in range(steps):
        g = torch.Generator().manual_seed(100 + s)
        eps_c, eps_a = torch.randn(B, A, generator=g), torch.randn(B, A, generator=g)
        agent.inject_draws(None, None, eps_c, eps_a)
        mc = agent.update_critic(f32(feat), b["action"].cuda(), b["reward"].cuda(), b["discount"].cuda(), f32(feat_next), 2 * s)
        sc = agent._scal_dev.cpu().numpy()
        ma = agent.update_actor(f32(feat), 2 * s)
        sa = agent._scal_dev.cpu().numpy()
        oc = o.update_critic(feat, b["action"], b["reward"], b["discount"], feat_next, 2 * s, eps_c)
        oa = o.update_actor(feat, 2 * s, eps_a)
        # each optimiser's own step count reaches the device (the stage API used t = 1 forever)
        want = utils.adam_scalars(lr, s + 1)
        assert np.array_equal(sc[SCAL_OFF["critic"]:SCAL_OFF["critic"] + 6], want[:6])
        assert np.array_equal(sa[SCAL_OFF["actor"]:SCAL_OFF["actor"] + 6], want[:6])
        tol = tol0 if s == 0 else 2e-2            # after an Adam step everything carries its sign-like noise
        for k in oc:
            assert abs(mc[k] - oc[k]) <= tol * abs(oc[k]) + 1e-4, (s, k, mc[k], oc[k])
        for k in ("actor_loss", "actor_logprob", "actor_ent"):
            assert abs(ma[k] - oa[k]) <= max(tol, 5e-3) * abs(oa[k]) + 1e-4, (s, k, ma[k], oa[k])
    torch.cuda.synchronize()
    assert agent._opt_steps == dict(encoder=0, critic=steps, actor=steps)
    # detached features: the encoder did not move; update_actor did not touch the target
    for k, v in agent.encoder.state_dict().items():
        assert torch.equal(v, enc0[k]), k
    for k, v in agent.critic_target.state_dict().items():
        assert torch.equal(v, tgt0[k]), k
    # the accumulated parameter step of the large weight tensors follows the oracle's (a wrong bias correction
    # scales the second and third step by 1.4x / 1.6x)
    report = {}
    for net, names in (("critic", ("Q1.0.weight", "Q1.2.weight", "Q2.2.weight")), ("actor", ("policy.0.weight", "policy.2.weight"))):
        for name in names:
            got = dict(getattr(agent, net).named_parameters())[name].detach().cpu().double() - p0[net][name]
            want = o.p[net][name] - p0[net][name]
            report[f"{net}.{name}"] = rel_l2(got.numpy(), want.numpy())
            assert (got - want).abs().max().item() <= 2.5 * lr * steps
    with capsys.disabled():
        print(f"\nstage API ({mode}) parameter-step rel-L2 vs oracle after {steps} steps:", {k: float(f"{v:.2e}") for k, v in report.items()})
    for k, v in report.items():
        assert v <= 0.15, (k, v)


def test_stage_api_own_features_update_the_encoder(dev):
    """fp32 mode: update_critic on the agent's own feature buffer back-propagates into the encoder and steps
    encoder_opt, as autograd does inside the reference's update() (drqv2.py:198-202)."""
    from drqv2_b200 import DrQV2Agent
    A, Fd, H, B, lr = 6, 50, 128, 8, 1e-4
    params = O.synthetic_params(9, A, Fd, H, seed=7)
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", lr, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True, use_cuda_graph=False,
                       seed=5, mode="fp32")
    for net in ("encoder", "actor", "critic", "critic_target"):
        getattr(agent, net).load_state_dict(params[net])
    b = O.synthetic_batch(B, A, seed=31)
    ws = agent.workspace(B)
    ws.obs[:B].copy_(b["obs"])
    ws.obs[B:].copy_(b["next_obs"])
    ws.shift[:B].copy_(b["shift_obs"])
    ws.shift[B:].copy_(b["shift_next"])
    agent._encode(ws)
    agent.inject_draws(None, None, b["eps_critic"], b["eps_actor"])
    m = agent.update_critic(ws.feat[:B], b["action"].cuda(), b["reward"].cuda(), b["discount"].cuda(), ws.feat[B:], 0)
    o = O.OracleAgent(params, lr, 0.01, SCHED, 0.3, dtype=torch.float64)
    enc = {k: v.requires_grad_(True) for k, v in o.p["encoder"].items()}
    feat = o.encode(b["obs"], b["shift_obs"], enc)
    with torch.no_grad():
        feat_next = o.encode(b["next_obs"], b["shift_next"], enc)
    mo = o.update_critic(feat, b["action"], b["reward"], b["discount"], feat_next, 0, b["eps_critic"], enc=enc)
    for k in mo:
        assert abs(m[k] - mo[k]) <= 1e-4 * abs(mo[k]) + 1e-6, (k, m[k], mo[k])
    assert agent._opt_steps == dict(encoder=1, critic=1, actor=0)
    for name, p in agent.encoder.named_parameters():
        assert rel_l2(p.grad.cpu().numpy(), o.grads["encoder"][name].numpy()) <= 3e-3, name
        assert not torch.equal(p.detach().cpu(), params["encoder"][name]), name
