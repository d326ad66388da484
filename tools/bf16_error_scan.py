"""How the bf16-mode gradient error vs the fp64 oracle scales with batch size (lr = 0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import DrQV2Agent  # noqa: E402
from oracle import drq_oracle as O  # noqa: E402
from tests.helpers import rel_l2  # noqa: E402

SCHED = "linear(1.0,0.1,100000)"
A, Fd, H = 6, 50, 256
params = O.synthetic_params(9, A, Fd, H, seed=4)
for mode in ("fp32", "bf16"):
    for B in (16, 64, 192):
        agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 0.0, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True,
                           use_cuda_graph=False, seed=5, mode=mode)
        for net in ("encoder", "actor", "critic", "critic_target"):
            getattr(agent, net).load_state_dict(params[net])
        o64 = O.OracleAgent(params, 0.0, 0.01, SCHED, 0.3, dtype=torch.float64)
        b = O.synthetic_batch(B, A, seed=10)
        agent.inject_draws(b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
        m = agent.update(iter([(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"])]), 0)
        m64 = o64.update(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"], 0, b["shift_obs"],
                         b["shift_next"], b["eps_critic"], b["eps_actor"])
        keys = [("encoder", "convnet.0.weight"), ("encoder", "convnet.6.weight"), ("critic", "trunk.0.weight"),
                ("critic", "Q1.0.weight"), ("critic", "Q1.2.weight"), ("actor", "trunk.0.weight"), ("actor", "policy.4.weight")]
        errs = {f"{n}.{k}": rel_l2(dict(getattr(agent, n).named_parameters())[k].grad.cpu().numpy(), o64.grads[n][k].numpy()) for n, k in keys}
        print(mode, "B", B, "closs", abs(m["critic_loss"] - m64["critic_loss"]) / m64["critic_loss"],
              "aloss", abs(m["actor_loss"] - m64["actor_loss"]) / abs(m64["actor_loss"]),
              {k: float(f"{v:.2e}") for k, v in errs.items()}, flush=True)
