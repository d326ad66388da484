"""bf16-mode gradient error vs fp64 oracle on smooth (natural-image-like) frames, all encoder layers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import DrQV2Agent  # noqa: E402
from oracle import drq_oracle as O  # noqa: E402
from tests.helpers import rel_l2  # noqa: E402

SCHED = "linear(1.0,0.1,100000)"
A, Fd, H, B = 6, 50, 256, 64
params = O.synthetic_params(9, A, Fd, H, seed=4)
for kind in ("noise", "smooth"):
    b = O.synthetic_batch(B, A, seed=10)
    if kind == "smooth":
        g = torch.Generator().manual_seed(3)
        def smooth():
            low = torch.rand(B, 9, 12, 12, generator=g)
            up = torch.nn.functional.interpolate(low, size=(84, 84), mode="bicubic", align_corners=False)
            return (up.clamp(0, 1) * 255).round().to(torch.uint8)
        b["obs"], b["next_obs"] = smooth(), smooth()
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 0.0, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True,
                       use_cuda_graph=False, seed=5, mode="bf16")
    for net in ("encoder", "actor", "critic", "critic_target"):
        getattr(agent, net).load_state_dict(params[net])
    o64 = O.OracleAgent(params, 0.0, 0.01, SCHED, 0.3, dtype=torch.float64)
    agent.inject_draws(b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
    m = agent.update(iter([(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"])]), 0)
    m64 = o64.update(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"], 0, b["shift_obs"],
                     b["shift_next"], b["eps_critic"], b["eps_actor"])
    errs = {}
    for n in ("encoder", "critic", "actor"):
        for k, p in getattr(agent, n).named_parameters():
            if k.endswith("weight"):
                errs[f"{n}.{k}"] = rel_l2(p.grad.cpu().numpy(), o64.grads[n][k].numpy())
    print(kind, {k: float(f"{v:.2e}") for k, v in errs.items()}, flush=True)
