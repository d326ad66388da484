"""Independent-agent ensembles on one GPU (BASELINE configs[3]).

The reference runs an ensemble as one hydra sweep job per seed (cfgs/config.yaml:4,55-62): every member
has its own parameters, optimiser state, replay buffer and RNG stream, and nothing is exchanged.  Here the
members of one GPU live in one process; each keeps its own CUDA graph and is replayed on its own stream.
A single batch-256 update leaves SMs idle in its launch gaps and in its small kernels (heads, LayerNorm,
gather, optimiser tails); with several members in flight those gaps are filled by another member's work:
measured 1574 -> 1940 (2 members) -> 2146 (4 members) aggregate updates/s on one B200.  Across GPUs the
ensemble is one process per GPU with no collective (bench.py --parallel ensemble)."""
from __future__ import annotations

import torch

from .drqv2 import DrQV2Agent


class AgentEnsemble:
    """`n_agents` DrQV2Agent members on one device.  `agent_args` / `agent_kw` are DrQV2Agent's constructor
    arguments (drqv2.py:128-133); member k gets seed `seeds[k]` (default base_seed + k)."""

    def __init__(self, n_agents, *agent_args, seeds=None, base_seed=0, **agent_kw):
        if n_agents < 1:
            raise ValueError("an ensemble needs at least one member")
        seeds = list(seeds) if seeds is not None else [base_seed + k for k in range(n_agents)]
        if len(seeds) != n_agents:
            raise ValueError("one seed per member")
        agent_kw.pop("seed", None)
        self.agents = []
        for k in range(n_agents):
            torch.manual_seed(seeds[k])                       # parameter init differs per member, as per seed in the reference
            self.agents.append(DrQV2Agent(*agent_args, seed=seeds[k], **agent_kw))
        dev = self.agents[0]._dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_agents)]

    def __len__(self):
        return len(self.agents)

    def __getitem__(self, k):
        return self.agents[k]

    def update(self, replay_iters, step):
        """One update of every member (member k draws from replay_iters[k]); returns the members' metrics
        dicts.  All updates are enqueued before any metric is read back, so they overlap on the device."""
        if len(replay_iters) != len(self.agents):
            raise ValueError("one replay iterator per member")
        cur = torch.cuda.current_stream()
        pending = []
        for ag, it, s in zip(self.agents, replay_iters, self.streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                pending.append(ag.update_async(it, step))
        out = []
        for ag, ws, s in zip(self.agents, pending, self.streams):
            if ws is not None and ag.use_tb:
                with torch.cuda.stream(s):
                    out.append(ag.read_metrics(ws))
            else:
                out.append(dict())
            cur.wait_stream(s)
        return out

    def act(self, observations, step, eval_mode):
        """Member k acts on observations[k] (DrQV2Agent.act, drqv2.py:159-175)."""
        return [ag.act(o, step, eval_mode) for ag, o in zip(self.agents, observations)]

    def train(self, training=True):
        for ag in self.agents:
            ag.train(training)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()
