"""drqv2_b200 — the DrQ-v2 agent-update hot path as hand-written sm_100a CUDA kernels
behind the reference's Python API (drqv2.py / replay_buffer.py / utils.py)."""
from . import utils  # noqa: F401
from .drqv2 import Actor, Critic, DrQV2Agent, Encoder, RandomShiftsAug  # noqa: F401
from .ensemble import AgentEnsemble  # noqa: F401
from .loop import BatchedEnv, TrainLoop, evaluate  # noqa: F401
from .replay_buffer import ReplayBufferStorage, make_replay_loader  # noqa: F401

__all__ = ["DrQV2Agent", "RandomShiftsAug", "Encoder", "Actor", "Critic", "ReplayBufferStorage",
           "make_replay_loader", "utils", "AgentEnsemble", "TrainLoop", "BatchedEnv", "evaluate"]
