"""the reference's train.py loop (train.py:138-190) as the pipelined TrainLoop, and its
eval loop (train.py:98-122) over a batch of environments - driven with a fake pixel environment (no MuJoCo here)."""
import numpy as np
import pytest
import torch

from tests.fake_env import FakePixelEnv, data_specs

SCHED = "linear(1.0,0.1,100000)"


class _StubAgent:
    """act() contract only: deterministic actions from the observation batch (CPU)."""
    training = True
    use_tb = False

    def __init__(self, A):
        self.A, self.calls = A, []

    def train(self, training=True):
        self.training = training

    def act(self, obs, step, eval_mode):
        obs = np.asarray(obs)
        batched = obs.ndim == 4
        o = obs if batched else obs[None]
        self.calls.append(o.shape[0])
        a = np.tanh((o.reshape(o.shape[0], -1)[:, :self.A].astype(np.float32) - 128.0) / 128.0)
        return a if batched else a[0]


def test_batched_eval_equals_serial_eval_cpu():
    """evaluate() over a BatchedEnv gives the reference's serial eval record (train.py:98-122) and serves every
    step of all live environments with one act call."""
    from drqv2_b200.loop import BatchedEnv, evaluate
    A, T, n, episodes = 4, 7, 5, 12
    agent = _StubAgent(A)
    rec = evaluate(agent, BatchedEnv([FakePixelEnv(A, T, seed=i) for i in range(n)]), episodes, 100, action_repeat=2)
    assert rec["episode_length"] == T * 2 and rec["step"] == 100
    assert max(agent.calls) == n and len(agent.calls) < episodes * T        # batched calls, not one per env step
    # serial yardstick: the reference's loop, one environment at a time (each env runs its episodes in order)
    envs = [FakePixelEnv(A, T, seed=i) for i in range(n)]
    serial = _StubAgent(A)
    total, count = 0.0, 0
    per_env = [len(range(i, episodes, n)) for i in range(n)]                # slot i is revived while episodes remain
    for env, k in zip(envs, per_env):
        for _ in range(k):
            ts = env.reset()
            while not ts.last():
                ts = env.step(serial.act(ts.observation, 100, True))
                total += float(ts.reward[0])
            count += 1
    assert count == episodes
    assert rec["episode_reward"] == pytest.approx(total / episodes, rel=1e-6)


def _make(tmp_path, tag, mode, B, H, T, A=6, step_seconds=0.0, use_tb=True):
    from drqv2_b200 import DrQV2Agent, ReplayBufferStorage, make_replay_loader
    torch.manual_seed(11)
    np.random.seed(11)
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, 50, H, 0.01, 20, 2, SCHED, 0.3, use_tb, seed=3, mode=mode)
    storage = ReplayBufferStorage(data_specs(A), tmp_path / tag)
    loader = make_replay_loader(tmp_path / tag, 4096, B, 0, False, 3, 0.99)
    return agent, FakePixelEnv(A, T, seed=5, step_seconds=step_seconds), storage, loader


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_pipelined_loop_equals_blocking_loop(tmp_path, mode):
    """The pipelined driver reorders no arithmetic: after the same number of steps the parameters equal the
    reference-ordered blocking loop's bit for bit (every action depends on the weights of the update before it,
    and every frame on the action)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from drqv2_b200.loop import TrainLoop
    agents, logs = [], []
    for pipelined in (False, True):
        agent, env, storage, loader = _make(tmp_path, f"buf{int(pipelined)}", mode, B=8, H=64, T=12)
        log = []
        loop = TrainLoop(agent, env, storage, iter(loader), num_train_frames=90, num_seed_frames=30, action_repeat=1,
                         log=lambda m, step, ty, log=log: log.append((ty, step, dict(m))), pipelined=pipelined)
        out = loop.run()
        assert out["steps"] == 90 and out["updates"] == 30           # steps 30..88, every second step
        assert out["episodes"] == 7                                  # 12-step episodes
        agents.append(agent)
        logs.append([(ty, step, m) for ty, step, m in log if ty == "train" and "critic_loss" in m])
    for net in ("encoder", "actor", "critic", "critic_target"):
        for (n1, p1), (_, p2) in zip(getattr(agents[0], net).named_parameters(), getattr(agents[1], net).named_parameters()):
            assert torch.equal(p1, p2), (net, n1)
    # the same metrics, logged at the same frames (the pipelined loop reads them one iteration later)
    assert len(logs[0]) == len(logs[1]) == 30
    for (_, s0, m0), (_, s1, m1) in zip(*logs):
        assert s0 == s1 and m0 == m1


@pytest.mark.gpu
def test_env_step_overlaps_the_update(tmp_path, capsys):
    """At the BENCH configuration (B=256, H=1024, bf16) the environment step starts while the update enqueued
    before it is still running on the GPU, and an iteration costs less than in the blocking order."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import time
    from drqv2_b200.loop import TrainLoop
    report = {}
    for pipelined in (False, True):
        agent, env, storage, loader = _make(tmp_path, f"ov{int(pipelined)}", "bf16", B=256, H=1024, T=50,
                                            step_seconds=0.6e-3, use_tb=True)
        agent.update_every_steps = 1
        loop = TrainLoop(agent, env, storage, iter(loader), num_train_frames=460, num_seed_frames=110,
                         pipelined=pipelined)
        loop.run(steps=160)                          # seed episodes, the eager pass, the capture, first replays
        loop.overlapped, u0, s0 = 0, loop.updates, loop.global_step
        t0 = time.perf_counter()
        out = loop.run()
        dt = time.perf_counter() - t0
        report[pipelined] = dict(ms_per_step=1e3 * dt / (out["steps"] - s0), overlapped=out["overlapped_env_steps"],
                                 updates=out["updates"] - u0)
    with capsys.disabled():
        print("\ntrain loop, fake env 0.6 ms/step, B=256 bf16 update every step:", report)
    assert report[True]["updates"] == 300
    assert report[True]["overlapped"] >= 0.8 * report[True]["updates"]
    assert report[True]["ms_per_step"] < 0.9 * report[False]["ms_per_step"]
