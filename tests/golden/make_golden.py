"""Generate the committed golden vectors by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the outputs
(tests/golden/*.npz, *.json) are committed so the tests run anywhere.

    python tests/golden/make_golden.py

The reference (drqv2.py / utils.py / replay_buffer.py / dmc.py) is imported with empty
stub modules for the packages it imports but does not use on this path (hydra,
omegaconf, dm_env, dm_control) — SURVEY.md §8c.  Random draws are injected by
patching torch.randint / utils._standard_normal / np.random.randint exactly where the
reference calls them, so that the same draws can be fed to the oracle and the CUDA path.
"""
import collections
import enum
import hashlib
import json
import os
import pathlib
import sys
import tempfile
import types

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = os.environ.get("DRQ_REFERENCE", "/root/reference")
sys.path.insert(0, str(ROOT))


def import_reference():
    for n in ("hydra", "omegaconf"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["omegaconf"].OmegaConf = object
    # minimal dm_env / dm_control stubs so that dmc.py's wrappers import
    dm_env = types.ModuleType("dm_env")

    class StepType(enum.IntEnum):
        FIRST = 0
        MID = 1
        LAST = 2

    class TimeStep(collections.namedtuple("TimeStep", "step_type reward discount observation")):
        def first(self):
            return self.step_type == StepType.FIRST

        def last(self):
            return self.step_type == StepType.LAST

    class Environment:
        pass

    specs = types.ModuleType("dm_env.specs")

    class Array:
        def __init__(self, shape, dtype, name=None):
            self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

    class BoundedArray(Array):
        def __init__(self, shape, dtype, minimum, maximum, name=None):
            super().__init__(shape, dtype, name)
            self.minimum, self.maximum = minimum, maximum

    specs.Array, specs.BoundedArray = Array, BoundedArray
    dm_env.StepType, dm_env.TimeStep, dm_env.Environment, dm_env.specs = StepType, TimeStep, Environment, specs
    sys.modules["dm_env"], sys.modules["dm_env.specs"] = dm_env, specs
    for n in ("dm_control", "dm_control.suite", "dm_control.manipulation", "dm_control.suite.wrappers",
              "dm_control.suite.wrappers.action_scale", "dm_control.suite.wrappers.pixels"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["dm_control"].manipulation = sys.modules["dm_control.manipulation"]
    sys.modules["dm_control"].suite = sys.modules["dm_control.suite"]
    sys.modules["dm_control.suite"].wrappers = sys.modules["dm_control.suite.wrappers"]
    sys.modules["dm_control.suite.wrappers"].action_scale = sys.modules["dm_control.suite.wrappers.action_scale"]
    sys.modules["dm_control.suite.wrappers"].pixels = sys.modules["dm_control.suite.wrappers.pixels"]
    sys.path.insert(0, REF)
    import drqv2, utils, replay_buffer, dmc  # noqa: E401
    return drqv2, utils, replay_buffer, dmc, dm_env


def frame_formula(e, t):
    """Deterministic u8 frame [3,84,84] for episode e, row t (integer arithmetic only)."""
    c = np.arange(3).reshape(3, 1, 1)
    y = np.arange(84).reshape(1, 84, 1)
    x = np.arange(84).reshape(1, 1, 84)
    return ((31 * e + 17 * t + 7 * c + 3 * y + 5 * x + (x * y) % 11 + (e + 1) * (t + 2) * (x + y) % 13) % 256).astype(np.uint8)


def scalar_formula(e, t, A):
    action = (((np.arange(A) * 37 + e * 11 + t * 5) % 200) / 100.0 - 1.0).astype(np.float32)
    reward = np.float32(((e * 7 + t * 13) % 97) / 97.0)
    discount = np.float32(1.0 if (t + e) % 5 else 0.9)   # exercise the discount chain
    return action, reward, discount


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_replay_golden(replay_buffer, dmc, dm_env):
    """Fake pixel env -> reference FrameStackWrapper -> ExtendedTimeStepWrapper ->
    ReplayBufferStorage -> npz -> ReplayBuffer._sample (replay_buffer.py:142-160)."""
    A = 6
    lens = [9, 5, 12]   # transitions per episode (rows = len + 1)

    class FakeEnv(dm_env.Environment):
        def __init__(self):
            self.e, self.t = -1, 0

        def _ts(self, step_type):
            _, rew, disc = scalar_formula(self.e, self.t, A)
            obs = collections.OrderedDict(pixels=frame_formula(self.e, self.t).transpose(1, 2, 0).copy())
            if step_type == dm_env.StepType.FIRST:
                return dm_env.TimeStep(step_type, None, None, obs)
            return dm_env.TimeStep(step_type, rew, disc, obs)

        def reset(self):
            self.e += 1
            self.t = 0
            return self._ts(dm_env.StepType.FIRST)

        def step(self, action):
            self.t += 1
            last = self.t == lens[self.e]
            return self._ts(dm_env.StepType.LAST if last else dm_env.StepType.MID)

        def observation_spec(self):
            return collections.OrderedDict(pixels=dm_env.specs.Array((84, 84, 3), np.uint8, "pixels"))

        def action_spec(self):
            return dm_env.specs.BoundedArray((A,), np.float32, -1.0, 1.0, "action")

    env = dmc.FrameStackWrapper(FakeEnv(), 3, "pixels")
    env = dmc.ExtendedTimeStepWrapper(env)
    specs = (env.observation_spec(), env.action_spec(),
             dm_env.specs.Array((1,), np.float32, "reward"), dm_env.specs.Array((1,), np.float32, "discount"))
    out = {"lens": lens, "A": A, "samples": []}
    with tempfile.TemporaryDirectory() as d:
        d = pathlib.Path(d)
        storage = replay_buffer.ReplayBufferStorage(specs, d / "buffer")
        for e in range(len(lens)):
            ts = env.reset()
            storage.add(ts)
            while not ts.last():
                a, _, _ = scalar_formula(e, ts_t(env) + 1, A)
                ts = env.step(a)
                storage.add(ts)
        assert len(storage) == sum(lens)
        files = sorted((d / "buffer").glob("*.npz"), key=lambda f: int(f.stem.split("_")[1]))
        episodes = [replay_buffer.load_episode(f) for f in files]
        # one episode file exactly as the reference's save_episode wrote it (replay_buffer.py:22-27): the fixture of
        # the npz interoperability tests ({ts}_{idx}_{len}.npz; the timestamp is dropped from the fixture's name)
        import shutil
        shutil.copyfile(files[1], HERE / "ref_episode_1_5.npz")
        for nstep in (1, 3):
            rb = replay_buffer.ReplayBuffer(d / "buffer", 10 ** 6, 0, nstep, 0.99, fetch_every=10 ** 9,
                                            save_snapshot=True)
            rb._samples_since_last_fetch = 0   # skip _try_fetch: episodes injected below
            for e, ep in enumerate(episodes):
                T = replay_buffer.episode_len(ep)
                for idx in range(1, T - nstep + 2):
                    rb._sample_episode = lambda ep=ep: ep
                    orig = np.random.randint
                    np.random.randint = lambda lo, hi, idx=idx: idx - 1   # replay_buffer.py:150
                    try:
                        obs, action, reward, discount, nxt = rb._sample()
                    finally:
                        np.random.randint = orig
                    out["samples"].append(dict(
                        episode=e, idx=idx, nstep=nstep, obs_sha=sha(obs), next_sha=sha(nxt),
                        action=[float(v) for v in action], reward=float(reward[0]),
                        discount=float(discount[0]),
                        reward_hex=np.float32(reward[0]).tobytes().hex(),
                        discount_hex=np.float32(discount[0]).tobytes().hex()))
    return out


def ts_t(env):
    return env._env._env.t   # ExtendedTimeStepWrapper -> FrameStackWrapper -> FakeEnv


def aug_input(N, C):
    n = np.arange(N).reshape(N, 1, 1, 1)
    c = np.arange(C).reshape(1, C, 1, 1)
    y = np.arange(84).reshape(1, 1, 84, 1)
    x = np.arange(84).reshape(1, 1, 1, 84)
    return ((n * 53 + c * 29 + y * y * 3 + x * 7 + (x * y) % 17) % 256).astype(np.uint8)


def make_aug_golden(drqv2):
    N, C = 3, 2
    x = torch.from_numpy(aug_input(N, C)).float()
    shifts = torch.tensor([[0, 8], [4, 4], [7, 1]], dtype=torch.int32)
    orig = torch.randint

    def fake_randint(lo, hi, size, device=None, dtype=None):
        assert (lo, hi) == (0, 9) and tuple(size) == (N, 1, 1, 2)
        return shifts.view(N, 1, 1, 2).to(dtype)

    torch.randint = fake_randint
    try:
        y = drqv2.RandomShiftsAug(pad=4)(x)
    finally:
        torch.randint = orig
    return dict(shifts=shifts.numpy(), out=y.numpy())


class ExactShift(torch.nn.Module):
    """REF-X: agent.aug replaced by the integer shift, consuming the same randint draw."""

    def __init__(self, pad):
        super().__init__()
        self.pad = pad

    def forward(self, x):
        n, c, h, w = x.size()
        shift = torch.randint(0, 2 * self.pad + 1, size=(n, 1, 1, 2), device=x.device, dtype=x.dtype)
        xp = torch.nn.functional.pad(x, (self.pad,) * 4, "replicate")
        out = torch.empty_like(x)
        for i in range(n):
            sx, sy = int(shift[i, 0, 0, 0]), int(shift[i, 0, 0, 1])
            out[i] = xp[i, :, sy:sy + h, sx:sx + w]
        return out


def probe_indices(numel, k=24):
    """Fixed pseudo-random probe positions (integer LCG; machine independent)."""
    idx, s = [], 12345
    for _ in range(min(k, numel)):
        s = (s * 1103515245 + 12345) % (2 ** 31)
        idx.append(s % numel)
    return np.array(sorted(set(idx)), dtype=np.int64)


def summarize(t):
    t = t.detach().double().flatten()
    pi = probe_indices(t.numel())
    return dict(l2=float(t.norm()), sum=float(t.sum()), absmax=float(t.abs().max()),
                probe_idx=pi.tolist(), probe=[float(v) for v in t[pi]])


def make_update_golden(drqv2, utils, case):
    from oracle import drq_oracle as O
    B, A, Fd, H, lr, steps = case["B"], case["A"], case["F"], case["H"], case["lr"], case["steps"]
    out = dict(case=case, variants={})
    for variant in ("ref", "refx"):
        torch.manual_seed(0)
        agent = drqv2.DrQV2Agent((9, 84, 84), (A,), "cpu", lr, Fd, H, 0.01, 2000, 2, "linear(1.0,0.1,100000)", 0.3, True)
        params = O.synthetic_params(9, A, Fd, H, seed=case["pseed"])
        agent.encoder.load_state_dict(params["encoder"])
        agent.actor.load_state_dict(params["actor"])
        agent.critic.load_state_dict(params["critic"])
        agent.critic_target.load_state_dict(params["critic_target"])
        if variant == "refx":
            agent.aug = ExactShift(4)
        per_step = []
        for s in range(steps):
            batch = O.synthetic_batch(B, A, seed=case["bseed"] + s)
            draws_int = [batch["shift_obs"], batch["shift_next"]]
            draws_f = [batch["eps_critic"], batch["eps_actor"]]
            orig_randint, orig_sn = torch.randint, utils._standard_normal

            def fake_randint(lo, hi, size, device=None, dtype=None):
                return draws_int.pop(0).view(tuple(size)).to(dtype)

            def fake_sn(shape, dtype, device):
                return draws_f.pop(0).view(tuple(shape)).to(dtype).clone()

            torch.randint, utils._standard_normal = fake_randint, fake_sn
            captured = {}
            orig_zero = agent.actor_opt.zero_grad

            def grab_and_zero(set_to_none=True):
                # drqv2.py:219 runs before the actor backward: encoder/critic grads of this
                # update are still alive here
                captured["encoder"] = {k: p.grad.clone() for k, p in agent.encoder.named_parameters()}
                captured["critic"] = {k: p.grad.clone() for k, p in agent.critic.named_parameters()}
                orig_zero(set_to_none=set_to_none)

            agent.actor_opt.zero_grad = grab_and_zero
            try:
                it = iter([(batch["obs"].numpy(), batch["action"].numpy(), batch["reward"].numpy(),
                            batch["discount"].numpy(), batch["next_obs"].numpy())])
                metrics = agent.update(it, 2 * s)
            finally:
                torch.randint, utils._standard_normal = orig_randint, orig_sn
                agent.actor_opt.zero_grad = orig_zero
            assert not draws_int and not draws_f
            captured["actor"] = {k: p.grad.clone() for k, p in agent.actor.named_parameters()}
            rec = dict(metrics={k: float(v) for k, v in metrics.items()},
                       grads={net: {k: summarize(g) for k, g in gs.items()} for net, gs in captured.items()},
                       params={net: {k: summarize(p) for k, p in getattr(agent, net).named_parameters()}
                               for net in ("encoder", "actor", "critic", "critic_target")})
            per_step.append(rec)
        # act(): eval-mode mean and a train-mode sample with injected noise (drqv2.py:164-175)
        obs1 = O.synthetic_batch(1, A, seed=99)["obs"][0].numpy()
        with torch.no_grad():
            a_eval = agent.act(obs1, 5000, True)
            eps = torch.linspace(-1.5, 1.5, A).view(1, A)
            orig_sn = utils._standard_normal
            utils._standard_normal = lambda shape, dtype, device: eps.clone().to(dtype)
            try:
                a_train = agent.act(obs1, 5000, False)
            finally:
                utils._standard_normal = orig_sn
        out["variants"][variant] = dict(steps=per_step, act_eval=[float(v) for v in a_eval],
                                        act_train=[float(v) for v in a_train])
    return out


def main():
    torch.set_num_threads(1)   # fixed summation order for the recorded floats
    drqv2, utils, replay_buffer, dmc, dm_env = import_reference()
    rep = make_replay_golden(replay_buffer, dmc, dm_env)
    (HERE / "replay_golden.json").write_text(json.dumps(rep, indent=0))
    aug = make_aug_golden(drqv2)
    np.savez_compressed(HERE / "aug_golden.npz", **aug)
    cases = [
        dict(name="walker_small", B=8, A=6, F=50, H=1024, lr=1e-4, steps=2, pseed=0, bseed=100),
        dict(name="humanoid_small", B=4, A=21, F=100, H=256, lr=8e-5, steps=1, pseed=3, bseed=200),
    ]
    upd = [make_update_golden(drqv2, utils, c) for c in cases]
    (HERE / "update_golden.json").write_text(json.dumps(upd))
    sched = {s: [utils.schedule(s, st) for st in (0, 1, 50000, 100000, 250000)]
             for s in ("linear(1.0,0.1,100000)", "linear(1.0,0.1,500000)", "0.2",
                       "step_linear(1.0,0.5,1000,0.1,200000)")}
    (HERE / "schedule_golden.json").write_text(json.dumps(sched))
    print("golden written:", [p.name for p in HERE.iterdir()])


if __name__ == "__main__":
    main()
