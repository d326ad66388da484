"""The callers either side of the update path: the actor/learner loop of the reference's
train.py:138-190 as a pipelined driver, and the evaluation loop of train.py:98-122 over a batch of environments.

The reference alternates  act -> update (blocking: 8 .item() reads) -> env.step -> storage.add  on one thread, so
the GPU idles while the CPU simulates and the CPU idles while the GPU updates.  Here the same steps run in the
same order with the same data dependencies, but

  * `update_async` only enqueues the update (one CUDA-graph replay); the environment step and
    `ReplayBufferStorage.add` (which ingests finished episodes into the HBM ring) run on the CPU while the GPU
    works, and the update's metrics are read one iteration later (`read_metrics`), when they are already there;
  * the next `act` is stream-ordered behind the update, so it sees the newest weights - exactly what the
    reference's synchronous loop produces (the parameters after N steps are bit-identical, tests/test_loop.py);
  * evaluation steps `n_envs` environments at once and calls `act` on the whole batch (the batch-1024 graph of
    BASELINE configs[2]) instead of once per environment step.

Environments are duck-typed as the reference's wrapped dm_env (dmc.py:20-33,180-210): `reset()` / `step(action)`
return a time step with `.observation` (uint8 [C, 84, 84]), `.reward`, `.last()` and item access by spec name.
hydra, the logger and the video recorders of train.py stay out of scope; `log` is any callable(dict, step, ty).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import utils


def _scalar(x):
    return 0.0 if x is None else float(np.asarray(x).reshape(-1)[0])


class Until:
    """utils.Until (reference utils.py:64-74)"""

    def __init__(self, until, action_repeat=1):
        self._until, self._action_repeat = until, action_repeat

    def __call__(self, step):
        if self._until is None:
            return True
        return step < self._until // self._action_repeat


class Every:
    """utils.Every (reference utils.py:77-87)"""

    def __init__(self, every, action_repeat=1):
        self._every, self._action_repeat = every, action_repeat

    def __call__(self, step):
        if self._every is None:
            return False
        return step % (self._every // self._action_repeat) == 0


class BatchedEnv:
    """`n` independent environments stepped together: observations stacked to uint8 [n, C, 84, 84] so that one
    `agent.act` call serves all of them.  An environment whose episode ended is reset on the next step and its
    `done` flag reported once."""

    def __init__(self, envs):
        self.envs = list(envs)
        if not self.envs:
            raise ValueError("BatchedEnv needs at least one environment")
        self._obs = None
        self._live = None

    def __len__(self):
        return len(self.envs)

    def reset(self):
        steps = [e.reset() for e in self.envs]
        self._obs = np.stack([np.asarray(ts.observation) for ts in steps])
        self._live = np.ones(len(self.envs), dtype=bool)
        return self._obs

    def step(self, actions):
        """actions float32 [n, A] -> (observations [n, C, 84, 84], rewards [n], done [n]).  Environments that are
        already done (not live) are not stepped; their reward is 0."""
        n = len(self.envs)
        rewards, done = np.zeros(n, np.float64), np.zeros(n, dtype=bool)
        for i, env in enumerate(self.envs):
            if not self._live[i]:
                continue
            ts = env.step(actions[i])
            self._obs[i] = np.asarray(ts.observation)
            rewards[i] = _scalar(ts.reward)
            if ts.last():
                done[i] = True
                self._live[i] = False
        return self._obs, rewards, done

    def revive(self, i):
        ts = self.envs[i].reset()
        self._obs[i] = np.asarray(ts.observation)
        self._live[i] = True

    @property
    def live(self):
        return self._live


def evaluate(agent, envs, num_episodes, global_step, action_repeat=1):
    """train.py:98-122 over a BatchedEnv: `num_episodes` evaluation episodes, the deterministic action
    (eval_mode=True) for every live environment from ONE batched `act` per step.  Returns the reference's eval
    record: mean episode reward and mean episode length in frames."""
    if not isinstance(envs, BatchedEnv):
        envs = BatchedEnv(envs)
    n = len(envs)
    started = min(n, num_episodes)
    obs = envs.reset()
    for i in range(started, n):                      # more environments than episodes asked for
        envs.live[i] = False
    finished, steps, total_reward = 0, 0, 0.0
    while finished < num_episodes:
        with torch.no_grad(), utils.eval_mode(agent):
            actions = agent.act(obs, global_step, eval_mode=True)        # [n, A]: the batched act graph
        live_before = envs.live.copy()
        obs, rewards, done = envs.step(actions)
        total_reward += float(rewards.sum())
        steps += int(live_before.sum())
        for i in np.nonzero(done)[0]:
            finished += 1
            if started < num_episodes:               # start the next episode in the freed slot
                envs.revive(i)
                started += 1
        if not envs.live.any() and finished < num_episodes:
            raise RuntimeError("evaluate: all environments finished before num_episodes was reached")
    return dict(episode_reward=total_reward / num_episodes, episode_length=steps * action_repeat / num_episodes,
                step=global_step)


class TrainLoop:
    """train.py:138-190 with the update pipelined against the environment step.

    agent          : DrQV2Agent (this package's)
    train_env      : the wrapped environment (dmc.make(...), train.py:58-59)
    replay_storage : ReplayBufferStorage; replay_iter: iter(make_replay_loader(...))
    pipelined=False runs the reference's blocking order (update() with its metrics read inside the step) - same
    arithmetic, used by the tests as the yardstick."""

    def __init__(self, agent, train_env, replay_storage, replay_iter, num_train_frames, num_seed_frames,
                 action_repeat=1, eval_every_frames=None, eval_fn=None, log=None, pipelined=True):
        self.agent, self.env = agent, train_env
        self.storage, self.replay_iter = replay_storage, replay_iter
        self.action_repeat = action_repeat
        self.train_until_step = Until(num_train_frames, action_repeat)
        self.seed_until_step = Until(num_seed_frames, action_repeat)
        self.eval_every_step = Every(eval_every_frames, action_repeat)
        self.eval_fn, self.log = eval_fn, log or (lambda metrics, step, ty: None)
        self.pipelined = pipelined
        self.global_step = 0
        self.global_episode = 0
        self.updates = 0
        self.overlapped = 0          # env steps that began while the update enqueued before them was still running
        self.timing = dict(act=0.0, update_enqueue=0.0, env=0.0, add=0.0, metrics=0.0)
        self._time_step, self._pending = None, None
        self._episode_step, self._episode_reward, self._t_episode = 0, 0.0, 0.0

    @property
    def global_frame(self):
        return self.global_step * self.action_repeat

    def run(self, steps=None):
        """Run until num_train_frames (train.py:156), or for at most `steps` further environment steps; a later
        call continues where this one stopped (same episode, same pending metrics)."""
        agent, t = self.agent, self.timing
        stop = None if steps is None else self.global_step + steps
        if self._time_step is None:
            self._time_step = self.env.reset()
            self.storage.add(self._time_step)
            self._t_episode = time.perf_counter()
        time_step, pending = self._time_step, self._pending
        while self.train_until_step(self.global_step) and (stop is None or self.global_step < stop):
            if time_step.last():
                self.global_episode += 1
                frames = self._episode_step * self.action_repeat
                self.log(dict(fps=frames / max(time.perf_counter() - self._t_episode, 1e-9),
                              episode_reward=self._episode_reward, episode_length=frames, episode=self.global_episode,
                              buffer_size=len(self.storage), step=self.global_step), self.global_frame, "train")
                time_step = self.env.reset()
                self.storage.add(time_step)
                self._episode_step, self._episode_reward, self._t_episode = 0, 0.0, time.perf_counter()
            if self.eval_fn is not None and self.eval_every_step(self.global_step):
                self.log(self.eval_fn(self.global_step), self.global_frame, "eval")
            # sample action (train.py:174-177): stream-ordered behind the last update, i.e. with its weights
            t0 = time.perf_counter()
            with torch.no_grad(), utils.eval_mode(agent):
                action = agent.act(time_step.observation, self.global_step, eval_mode=False)
            t1 = time.perf_counter()
            t["act"] += t1 - t0
            # the previous update finished before that act returned: its metrics are read without waiting
            if pending is not None:
                ws, at_step = pending
                if agent.use_tb:
                    self.log(agent.read_metrics(ws), at_step * self.action_repeat, "train")
                pending = None
            t2 = time.perf_counter()
            t["metrics"] += t2 - t1
            # try to update the agent (train.py:180-182)
            done_event = None
            if not self.seed_until_step(self.global_step):
                if self.pipelined:
                    ws = agent.update_async(self.replay_iter, self.global_step)
                    if ws is not None:
                        self.updates += 1
                        pending = (ws, self.global_step)
                        done_event = torch.cuda.Event()
                        done_event.record()
                else:
                    metrics = agent.update(self.replay_iter, self.global_step)
                    if self.global_step % agent.update_every_steps == 0:
                        self.updates += 1
                    self.log(metrics, self.global_frame, "train")
            t3 = time.perf_counter()
            t["update_enqueue"] += t3 - t2
            # take env step (train.py:185-188) - on the CPU, beside the update
            if done_event is not None and not done_event.query():
                self.overlapped += 1
            time_step = self.env.step(action)
            t4 = time.perf_counter()
            t["env"] += t4 - t3
            self._episode_reward += _scalar(time_step.reward)
            self.storage.add(time_step)
            t["add"] += time.perf_counter() - t4
            self._episode_step += 1
            self.global_step += 1
        self._time_step, self._pending = time_step, pending
        if not self.train_until_step(self.global_step):
            if pending is not None and agent.use_tb:
                self.log(agent.read_metrics(pending[0]), pending[1] * self.action_repeat, "train")
            self._pending = None
        torch.cuda.synchronize()
        return dict(steps=self.global_step, episodes=self.global_episode, updates=self.updates,
                    overlapped_env_steps=self.overlapped, timing=dict(t))
