"""A deterministic stand-in for the reference's wrapped DMC environment (dmc.py:180-210: action repeat, pixels,
frame stack of 3, extended time step) - there is no MuJoCo here.  Observations are uint8 [9, 84, 84] stacks of the
last three frames with the reset frame repeated (dmc.py:98-109); frames depend on the actions taken, so a training
loop's parameters depend on every action it chose."""
import time
from collections import deque, namedtuple

import numpy as np

Spec = namedtuple("Spec", "shape dtype name")


def data_specs(A):
    """train.py:49-53"""
    return (Spec((9, 84, 84), np.uint8, "observation"), Spec((A,), np.float32, "action"),
            Spec((1,), np.float32, "reward"), Spec((1,), np.float32, "discount"))


class TimeStep:
    """dmc.ExtendedTimeStep (dmc.py:14-32): attribute and item access, first()/last()"""

    def __init__(self, step_type, reward, discount, observation, action):
        self.step_type, self.reward, self.discount = step_type, reward, discount
        self.observation, self.action = observation, action

    def first(self):
        return self.step_type == 0

    def last(self):
        return self.step_type == 2

    def __getitem__(self, attr):
        return getattr(self, attr)


class FakePixelEnv:
    def __init__(self, A, episode_len, seed=0, step_seconds=0.0):
        self.A, self.T, self.seed, self.step_seconds = A, episode_len, seed, step_seconds
        self.episode = -1
        self._base = (np.arange(3 * 84 * 84, dtype=np.int64).reshape(3, 84, 84) * 7) % 251

    def _frame(self):
        return ((self._base * (1 + self.t % 5) + 31 * self.episode + 17 * self.t + self.seed + self._drift) % 256).astype(np.uint8)

    def reset(self):
        self.episode += 1
        self.t, self._drift = 0, 0
        self._frames = deque([self._frame()] * 3, maxlen=3)
        return TimeStep(0, np.zeros(1, np.float32), np.ones(1, np.float32), np.concatenate(list(self._frames), 0),
                        np.zeros(self.A, np.float32))

    def step(self, action):
        if self.step_seconds:
            end = time.perf_counter() + self.step_seconds      # a busy CPU, as a physics step is
            while time.perf_counter() < end:
                pass
        action = np.asarray(action, np.float32)
        assert action.shape == (self.A,) and np.all(np.abs(action) <= 1.0)
        self.t += 1
        self._drift = int(np.round(1000 * float(action.sum()))) % 256
        self._frames.append(self._frame())
        reward = np.float32(1.0 - min(1.0, float(np.abs(action).mean())))
        return TimeStep(2 if self.t == self.T else 1, np.full(1, reward, np.float32), np.ones(1, np.float32),
                        np.concatenate(list(self._frames), 0), action)
