"""GPU-resident uint8 replay ring behind the reference's replay_buffer.py contract.

``ReplayBufferStorage(data_specs, replay_dir).add(time_step)`` /
``make_replay_loader(replay_dir, max_size, batch_size, num_workers, save_snapshot,
nstep, discount)`` keep the reference's signatures (replay_buffer.py:37-61,173-190), so
``train.py`` is unchanged.  Instead of compressed ``*.npz`` episodes on disk that four
worker processes re-read, episodes go into one device ring of single frames
(21,168 B per step instead of a 63,504 B stack), and a batch is produced by two kernels:
``drq_ring_sample`` (episode, then start index, as replay_buffer.py:96-98,150) and
``drq_ring_gather_nstep`` (frame stack by index + the bit-exact fp32 n-step chain of
replay_buffer.py:154-159).  Storage and loader meet through ``replay_dir`` exactly as
the reference's do — here it keys a registry of rings instead of naming a directory of
files.

Differences from the reference's sampler that are statistical, not arithmetic
(SURVEY appendix): sampling is uniform over all resident episodes (not per worker
quarter), new episodes are visible immediately, eviction is FIFO by whole episode.
"""
from __future__ import annotations

import datetime
import io
import pathlib
from collections import defaultdict

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import IMG, call

_RINGS = {}
DEFAULT_CAPACITY = 100_000


def _stream():
    return torch.cuda.current_stream().cuda_stream


def episode_len(episode):
    # rows minus the dummy first transition (replay_buffer.py:17-19)
    return next(iter(episode.values())).shape[0] - 1


# ----------------------------------------------------------------------------- on-disk episode format
# The reference keeps every episode as `{timestamp}_{index}_{length}.npz` (np.savez_compressed of the spec
# arrays, observations as full frame stacks; replay_buffer.py:22-34,71-78).  The ring stores single frames;
# these helpers convert both ways so that an existing reference buffer directory can be loaded and a
# `save_snapshot` run leaves the same files behind.
def save_episode(episode, fn):
    """np.savez_compressed of the episode dict, byte-compatible with replay_buffer.py:22-27."""
    with io.BytesIO() as bs:
        np.savez_compressed(bs, **episode)
        bs.seek(0)
        with pathlib.Path(fn).open("wb") as f:
            f.write(bs.read())


def load_episode(fn):
    """replay_buffer.py:30-34."""
    with pathlib.Path(fn).open("rb") as f:
        episode = np.load(f)
        return {k: episode[k] for k in episode.keys()}


def frames_from_stacks(obs, stack):
    """Stacked observations u8 [rows, stack*C, H, W] (dmc.py:86-109: row t = frames max(t-stack+1..t, 0)) ->
    the frame rendered at each row, u8 [rows, C, H, W].  Raises if obs is not such a stack."""
    rows, SC = obs.shape[0], obs.shape[1]
    if SC % stack:
        raise ValueError("observation channels must be a multiple of frame_stack")
    C = SC // stack
    newest = obs[:, (stack - 1) * C:]
    for j in range(stack - 1):
        lag = stack - 1 - j
        want = newest[np.maximum(np.arange(rows) - lag, 0)]
        if not np.array_equal(obs[:, j * C:(j + 1) * C], want):
            raise ValueError("observations are not a frame stack of consecutive frames; "
                             "construct ReplayBufferStorage(..., frame_stack=1) to store them whole")
    return np.ascontiguousarray(newest)


def stacks_from_frames(frames, stack):
    """Inverse of frames_from_stacks: u8 [rows, C, H, W] -> u8 [rows, stack*C, H, W]."""
    rows = frames.shape[0]
    t = np.arange(rows)
    return np.concatenate([frames[np.maximum(t - (stack - 1 - j), 0)] for j in range(stack)], axis=1)


def episode_files(replay_dir):
    """Episode files of a reference buffer directory in the order the reference's loader keeps them
    (sorted by name = by timestamp, replay_buffer.py:112)."""
    return sorted(pathlib.Path(replay_dir).glob("*.npz"))


class GpuRing:
    """Device ring of single frames + per-step scalars, with an episode table."""

    def __init__(self, capacity, frame_c, stack, action_dim, device="cuda", max_episodes=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GpuRing needs a CUDA device (drqv2_b200 has no CPU fallback)")
        _lib.lib()
        self.capacity, self.frame_c, self.stack, self.A = int(capacity), frame_c, stack, action_dim
        self.device = torch.device(device)
        dev = self.device
        self.frames = torch.zeros(self.capacity, frame_c, IMG, IMG, dtype=torch.uint8, device=dev)
        self.action = torch.zeros(self.capacity, action_dim, device=dev)
        self.reward = torch.zeros(self.capacity, device=dev)
        self.discount = torch.zeros(self.capacity, device=dev)
        self.max_episodes = max_episodes or max(16, self.capacity // 2)
        self.ep_table = torch.zeros(self.max_episodes, 2, dtype=torch.int32, device=dev)
        self.n_episodes = torch.zeros(1, dtype=torch.int32, device=dev)
        self.head = 0
        self.episodes = []          # host mirror: (start slot, rows)
        self.min_len = 1

    def __len__(self):
        return sum(rows - 1 for _, rows in self.episodes)

    def add_episode(self, frames, action, reward, discount):
        """frames u8 [rows, frame_c, 84, 84]; action f32 [rows, A]; reward/discount f32 [rows]."""
        rows = frames.shape[0]
        if rows > self.capacity:
            raise ValueError(f"episode of {rows} rows does not fit a ring of {self.capacity} slots")
        start = self.head
        # evict (FIFO, whole episodes) everything the new rows overwrite — replay_buffer.py:106-110
        def overlaps(s, r):
            a0, a1 = start, start + rows
            for b0 in (s, s + self.capacity, s - self.capacity):
                if b0 < a1 and a0 < b0 + r:
                    return True
            return False
        self.episodes = [(s, r) for (s, r) in self.episodes if not overlaps(s, r)]
        first = min(rows, self.capacity - start)
        parts = ((self.frames, torch.as_tensor(frames)), (self.action, torch.as_tensor(action)),
                 (self.reward, torch.as_tensor(reward).reshape(-1)),
                 (self.discount, torch.as_tensor(discount).reshape(-1)))
        for dst, src in parts:
            dst[start:start + first].copy_(src[:first], non_blocking=True)
            if first < rows:
                dst[:rows - first].copy_(src[first:], non_blocking=True)
        self.head = (start + rows) % self.capacity
        self.episodes.append((start, rows))
        self._upload_table()

    def _upload_table(self):
        eligible = [(s, r - 1) for s, r in self.episodes if r - 1 >= self.min_len][-self.max_episodes:]
        n = len(eligible)
        if n:
            t = torch.tensor(eligible, dtype=torch.int32)
            self.ep_table[:n].copy_(t, non_blocking=False)
        self.n_episodes.fill_(n)
        self._n_eligible = n


class ReplayBufferStorage:
    """Accumulates one episode on the host and pushes it into the GPU ring on
    ``time_step.last()`` (the reference writes an npz at the same moment,
    replay_buffer.py:48-61,71-78)."""

    def __init__(self, data_specs, replay_dir, frame_stack=3, device="cuda"):
        self._data_specs = data_specs
        self._replay_dir = pathlib.Path(replay_dir)
        self._frame_stack = frame_stack
        self._device = device
        self._current_episode = defaultdict(list)
        self._num_episodes = 0
        self._num_transitions = 0
        self._key = str(self._replay_dir)
        entry = _RINGS.setdefault(self._key, dict(ring=None, capacity=None, storage=self))
        entry["storage"] = self
        entry.setdefault("save_snapshot", False)
        self._pending = []
        self._preload()

    def _preload(self):
        """Episodes a previous (reference or drqv2_b200, save_snapshot) run left in replay_dir
        (replay_buffer.py:63-69): counted now, pushed into the ring as soon as it exists."""
        if not self._replay_dir.is_dir():
            return
        for fn in episode_files(self._replay_dir):
            try:
                _, _, eps_len = fn.stem.split("_")
                self._num_transitions += int(eps_len)
            except ValueError:
                continue
            self._num_episodes += 1
            self._pending.append(fn)

    def _ingest_pending(self, ring):
        names = [s.name for s in self._data_specs]
        pending, self._pending = self._pending, []
        for fn in pending:
            ep = load_episode(fn)
            ring.add_episode(frames_from_stacks(ep[names[0]], ring.stack) if ring.stack > 1 else ep[names[0]],
                             ep[names[1]].astype(np.float32), ep[names[2]].astype(np.float32),
                             ep[names[3]].astype(np.float32))

    def __len__(self):
        return self._num_transitions

    def add(self, time_step):
        for spec in self._data_specs:
            value = time_step[spec.name]
            if np.isscalar(value):
                value = np.full(spec.shape, value, spec.dtype)
            assert tuple(spec.shape) == tuple(value.shape) and spec.dtype == value.dtype
            self._current_episode[spec.name].append(value)
        if time_step.last():
            episode = {spec.name: np.array(self._current_episode[spec.name], spec.dtype)
                       for spec in self._data_specs}
            self._current_episode = defaultdict(list)
            self._store_episode(episode)

    def _ring(self, obs_c, action_dim):
        entry = _RINGS[self._key]
        if entry["ring"] is None:
            cap = entry["capacity"] or DEFAULT_CAPACITY
            stack = self._frame_stack
            assert obs_c % stack == 0, "observation channels must be a multiple of frame_stack"
            entry["ring"] = GpuRing(cap, obs_c // stack, stack, action_dim, self._device)
            self._ingest_pending(entry["ring"])
        return entry["ring"]

    def load_existing(self):
        """Create the ring from the episodes found in replay_dir (resume without waiting for a new episode).
        Returns the number of episodes loaded."""
        if not self._pending:
            return 0
        n = len(self._pending)
        ep = load_episode(self._pending[0])
        names = [s.name for s in self._data_specs]
        self._ring(ep[names[0]].shape[1], ep[names[1]].shape[1])
        return n

    def _store_episode(self, episode):
        names = [s.name for s in self._data_specs]
        obs_key, act_key, rew_key, disc_key = names[0], names[1], names[2], names[3]
        obs = episode[obs_key]
        ring = self._ring(obs.shape[1], episode[act_key].shape[1])
        # the stack must be the de-duplicated history the ring assumes (dmc.py:86-109)
        newest = frames_from_stacks(obs, ring.stack) if ring.stack > 1 else np.ascontiguousarray(obs)
        eps_idx = self._num_episodes
        eps_len = episode_len(episode)
        self._num_episodes += 1
        self._num_transitions += eps_len
        ring.add_episode(newest, episode[act_key].astype(np.float32),
                         episode[rew_key].astype(np.float32), episode[disc_key].astype(np.float32))
        if _RINGS[self._key].get("save_snapshot"):
            # keep the reference's files (replay_buffer.py:71-78; its loader deletes them unless save_snapshot)
            self._replay_dir.mkdir(parents=True, exist_ok=True)
            ts = datetime.datetime.now().strftime("%Y%m%dT%H%M%S")
            save_episode(episode, self._replay_dir / f"{ts}_{eps_idx}_{eps_len}.npz")


class RingSrc(C.Structure):
    """drq_ring_src (include/drqv2_b200.h): a sampled batch as a view of the ring"""
    _fields_ = [("frames", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p), ("discount", C.c_void_p),
                ("capacity", C.c_int64), ("frame_c", C.c_int32), ("stack", C.c_int32), ("A", C.c_int32),
                ("nstep", C.c_int32), ("gamma", C.c_float), ("reserved", C.c_int32),
                ("ep_table", C.c_void_p), ("n_episodes", C.c_void_p), ("seed", C.c_uint64), ("counter", C.c_void_p),
                ("ep_start", C.c_void_p), ("idx", C.c_void_p)]


class RingIterator:
    """Endless iterator of (obs, action, reward, discount, next_obs) device tensors."""

    def __init__(self, loader):
        self._l = loader
        self.batch_size = loader.batch_size

    @property
    def graph_key(self):
        """what a CUDA graph captured over next_into() bakes in: the ring's buffers and this loader's sampler
        state and constants (DrQV2Agent keys its graphs by it)"""
        l = self._l
        ring = l.ring()
        return (ring.frames.data_ptr(), ring.ep_table.data_ptr(), l._counter.data_ptr(), l.nstep, l.discount, l.seed)

    def ring_source(self):
        """The drq_ring_src of this loader over its ring: lets an update read the sampled frame stacks straight from
        the ring (drq_update_prologue_ring + drq_conv1_*_bf16_ring) instead of gathering them first."""
        l = self._l
        ring = l.ring()
        return RingSrc(ring.frames.data_ptr(), ring.action.data_ptr(), ring.reward.data_ptr(), ring.discount.data_ptr(),
                       ring.capacity, ring.frame_c, ring.stack, ring.A, l.nstep, float(l.discount), 0,
                       ring.ep_table.data_ptr(), ring.n_episodes.data_ptr(), l.seed, l._counter.data_ptr(),
                       l._ep_start.data_ptr(), l._idx.data_ptr())

    def check_ready(self):
        if getattr(self._l.ring(), "_n_eligible", 0) == 0:
            raise RuntimeError("replay ring has no episode long enough to sample from")

    def __iter__(self):
        return self

    def __next__(self):
        l = self._l
        l._ensure_out()
        self.next_into(l._obs, l._action, l._reward, l._discount, l._next_obs)
        return l._obs, l._action, l._reward, l._discount, l._next_obs

    def next_into(self, obs, action, reward, discount, next_obs, ep_start=None, idx=None):
        """Enqueue sample + gather into the given device buffers on the current stream
        (graph-capturable).  ep_start/idx (int32 device tensors [B]) bypass the sampler."""
        l = self._l
        ring = l.ring()
        B = l.batch_size
        if ep_start is None and getattr(ring, "_n_eligible", 0) == 0:
            raise RuntimeError("replay ring has no episode long enough to sample from")
        with torch.cuda.device(ring.device):           # kernels launch on the current device: the ring's
            s = _stream()
            if ep_start is None:
                call("drq_ring_sample_step", ring.ep_table.data_ptr(), ring.n_episodes.data_ptr(), l.nstep, l.seed,
                     l._counter.data_ptr(), l._ep_start.data_ptr(), l._idx.data_ptr(), B, s)     # draws, then counter += 1
                ep_start, idx = l._ep_start, l._idx
            call("drq_ring_gather_nstep", ring.frames.data_ptr(), ring.action.data_ptr(), ring.reward.data_ptr(),
                 ring.discount.data_ptr(), ring.capacity, ring.frame_c, ring.stack, ring.A, ep_start.data_ptr(),
                 idx.data_ptr(), B, l.nstep, float(l.discount), obs.data_ptr(), next_obs.data_ptr(),
                 action.data_ptr(), reward.data_ptr(), discount.data_ptr(), s)


class RingLoader:
    def __init__(self, replay_dir, max_size, batch_size, nstep, discount, seed=None):
        self.key = str(pathlib.Path(replay_dir))
        self.batch_size, self.nstep, self.discount = int(batch_size), int(nstep), float(discount)
        entry = _RINGS.setdefault(self.key, dict(ring=None, capacity=None, storage=None))
        entry["capacity"] = int(max_size)
        if seed is None:
            seed = int(np.random.get_state()[1][0])   # as replay_buffer.py:167-170 seeds its workers
        self.seed = seed
        self._obs = None
        ring = entry.get("ring")
        dev = ring.device if ring is not None else torch.device("cuda", torch.cuda.current_device())
        self._counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ep_start = torch.zeros(self.batch_size, dtype=torch.int32, device=dev)
        self._idx = torch.zeros(self.batch_size, dtype=torch.int32, device=dev)

    def ring(self):
        ring = _RINGS[self.key]["ring"]
        if ring is None:
            raise RuntimeError("replay ring is empty: no episode has been stored yet")
        if ring.min_len != self.nstep:
            ring.min_len = self.nstep
            ring._upload_table()
        return ring

    def _ensure_out(self):
        if self._obs is None:
            ring = self.ring()
            B, dev = self.batch_size, ring.device
            c = ring.frame_c * ring.stack
            self._obs = torch.zeros(B, c, IMG, IMG, dtype=torch.uint8, device=dev)
            self._next_obs = torch.zeros_like(self._obs)
            self._action = torch.zeros(B, ring.A, device=dev)
            self._reward = torch.zeros(B, 1, device=dev)
            self._discount = torch.zeros(B, 1, device=dev)

    def __iter__(self):
        return RingIterator(self)


def make_replay_loader(replay_dir, max_size, batch_size, num_workers, save_snapshot, nstep, discount):
    """Same positional signature as the reference (replay_buffer.py:173-190; train.py:68-71).
    ``num_workers`` has no meaning for a device-resident ring.  ``save_snapshot``: every stored episode is
    also written to ``replay_dir`` in the reference's npz format (the reference's loader deletes episode
    files once loaded unless this is set, replay_buffer.py:116-117)."""
    del num_workers
    loader = RingLoader(replay_dir, max_size, batch_size, nstep, discount)
    _RINGS[loader.key]["save_snapshot"] = bool(save_snapshot)
    return loader
