"""Pins the oracle (oracle/drq_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import json

import numpy as np
import pytest
import torch

from oracle import drq_oracle as O
from tests.helpers import aug_input, episode_arrays


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_replay_nstep_matches_reference_bit_exact(golden_dir):
    g = json.loads((golden_dir / "replay_golden.json").read_text())
    eps = [episode_arrays(e, T, g["A"]) for e, T in enumerate(g["lens"])]
    assert len(g["samples"]) > 30
    for s in g["samples"]:
        frames, action, reward, discount = eps[s["episode"]]
        obs, act, rew, disc, nxt = O.nstep_sample(frames, action, reward, discount, s["idx"], s["nstep"], 0.99)
        assert sha(obs) == s["obs_sha"] and sha(nxt) == s["next_sha"]
        assert np.array_equal(act, np.array(s["action"], np.float32))
        assert np.float32(rew[0]).tobytes().hex() == s["reward_hex"]
        assert np.float32(disc[0]).tobytes().hex() == s["discount_hex"]


def test_ring_gather_equals_episode_sample(golden_dir):
    """The ring (slots modulo capacity, wrap-around) gives the same bytes as the per-episode
    sample — incl. an episode that wraps around the end of the ring."""
    g = json.loads((golden_dir / "replay_golden.json").read_text())
    A = g["A"]
    rows = [T + 1 for T in g["lens"]]
    cap = sum(rows)
    first = 7   # place episode 0 so that the last episode wraps around
    ring_f = np.zeros((cap, 3, 84, 84), np.uint8)
    ring_a = np.zeros((cap, A), np.float32)
    ring_r = np.zeros((cap,), np.float32)
    ring_d = np.zeros((cap,), np.float32)
    starts, pos = [], first
    for e, T in enumerate(g["lens"]):
        f, a, r, d = episode_arrays(e, T, A)
        starts.append(pos % cap)
        for t in range(T + 1):
            s = (pos + t) % cap
            ring_f[s], ring_a[s], ring_r[s], ring_d[s] = f[t], a[t], r[t, 0], d[t, 0]
        pos += T + 1
    for nstep in (1, 3):
        ss = [s for s in g["samples"] if s["nstep"] == nstep]
        ep_start = np.array([starts[s["episode"]] for s in ss], np.int32)
        idx = np.array([s["idx"] for s in ss], np.int32)
        obs, act, rew, disc, nxt = O.ring_gather(ring_f, ring_a, ring_r, ring_d, ep_start, idx, nstep, 0.99)
        for b, s in enumerate(ss):
            assert sha(obs[b]) == s["obs_sha"] and sha(nxt[b]) == s["next_sha"]
            assert np.float32(rew[b, 0]).tobytes().hex() == s["reward_hex"]
            assert np.float32(disc[b, 0]).tobytes().hex() == s["discount_hex"]
            assert np.array_equal(act[b], np.array(s["action"], np.float32))


def test_aug_matches_reference(golden_dir):
    g = np.load(golden_dir / "aug_golden.npz")
    x = torch.from_numpy(aug_input(3, 2)).float()
    shifts = torch.from_numpy(g["shifts"])
    ref = torch.from_numpy(g["out"])
    grid = O.random_shift_grid_sample(x, shifts)
    assert torch.allclose(grid, ref, atol=2e-4, rtol=0)            # same float algorithm
    exact = O.random_shift_exact(x, shifts)
    # the reference's bilinear grid_sample is the integer shift to ~3.6e-3 on 0..255
    assert (exact - ref).abs().max().item() < 8e-3
    # and the integer shift is exactly replicate-pad + crop
    xp = torch.nn.functional.pad(x, (4,) * 4, "replicate")
    for i in range(3):
        sx, sy = int(shifts[i, 0]), int(shifts[i, 1])
        assert torch.equal(exact[i], xp[i, :, sy:sy + 84, sx:sx + 84])


def test_schedule_matches_reference(golden_dir):
    g = json.loads((golden_dir / "schedule_golden.json").read_text())
    for s, vals in g.items():
        for st, v in zip((0, 1, 50000, 100000, 250000), vals):
            assert O.schedule(s, st) == v


def _check_summary(t, summ, rtol_l2, atol_probe):
    t = t.detach().double().flatten()
    probe = t[torch.tensor(summ["probe_idx"])].numpy()
    want = np.array(summ["probe"])
    scale = max(summ["absmax"], 1e-30)
    assert abs(float(t.norm()) - summ["l2"]) <= rtol_l2 * max(summ["l2"], 1e-30) + 1e-12
    assert np.max(np.abs(probe - want)) <= atol_probe * scale + 1e-12


@pytest.mark.parametrize("case_idx", [0, 1])
@pytest.mark.parametrize("variant,aug", [("ref", "grid"), ("refx", "exact")])
def test_update_matches_reference(golden_dir, case_idx, variant, aug):
    torch.set_num_threads(1)
    g = json.loads((golden_dir / "update_golden.json").read_text())[case_idx]
    c = g["case"]
    params = O.synthetic_params(9, c["A"], c["F"], c["H"], seed=c["pseed"])
    agent = O.OracleAgent(params, c["lr"], 0.01, "linear(1.0,0.1,100000)", 0.3, aug=aug)
    for s, rec in enumerate(g["variants"][variant]["steps"]):
        b = O.synthetic_batch(c["B"], c["A"], seed=c["bseed"] + s)
        m = agent.update(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"], 2 * s,
                         b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
        for k, v in rec["metrics"].items():
            assert m[k] == pytest.approx(v, rel=2e-5, abs=1e-6), k
        for net, gs in rec["grads"].items():
            for k, summ in gs.items():
                _check_summary(agent.grads[net][k], summ, 1e-4, 1e-4)
        for net, ps in rec["params"].items():
            for k, summ in ps.items():
                # post-Adam: entries with |g| ~ 0 may flip by up to 2*lr
                t = agent.p[net][k].detach().double().flatten()
                probe = t[torch.tensor(summ["probe_idx"])].numpy()
                assert np.max(np.abs(probe - np.array(summ["probe"]))) <= 2.5 * c["lr"] * (s + 1), (net, k)
                assert abs(float(t.norm()) - summ["l2"]) <= 1e-4 * summ["l2"] + 1e-6
    # act()
    obs1 = O.synthetic_batch(1, c["A"], seed=99)["obs"][0]
    a_eval = agent.act(obs1, 5000, True)[0].numpy()
    assert np.allclose(a_eval, g["variants"][variant]["act_eval"], atol=2e-5)
    eps = torch.linspace(-1.5, 1.5, c["A"]).view(1, -1)
    a_train = agent.act(obs1, 5000, False, eps=eps)[0].numpy()
    assert np.allclose(a_train, g["variants"][variant]["act_train"], atol=2e-5)
