"""The ring-direct head of a tensor-core update (north_star (1)/(2): frame stack by index, augmented frames never
materialised): drq_update_prologue_ring + drq_conv1_*_bf16_ring against the three-launch path they replace
(drq_update_prologue, drq_ring_sample_step, drq_ring_gather_nstep) followed by the plain conv1 kernels - bit for
bit, on a ring whose episodes wrap around the end of the buffer, and through a whole update."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import drq_oracle as O

pytestmark = pytest.mark.gpu
SCHED = "linear(1.0,0.1,100000)"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ring_with_wrapping_episodes(key, A, dev, seed=3):
    """capacity 100; episodes of 31 rows written one after the other so that the fourth wraps past the end"""
    from drqv2_b200 import replay_buffer as R
    g = np.random.default_rng(seed)
    ring = R.GpuRing(100, 3, 3, A, dev)
    for e in range(4):
        rows = 31
        ring.add_episode(g.integers(0, 256, (rows, 3, 84, 84), dtype=np.uint8), g.random((rows, A), dtype=np.float32) * 2 - 1,
                         g.random(rows, dtype=np.float32), np.where(g.random(rows) < 0.8, 1.0, 0.9).astype(np.float32))
    R._RINGS[key] = dict(ring=ring, capacity=100, storage=None)
    assert any(s + r > ring.capacity for s, r in ring.episodes), "an episode must wrap"
    return ring


@pytest.mark.parametrize("nstep", [1, 3])
def test_prologue_ring_and_conv1_from_ring_bitwise(dev, nstep):
    from drqv2_b200 import _lib, make_replay_loader
    A, B, pad = 6, 24, 4
    key = f"/test/ringdirect{nstep}"
    _ring_with_wrapping_episodes(key, A, dev)
    np.random.seed(5)
    loader = make_replay_loader(key, 100, B, 0, False, nstep, 0.99)
    it = iter(loader)
    ring = loader.ring()
    seed = 1234
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    scal_ring = torch.arange(4 * 32, dtype=torch.float32).view(4, 32).pin_memory()

    def head(direct):
        loader._counter.fill_(7)
        cursor, counter = z(1, dt=torch.int64) + 2, z(1, dt=torch.int64) + 11
        out = dict(scal=z(32), shift_o=z(B, 2, dt=torch.int32), shift_n=z(B, 2, dt=torch.int32), eps_c=z(B, A), eps_a=z(B, A),
                   action=z(B, A), reward=z(B, 1), discount=z(B, 1), obs=z(B, 9, 84, 84, dt=torch.uint8),
                   nxt=z(B, 9, 84, 84, dt=torch.uint8))
        args = (scal_ring.data_ptr(), 4, cursor.data_ptr(), out["scal"].data_ptr(), seed, counter.data_ptr(), pad,
                out["shift_o"].data_ptr(), out["shift_n"].data_ptr(), out["eps_c"].data_ptr(), out["eps_a"].data_ptr(), B, A)
        if direct:
            src = it.ring_source()
            tail = (C.byref(src), out["action"].data_ptr(), out["reward"].data_ptr(), out["discount"].data_ptr())
            if direct == "parts":       # the two launches of the update's schedule: shifts + sample, then the rest
                _lib.call("drq_update_prologue_ring_part", *args, *tail, 1, _stream())
                assert int(cursor) == 2 and int(counter) == 11, "part 1 leaves the scalar cursor and the draw counter alone"
                _lib.call("drq_update_prologue_ring_part", *args, *tail, 2, _stream())
            else:
                _lib.call("drq_update_prologue_ring", *args, *tail, _stream())
            out["src"] = src
        else:
            _lib.call("drq_update_prologue", *args, _stream())
            it.next_into(out["obs"], out["action"], out["reward"], out["discount"], out["nxt"])
        torch.cuda.synchronize()
        out.update(ep_start=loader._ep_start.clone(), idx=loader._idx.clone(), cursor=cursor.clone(), counter=counter.clone(),
                   lcounter=loader._counter.clone())
        return out

    a, b, c = head(False), head(True), head("parts")
    for k in ("scal", "shift_o", "shift_n", "eps_c", "eps_a", "action", "reward", "discount", "ep_start", "idx", "cursor",
              "counter", "lcounter"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a[k], c[k]), ("parts", k)
    assert int(a["lcounter"]) == 8 and int(a["counter"]) == 12 and int(a["cursor"]) == 3
    # the sampled windows include wrapped ones
    es, ix = a["ep_start"].cpu().numpy().astype(np.int64), a["idx"].cpu().numpy()
    assert ((es + ix + nstep - 1) >= ring.capacity).any() and ((es + ix) < ring.capacity).any()
    # conv1 forward / weight gradient: rows from the ring == rows from the gathered stacks
    g = torch.Generator().manual_seed(1)
    w = ((torch.rand(32, 9, 3, 3, generator=g) - 0.5) * 0.3).to(dev)
    bias = ((torch.rand(32, generator=g) - 0.5) * 0.1).to(dev)
    wp = torch.zeros(_lib.lib().drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
    _lib.call("drq_pack_conv1_w_bf16", w.data_ptr(), bias.data_ptr(), wp.data_ptr(), 9, _stream())
    stacks = torch.cat([a["obs"], a["nxt"]])
    shift = torch.cat([a["shift_o"], a["shift_n"]]).contiguous()
    nel = _lib.lib().drq_wb_elems(2 * B)
    y0, y1 = z(nel, dt=torch.bfloat16), z(nel, dt=torch.bfloat16)
    _lib.call("drq_conv1_fwd_bf16", stacks.data_ptr(), shift.data_ptr(), wp.data_ptr(), y0.data_ptr(), 2 * B, 9, pad, _stream())
    _lib.call("drq_conv1_fwd_bf16_ring", C.byref(b["src"]), B, shift.data_ptr(), wp.data_ptr(), y1.data_ptr(), 2 * B, pad, _stream())
    torch.cuda.synchronize()
    assert torch.equal(y0.view(torch.int16), y1.view(torch.int16)) and y0.float().abs().sum() > 0
    # stack-by-index against the oracle's gather (replay_buffer.py:150-153 on dmc.py:98-109 stacks)
    want = O.ring_gather(ring.frames.cpu().numpy(), ring.action.cpu().numpy(), ring.reward.cpu().numpy(),
                         ring.discount.cpu().numpy(), es, ix, nstep, 0.99)
    for got, w_ in zip((a["obs"], b["action"], b["reward"], b["discount"], a["nxt"]), want):
        assert np.array_equal(got.cpu().numpy(), w_)
    d = ((torch.rand(4, B * 1776 + 128, 8, generator=g) - 0.5) * 1e-2).to(torch.bfloat16).to(dev)
    ws0, ws1 = z(_lib.lib().drq_conv1_wgrad_bf16_ws_floats()), z(_lib.lib().drq_conv1_wgrad_bf16_ws_floats())
    dw0, db0, dw1, db1 = z(32, 9, 3, 3), z(32), z(32, 9, 3, 3), z(32)
    _lib.call("drq_conv1_wgrad_bf16", stacks.data_ptr(), shift.data_ptr(), d.data_ptr(), ws0.data_ptr(), dw0.data_ptr(),
              db0.data_ptr(), B, 9, pad, _stream())
    _lib.call("drq_conv1_wgrad_bf16_ring", C.byref(b["src"]), B, shift.data_ptr(), d.data_ptr(), ws1.data_ptr(), dw1.data_ptr(),
              db1.data_ptr(), B, pad, _stream())
    torch.cuda.synchronize()
    assert torch.equal(dw0, dw1) and torch.equal(db0, db1) and dw0.abs().sum() > 0


def test_update_from_ring_direct_equals_gathered(dev):
    """Whole tensor-core updates fed from the ring: reading the stacks by index gives the parameters of the
    gather-first path bit for bit (eager and captured)."""
    from drqv2_b200 import DrQV2Agent, make_replay_loader
    A, Fd, H, B = 6, 50, 128, 16
    params = O.synthetic_params(9, A, Fd, H, seed=6)
    agents = []
    for direct in (False, True):
        key = f"/test/ringupd{int(direct)}"
        _ring_with_wrapping_episodes(key, A, dev)
        np.random.seed(9)
        it = iter(make_replay_loader(key, 100, B, 0, False, 3, 0.99))
        agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True, use_cuda_graph=True,
                           seed=5, mode="bf16")
        agent.ring_direct = direct
        for net in ("encoder", "actor", "critic", "critic_target"):
            getattr(agent, net).load_state_dict(params[net])
        ms = [agent.update(it, 2 * s) for s in range(5)]
        assert all(np.isfinite(m["critic_loss"]) for m in ms)
        assert any(isinstance(v, torch.cuda.CUDAGraph) for v in agent._graphs.values())
        agents.append((agent, ms))
    torch.cuda.synchronize()
    (a0, m0), (a1, m1) = agents
    assert m0 == m1
    for net in ("encoder", "actor", "critic", "critic_target"):
        for (n1, p1), (_, p2) in zip(getattr(a0, net).named_parameters(), getattr(a1, net).named_parameters()):
            assert torch.equal(p1, p2), (net, n1)
