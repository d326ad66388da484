#!/usr/bin/env python
"""Benchmark of the DrQ-v2 agent update (BASELINE.json metric: updates/sec at batch 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one DrQV2Agent.update (critic + actor + soft target update) on a fresh
batch of the walker_walk shape (B=256, 9x84x84 uint8 stacks, A=6, F=50, H=1024, n-step 3).

ours      : `value` = updates/s with the replay ring resident in HBM (sample + n-step
            gather + update captured in one CUDA graph, timed with CUDA events);
            `e2e` = the same update through the public API fed from HOST batches
            (pinned memory -> H2D inside the timed region, metrics read back D2H).
reference : the CPU restatement of the reference's update (oracle/ port; the reference is
            pure PyTorch, there is nothing to compile) on the host cores, bounded sample.
Multi-GPU : one process per GPU (torchrun); independent agents (ensemble members) with no
            collective, value = sum over ranks / max-over-ranks time, scaling "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCHED = "linear(1.0,0.1,100000)"
E_F, E_B = 84_561_984, 160_409_664     # encoder fwd / bwd FLOP per sample
CONV_MACS = {39: 14_017_536, 37: 12_616_704, 35: 11_289_600}


def update_flops(B, A, F, H):
    T = 2 * 39200 * F
    Q = 2 * ((F + A) * H + H * H + H)
    P = 2 * (F * H + H * H + H * A)
    return B * (2 * E_F + E_B + 8 * T + (2 * P + 6 * Q) + (6 * Q + 2 * P))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fill_ring(loader_dir, A, episodes, rows, device, seed=1):
    """SURVEY §8d synthetic replay: E episodes x rows steps of uniform u8 frames, action~U(-1,1),
    reward~U(0,1), discount 1; row 0 is the reset dummy.  Written straight into the ring."""
    from drqv2_b200 import replay_buffer as rb
    cap = episodes * rows
    ring = rb.GpuRing(cap, 3, 3, A, device)
    g = torch.Generator(device=device).manual_seed(seed)
    ring.frames.copy_(torch.randint(0, 256, ring.frames.shape, dtype=torch.uint8, device=device, generator=g))
    ring.action.copy_(torch.rand(cap, A, device=device, generator=g) * 2 - 1)
    ring.reward.copy_(torch.rand(cap, device=device, generator=g))
    ring.discount.fill_(1.0)
    starts = torch.arange(episodes, device=device) * rows
    ring.action[starts] = 0
    ring.reward[starts] = 0
    ring.episodes = [(int(e * rows), rows) for e in range(episodes)]
    ring.head = 0
    ring._upload_table()
    rb._RINGS[str(loader_dir)] = dict(ring=ring, capacity=cap, storage=None)
    return ring


def time_kernel(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def run_ours(args, rank, world):
    from drqv2_b200 import DrQV2Agent, _lib, make_replay_loader
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dp = args.parallel == "dp" and world > 1
    A, Fd, H = args.action_dim, args.feature_dim, args.hidden_dim
    B = args.batch // world if dp else args.batch  # dp: --batch is the global batch, sharded over the ranks
    if dp and args.batch % world:
        raise SystemExit(f"--batch {args.batch} does not split over {world} ranks")
    torch.manual_seed(0 if dp else rank)          # ensemble member = independent seed; dp = one agent
    np.random.seed(7 + rank)
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, SCHED, 0.3, False,
                       use_cuda_graph=True, seed=0 if dp else rank, mode=args.mode, data_parallel=dp)
    key = f"/bench/ring{rank}"
    fill_ring(key, A, args.episodes, 501, dev, seed=1 + rank)
    loader = make_replay_loader(key, args.episodes * 501, B, 0, False, 3, 0.99)
    it = iter(loader)
    K = 1 if dp else max(1, args.agents_per_gpu)
    # further ensemble members of this GPU: own parameters, optimiser state, ring, RNG stream, graph and CUDA stream
    members, member_its = [agent], [it]
    for k in range(1, K):
        torch.manual_seed(1000 * k + rank)
        members.append(DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, SCHED, 0.3, False,
                                  use_cuda_graph=True, seed=1000 * k + rank, mode=args.mode))
        fill_ring(f"{key}_m{k}", A, args.episodes, 501, dev, seed=1 + rank + 1000 * k)
        member_its.append(iter(make_replay_loader(f"{key}_m{k}", args.episodes * 501, B, 0, False, 3, 0.99)))
    member_streams = [torch.cuda.Stream(device=dev) for _ in range(K)]

    def update_all(iters, step):
        """one update of every member of this GPU; returns member 0's metrics"""
        if K == 1:
            return agent.update(iters[0], step)
        cur = torch.cuda.current_stream()
        pending = []
        for ag, mit, st in zip(members, iters, member_streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                pending.append(ag.update_async(mit, step))
        m = dict()
        for ag, ws_, st in zip(members, pending, member_streams):
            if ag.use_tb:
                with torch.cuda.stream(st):
                    mk = ag.read_metrics(ws_)
                m = m or mk
            cur.wait_stream(st)
        return m

    # count kernel launches of one update (eager pass == what the graph replays): every C-ABI call of the
    # update goes through _lib.call; entry points that launch two kernels are listed below
    n_calls = [0]
    orig_call = _lib.call
    two_kernel_calls = ("drq_conv3x3_wgrad_bf16", "drq_conv1_wgrad_bf16", "drq_conv3x3_wgrad_f32", "drq_conv1_wgrad_f32")

    def counting_call(name, *a):
        # drq_ln_tanh_bwd launches its parameter-gradient kernel only when dgamma (argument 8) is given
        # the bf16 wgrad entry points skip their reduce kernel when dw (argument 4) is NULL (reduced by one launch later)
        two = (name in two_kernel_calls and (not name.endswith("_bf16") or a[4])) or (name == "drq_ln_tanh_bwd" and a[8])
        n_calls[0] += 2 if two else 1
        return orig_call(name, *a)

    import drqv2_b200._bf16 as BF
    import drqv2_b200.drqv2 as D
    import drqv2_b200.replay_buffer as R
    step = 0
    agent.update(it, step); step += 2             # eager warm-up
    D.call = R.call = BF.call = counting_call
    agent.use_cuda_graph = False
    agent.update(it, step); step += 2             # eager, counted
    agent.use_cuda_graph = True
    launches_per_update = n_calls[0]
    D.call = R.call = BF.call = orig_call
    for _ in range(max(args.warmup, 3)):
        update_all(member_its, step); step += 2
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:                     # one nvidia-smi poller per job: NVML queries perturb the GPUs they touch, and in
        sampler.start()               # data-parallel mode a stall on any rank stalls every rank twice per update
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        update_all(member_its, step); step += 2
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = (1 if dp else world * K) * 1e3 / ms_per_step  # dp: global-batch updates/s; ensemble: sum over agents

    # ---- e2e: public API fed from host batches (pinned), metrics read back.  prefetch: the agent pulls the next
    # host batch one update ahead and overlaps its H2D copy with the running update (a DrQV2Agent option)
    for ag in members:
        ag.use_tb = True
        ag.prefetch = True
    g = torch.Generator().manual_seed(100 + rank)
    nhost = 4
    host_batches = []
    for _ in range(nhost):
        host_batches.append(tuple(t.pin_memory() for t in (
            torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, generator=g),
            torch.rand(B, A, generator=g) * 2 - 1, torch.rand(B, 1, generator=g),
            torch.full((B, 1), 0.970299065), torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, generator=g))))

    def host_iter():
        i = 0
        while True:
            yield host_batches[i % nhost]
            i += 1

    hits = [host_iter() for _ in range(K)]
    for _ in range(4):
        update_all(hits, step); step += 2
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0.record()
    for _ in range(args.steps):
        m = update_all(hits, step); step += 2
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_value = (1 if dp else world * K) * 1e3 / (e2e_ms / args.steps)
    h2d = K * (sum(t.numel() * t.element_size() for t in host_batches[0]) + 64)
    d2h = K * 8 * 4
    assert np.isfinite(m["critic_loss"])
    for ag in members:
        ag.use_tb = False
        ag.prefetch = False

    out = None
    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        flops = update_flops(B, A, Fd, H) * (world if dp else 1)     # per counted update
        # ---- dominant kernel, timed alone with CUDA events on its launch stream
        ws = agent.workspace(B)
        s = torch.cuda.current_stream().cuda_stream
        pe = lambda k: agent._p("encoder", k)
        ge = lambda k: agent._g("encoder", k)
        if args.mode == "bf16":
            bw, st = agent.bf16_workspace(B), agent._bf16
            acts = [a.data_ptr() for a in bw.acts]
            d = [t.data_ptr() for t in bw.dpre]
            cand = {
                "conv3x3_tc_kernel<fwd>(layer2,N=2B)": (lambda: _lib.call("drq_conv3x3_fwd_bf16", acts[0], st.conv_wf[0].data_ptr(), pe("convnet.2.bias"), acts[1], 2 * B, 39, 0, 0, 0, 0, s),
                                                        2 * CONV_MACS[39] * 2 * B),
                "conv3x3_tc_kernel<dgrad>(layer2,N=B)": (lambda: _lib.call("drq_conv3x3_dgrad_bf16", d[1], st.conv_wd[0].data_ptr(), acts[0], 2 * B, d[0], B, 39, s),
                                                         2 * CONV_MACS[39] * B),
                "conv3x3_wgrad_tc_kernel(layer2,N=B)": (lambda: _lib.call("drq_conv3x3_wgrad_bf16", acts[0], 2 * B, d[1], bw.wg_ws[1].data_ptr(), ge("convnet.2.weight"), ge("convnet.2.bias"), B, 39, s),
                                                        2 * CONV_MACS[39] * B),
            }
        else:
            acts = [a.data_ptr() for a in ws.acts]
            d = [t.data_ptr() for t in ws.dpre]
        cand = cand if args.mode == "bf16" else {
            "conv3x3_fwd_f32(layer2,N=2B)": (lambda: _lib.call("drq_conv3x3_fwd_f32", acts[0], pe("convnet.2.weight"), pe("convnet.2.bias"), acts[1], 2 * B, 39, 0, s),
                                              2 * CONV_MACS[39] * 2 * B),
            "conv3x3_dgrad_f32(layer2,N=B)": (lambda: _lib.call("drq_conv3x3_dgrad_f32", d[1], pe("convnet.2.weight"), acts[0], d[0], B, 39, s),
                                              2 * CONV_MACS[39] * B),
            "conv3x3_wgrad_f32(layer2,N=B)": (lambda: _lib.call("drq_conv3x3_wgrad_f32", acts[0], d[1], ws.wgrad_ws.data_ptr(), ge("convnet.2.weight"), ge("convnet.2.bias"), B, 39, s),
                                              2 * CONV_MACS[39] * B),
        }
        kt = {k: (time_kernel(fn), fl) for k, (fn, fl) in cand.items()}
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of the same
        # launches (profiles/r1_ncu_full_conv_kernels.txt, B = 256)
        traffic = {"conv3x3_tc_kernel<fwd>(layer2,N=2B)": 69.2e6, "conv3x3_tc_kernel<dgrad>(layer2,N=B)": 59.9e6,
                   "conv3x3_wgrad_tc_kernel(layer2,N=B)": 56.3e6} if (args.mode == "bf16" and B == 256) else {}
        # share of the step: fwd x3 layers x(2B), dgrad x3, wgrad x3 are of the same class
        dom = max(kt, key=lambda k: kt[k][0])
        dt, fl = kt[dom]
        roof = {"bound": "tensor", "kernel": dom, "achieved": fl / dt / 1e12, "peak": tf_burst,
                "unit": "TFLOP/s", "frac": fl / dt / 1e12 / tf_burst, "traffic": traffic.get(dom), "peak_source": src,
                "kernel_ms": dt * 1e3,
                "all_kernels_ms": {k: v[0] * 1e3 for k, v in kt.items()},
                "whole_update": {"flop": flops, "achieved_per_gpu": flops * value / world / 1e12,
                                 "frac_of_sustained": flops * value / world / 1e12 / tf_sust}}
        # ---- HBM-bound kernels of the update, timed alone the same way (algorithmic bytes per launch)
        ar = agent._arena
        F32 = 4
        off, n = ar.seg["actor"][0], ar.seg["actor"][2]
        coff, cn = ar.seg["critic"][0], ar.seg["critic"][2]
        eoff, en = ar.seg["encoder"][0], ar.seg["encoder"][2] + ar.seg["critic"][2]
        sc = agent._scal_dev.data_ptr()

        def adam_actor_ema():
            _lib.call("drq_adam_ema_step", ar.params.data_ptr() + F32 * off, ar.grads.data_ptr() + F32 * off,
                      ar.exp_avg.data_ptr() + F32 * off, ar.exp_avg_sq.data_ptr() + F32 * off, n, sc,
                      ar.params.data_ptr() + F32 * coff, ar.target.data_ptr(), cn, 0.01, 0.99, s)

        def adam_enc_critic():
            _lib.call("drq_adam_step", ar.params.data_ptr() + F32 * eoff, ar.grads.data_ptr() + F32 * eoff,
                      ar.exp_avg.data_ptr() + F32 * eoff, ar.exp_avg_sq.data_ptr() + F32 * eoff, en, sc, s)

        saved = [t.clone() for t in (ar.params, ar.exp_avg, ar.exp_avg_sq, ar.target)]
        opt_name, pack_bytes = "adam_ema_kernel(both optimiser phases)", 0
        if args.mode == "bf16":                    # the bf16 update steps and refreshes the bf16 operands in one launch
            st = agent._bf16
            adam_actor_ema, adam_enc_critic = st.step_actor_target, st.step_critic_encoder
            opt_name = "adam_pack_kernel(both optimiser phases, incl. bf16 operand refresh)"
            pack_bytes = 2 * (n + en + cn)
        t_adam = time_kernel(adam_actor_ema) + time_kernel(adam_enc_critic)
        for t, sv in zip((ar.params, ar.exp_avg, ar.exp_avg_sq, ar.target), saved):
            t.copy_(sv)                            # the timing launches stepped the optimiser: restore
        adam_bytes = 28 * (n + en) + 12 * cn + pack_bytes   # p,g,m,v read + p,m,v written; EMA: p, tp read + tp written; bf16 copies
        t_gather = time_kernel(lambda: it.next_into(ws.obs[:B], ws.action, ws.reward, ws.discount, ws.obs[B:]))
        gather_bytes = 2 * 2 * B * 9 * 84 * 84     # u8 stacks read from the ring + written to the batch
        # (timed back to back, so the 126 MB L2 holds part of the working set: fractions above 1 are L2 hits)
        roof["hbm_kernels"] = {
            opt_name: {"bytes": adam_bytes, "ms": t_adam * 1e3, "achieved_gbs": adam_bytes / t_adam / 1e9,
                                                       "peak_gbs": hbm, "frac": adam_bytes / t_adam / 1e9 / hbm},
            "ring_sample+ring_gather_kernel": {"bytes": gather_bytes, "ms": t_gather * 1e3, "achieved_gbs": gather_bytes / t_gather / 1e9,
                                               "peak_gbs": hbm, "frac": gather_bytes / t_gather / 1e9 / hbm}}
        cpu = cpu_baseline(args, steps=2)
        out = {"metric": "DrQ-v2 updates/sec at batch 256", "value": value, "unit": "updates/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "strong" if dp else "weak", "vs_baseline": None,
               "dtype": "bf16" if args.mode == "bf16" else "f32",
               "data": "synthetic",
               "config": {"workload": f"configs[1]: walker_walk-shape agent.update, B={B}, 9x84x84 u8 stacks, A={A}, "
                                      f"F={Fd}, H={H}, n-step 3, GPU-resident replay ring ({args.episodes} episodes x 501 "
                                      f"rows), CUDA-graphed, {'bf16 tensor-core mode (fp32 master weights, fp32 accumulation)' if args.mode == 'bf16' else 'fp32 parity mode'}" + (f"; global batch {args.batch} data-parallel over {world} GPUs (NCCL gradient all-reduce in the graph)" if dp else (f"; {world * K} independent agents (ensemble), {K} per GPU on their own streams, no collective" if world * K > 1 else "")),
                          "l2": "inputs larger than L2: each step gathers a fresh 32.5 MB batch from a "
                                f"{args.episodes * 501 * 21168 / 1e6:.0f} MB ring and streams ~700 MB of activations",
                          "mode": args.mode, "agents_per_gpu": K},
               "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": launches_per_update * args.steps * K, "launches_per_update": launches_per_update,
               "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    return out


def _load_reference():
    """The unmodified reference modules from baseline/_ref (baseline/install_ref.py), or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "drqv2.py")):
        return None
    import types
    for name in ("hydra", "omegaconf"):                     # imported by the reference, unused on this path
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object
    import importlib.util
    mods = {}
    for name in ("utils", "drqv2"):
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(ref, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        if name == "drqv2":
            saved = sys.modules.get("utils")
            sys.modules["utils"] = mods["utils"]            # the reference's `import utils`
            try:
                spec.loader.exec_module(m)
            finally:
                if saved is None:
                    sys.modules.pop("utils", None)
                else:
                    sys.modules["utils"] = saved
        else:
            spec.loader.exec_module(m)
        mods[name] = m
    return mods["drqv2"]


def cpu_baseline(args, steps=2, warm=1):
    """The reference's own agent.update (unmodified sources in baseline/_ref, device 'cpu', all host
    threads) on a bounded sample: `steps` updates of the same B=256 workload from pre-collated tensors.
    Falls back to the oracle port of the same path when the reference sources are not present."""
    torch.set_num_threads(os.cpu_count())
    B, A, Fd, H = args.batch, args.action_dim, args.feature_dim, args.hidden_dim
    ref = _load_reference()
    if ref is not None:
        torch.manual_seed(0)
        agent = ref.DrQV2Agent((9, 84, 84), (A,), "cpu", 1e-4, Fd, H, 0.01, 2000, 2, SCHED, 0.3, False)
        g = torch.Generator().manual_seed(1)
        batch = (torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, generator=g),
                 torch.rand(B, A, generator=g) * 2 - 1, torch.rand(B, 1, generator=g),
                 torch.full((B, 1), 0.970299065), torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, generator=g))

        def it():
            while True:
                yield batch
        ri = it()
        for i in range(warm):
            agent.update(ri, 2 * i)
        t = time.perf_counter()
        for i in range(steps):
            agent.update(ri, 2 * (i + warm))
        dt = (time.perf_counter() - t) / steps
        return {"value": 1.0 / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{steps} updates (after {warm} warm-up) of the reference's DrQV2Agent.update on CPU, B={B}, "
                          "pre-collated synthetic tensors"}
    from oracle import drq_oracle as O
    params = O.synthetic_params(9, A, Fd, H, seed=0)
    agent = O.OracleAgent(params, 1e-4, 0.01, SCHED, 0.3, dtype=torch.float32, aug="grid")
    b = O.synthetic_batch(B, A, seed=1)
    a = (b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"])
    k = (b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
    for i in range(warm):
        agent.update(*a, 2 * i, *k)
    t = time.perf_counter()
    for i in range(steps):
        agent.update(*a, 2 * (i + warm), *k)
    dt = (time.perf_counter() - t) / steps
    return {"value": 1.0 / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} updates (after {warm} warm-up) of the same B={B} workload, pre-collated tensors, "
                      "float grid_sample aug as the reference ships it"}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    cpu = cpu_baseline(args, steps=steps, warm=warm)
    B, A, Fd, H = args.batch, args.action_dim, args.feature_dim, args.hidden_dim
    return {"impl": "reference", "metric": "DrQ-v2 updates/sec at batch 256", "value": cpu["value"],
            "unit": "updates/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[0]: reference's CPU update path ({cpu['kind']}, torch CPU), B={B}, A={A}, "
                                   f"F={Fd}, H={H}"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--action-dim", type=int, default=6)
    ap.add_argument("--feature-dim", type=int, default=50)
    ap.add_argument("--hidden-dim", type=int, default=1024)
    ap.add_argument("--episodes", type=int, default=64)
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--agents-per-gpu", type=int, default=1,
                    help="ensemble members per GPU (BASELINE configs[3] uses 8): independent agents replayed on their own "
                         "streams; value = updates/s summed over all members of all GPUs")
    ap.add_argument("--parallel", default="ensemble", choices=["ensemble", "dp"],
                    help="N > 1: independent agents per GPU (weak scaling, no collective) or one agent with the "
                         "global --batch sharded over the GPUs and NCCL gradient all-reduce (strong scaling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")
    out = run_ours(args, rank, world)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        # graphs holding captured NCCL kernels were dropped with the agent (run_ours returned); a teardown
        # that still blocks must not hold the job: force-exit after a grace period
        import gc
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
