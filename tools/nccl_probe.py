"""All-reduce cost of the two gradient ranges of the data-parallel update (fp32, AVG), eager and inside a CUDA
graph.  Run under torchrun."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = {"encoder+critic (humanoid)": 6301774, "actor (humanoid)": 5094849}
out = []
for name, n in sizes.items():
    t = torch.randn(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 50 * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(10):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 50 * 1e3
    out.append(f"{name}: {n * 4 / 1e6:.1f} MB  eager {eager:.0f} us  in-graph {graph:.0f} us  (bus {2 * (dist.get_world_size() - 1) / dist.get_world_size() * n * 4 / eager / 1e3:.0f} GB/s)")
    del g
if rank == 0:
    print("\n".join(out), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
