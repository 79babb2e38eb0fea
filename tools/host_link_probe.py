"""Host-link probe for the e2e numbers at N > 1 (run under torchrun, one rank per GPU): pinned 134 MB H2D + D2H copies on
two streams (what one fvm_eqs!(du,u,p,t) call with host vectors moves at 4096^2), first one rank at a time, then all ranks
at once.  Prints per-rank and aggregate GB/s: the ceiling the banded pipeline of fvm_pipe.cu can reach on this host."""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 16777216
hin = torch.empty(n, dtype=torch.float64).pin_memory()
hout = torch.empty(n, dtype=torch.float64).pin_memory()
hin.fill_(1.0)
hout.fill_(0.0)
din = torch.empty(n, dtype=torch.float64, device="cuda")
dout = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(iters=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def bar():
    if world > 1:
        dist.barrier()


run(3)
alone = None
for r in range(world):
    bar()
    if r == rank:
        alone = run()
    bar()
bar()
together = run()
bar()
line = "rank %d  alone %.2f ms (%.1f GB/s both ways)  all ranks at once %.2f ms (%.1f GB/s)" % (
    rank, alone * 1e3, 2 * 8 * n / alone / 1e9, together * 1e3, 2 * 8 * n / together / 1e9)
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    line += "  gpu pci bus %02x" % bus
except Exception:
    pass
line += "  cpus %s" % (sorted(os.sched_getaffinity(0))[:3] + ["..."] + sorted(os.sched_getaffinity(0))[-1:])
print(line, flush=True)
if world > 1:
    v = torch.tensor([2 * 8 * n / together / 1e9], device="cuda")
    dist.all_reduce(v)
    if rank == 0:
        print("aggregate with all %d ranks copying: %.1f GB/s" % (world, float(v.item())), flush=True)
    dist.destroy_process_group()
