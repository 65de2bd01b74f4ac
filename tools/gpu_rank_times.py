"""Per-rank kernel times of the config-4 batch (65 536 problems per GPU) when all GPUs of the box run at once: shows
whether a slow multi-GPU step is one slow device (clocks, power) or all of them.  Run under torchrun."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm

rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl")
dev = torch.device("cuda", rank)
B = 65536
s = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(), device=dev)
rows = lm.dispersed_params(B, seed=11).rows(B, device=dev)
ms = []
for i in range(6):
    if world > 1:
        dist.barrier()
    r = s.solve_rows(rows, trajectories=False); ms.append(s.last_kernel_ms())
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(rank)
clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
print(f"rank {rank}: kernel ms {['%.1f' % m for m in ms]} sm clock now {clk} MHz power {pw:.0f} W", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
