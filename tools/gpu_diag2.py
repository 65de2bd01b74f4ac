import sys, time, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
rows = lm.dispersed_params(B).rows(B).cuda()
out = solver.alloc_outputs(B, True, True)
for _ in range(3): solver.solve_rows(rows, True, out=out)
torch.cuda.synchronize()
T = time.perf_counter
for step in range(4):
    t0 = T(); raw = solver.solve_rows(rows, True, out=out); t1 = T()
    ms = solver.last_kernel_ms(); t2 = T()
    c = int((raw["status"] == 0).sum()); t3 = T()
    i = int(raw["iterations"].sum()); t4 = T()
    print(f'step {step}: launch {1e3*(t1-t0):.2f} ms, wait {1e3*(t2-t1):.2f} (kernel {ms:.1f}), conv {1e3*(t3-t2):.2f}, iters {1e3*(t4-t3):.2f}')
# same with events around everything
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); ev0.record()
for step in range(5):
    raw = solver.solve_rows(rows, True, out=out); ms = solver.last_kernel_ms()
    c = int((raw["status"] == 0).sum()); i = int(raw["iterations"].sum())
ev1.record(); torch.cuda.synchronize()
print('events total per step', ev0.elapsed_time(ev1)/5)
torch.cuda.synchronize(); ev0.record()
for step in range(5):
    raw = solver.solve_rows(rows, True, out=out)
ev1.record(); torch.cuda.synchronize()
print('events total per step, no host reads', ev0.elapsed_time(ev1)/5)
