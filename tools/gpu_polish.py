import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
rows = lm.dispersed_params(B).rows(B).cuda()
for npol in (0, 1, 2, 4):
    res = {}
    for warm in (False, True):
        solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(warm_start=warm, n_polish=npol), device=0)
        raw = solver.solve_rows(rows); torch.cuda.synchronize()
        res[warm] = {k: (v.clone() if v is not None else None) for k, v in raw.items()}
        ms = solver.last_kernel_ms()
        print(f'n_polish {npol} warm {warm}: {ms:.1f} ms iters {raw["iterations"].double().mean():.2f} fails {(raw["status"]!=0).sum().item()} kkt max {raw["kkt"].max().item():.1e}')
    a, b = res[True]['traj'], res[False]['traj']
    scale = b.abs().amax(dim=1, keepdim=True)           # per problem, per row
    rel = ((a - b).abs() / scale).amax(dim=1)            # [10, B]
    print('   worst rel diff per row:', [f'{x:.1e}' for x in rel.amax(dim=1).tolist()])
    print('   problems with angledot diff > 1e-4:', int((rel[7] > 1e-4).sum()), ' > 1e-5:', int((rel[7] > 1e-5).sum()), ' control > 1e-3:', int((rel[9] > 1e-3).sum()))
