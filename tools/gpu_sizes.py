import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
ref = None
for B in (1023, 1024, 1025, 37887, 37888, 37889, 65535, 262144):
    rows = lm.dispersed_params(B, seed=11).rows(B).cuda()
    raw = solver.solve_rows(rows, trajectories=(B <= 65536)); torch.cuda.synchronize()
    st = raw['status']
    tf = raw['tf']
    if ref is None: ref = tf[:1000].clone()
    d = (tf[:1000] - ref).abs().max().item()
    print(f'B {B}: ms {solver.last_kernel_ms():.1f} converged {(st==0).sum().item()}/{B} max|tf-tf_ref| first 1000 {d:.2e}')
