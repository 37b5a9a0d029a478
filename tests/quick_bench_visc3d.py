import os, sys, time, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'python-fluid-simulation_b200'))
import torch, scenes
from solver import _native as N
from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
lib=N.load()
for Ng, dtype in ((128,torch.float32),(256,torch.float32),(256,torch.float64)):
    sc=scenes.buckling(Ng, device='cuda', mu=100.0)
    s=ViscosityCGSolver3D(sc['gres'], sc['bound_size'], dtype=dtype)
    s.max_iter=0
    v=[sc[k].clone() for k in ('vx','vy','vz')]
    try: s.solve(sc['dt'],100.0,sc['rho'],*v,sc['sphi'],None,None,sc['lvol'],tol=0.0)
    except ValueError: pass
    scale=sc['dt']/s.cell_vol/sc['rho']
    N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, 100.0, 10, 0),'w'); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    K=100
    e0.record(); N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, 100.0, K, torch.cuda.current_stream().cuda_stream),'b'); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/K
    n=Ng; F=3*n*n*(n+1); V7=F+n**3+3*(n+1)*(n+1)*n
    esz=4 if dtype==torch.float32 else 8
    gb=(11*F+V7)*esz/1e9
    st=N.CgStats(); lib.fs_visc3d_read_stats(s._e.h, ctypes.byref(st), 0)
    print(f"N={Ng} {dtype}: {ms:.4f} ms/iter  {1000/ms:.1f} it/s  {gb/ms*1000:.0f} GB/s algorithmic  iters={st.iterations} delta={st.delta:.3e}")
    del s, sc
