"""Drop-in replacement for the reference's ``solver`` package (hot path only).

Same module and class names as ``/root/reference/solver`` so that
``from solver.ViscosityCGSolver3D import ViscosityCGSolver3D`` etc. (3D_viscous_fluid_sim.ipynb:613-616)
keep working with ``python-fluid-simulation_b200/`` on ``sys.path``; every kernel underneath is
hand-written sm_100a CUDA reached through the C ABI in ``include/fluidsolver_b200.h``.
There is no CPU fallback: importing a solver module without the built library raises.
"""
