"""B200-native drop-in for the reference's ``solver/PressureCGSolver3D.py``.

Module functions ``initialize_solver``, ``matvecmul``, ``apply_pressure`` keep the reference's argument
order (:155, :161, :167); ``class PressureCGSolver3D(buf, gres, bound_size)`` / ``solve(vx, vy, vz, sphi, sv,
lphi, wx=None, wy=None, wz=None, tol=1e-3)`` keeps its surface (:173-226).  fp64 state, as the reference."""
from . import _pressure as P
from .SolidFraction3D import compute_solid_frac, edge_in_fraction  # noqa: F401  (same re-exports as the reference :4)


def initialize_solver(cell_size, gres, vx, vy, vz, sphi, sv, lphi, b, wx, wy, wz):
    P.initialize_solver(cell_size, gres, (vx, vy, vz), sphi, sv, lphi, b, (wx, wy, wz))


def matvecmul(gres, v, out, wx, wy, wz, lphi):
    P.matvecmul(gres, v, out, (wx, wy, wz), lphi)


def apply_pressure(gres, cell_size, vx, vy, vz, pv, wx, wy, wz, sv, lphi):
    P.apply_pressure(gres, cell_size, (vx, vy, vz), pv, (wx, wy, wz), sv, lphi)


class PressureCGSolver3D(P.PressureSolverBase):
    _dim = 3
    _raise_on_fail = True      # for ... else: raise ValueError("Failed to converge!")  (:222-223)

    def solve(self, vx, vy, vz, sphi, sv, lphi, wx=None, wy=None, wz=None, tol=1e-3):
        ws = None if (wx is None or wy is None or wz is None) else (wx, wy, wz)   # :193
        self._solve((vx, vy, vz), sphi, sv, lphi, ws, tol)
