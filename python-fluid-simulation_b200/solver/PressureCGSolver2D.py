"""B200-native drop-in for the reference's ``solver/PressureCGSolver2D.py`` (:122-179).

The 2-D CG loop of the reference has no ``else: raise`` (:165-177): on iteration exhaustion it silently
applies whatever pressure it has.  That behaviour is kept."""
from . import _pressure as P
from .SolidFraction2D import compute_solid_frac, edge_in_fraction  # noqa: F401


def initialize_solver(cell_size, gres, vx, vy, sphi, sv, lphi, b, wx, wy):
    P.initialize_solver(cell_size, gres, (vx, vy), sphi, sv, lphi, b, (wx, wy))


def matvecmul(gres, v, out, wx, wy, lphi):
    P.matvecmul(gres, v, out, (wx, wy), lphi)


def apply_pressure(gres, cell_size, vx, vy, pv, wx, wy, sv, lphi):
    P.apply_pressure(gres, cell_size, (vx, vy), pv, (wx, wy), sv, lphi)


class PressureCGSolver2D(P.PressureSolverBase):
    _dim = 2
    _raise_on_fail = False

    def solve(self, vx, vy, sphi, sv, lphi, wx=None, wy=None, tol=1e-3):
        ws = None if (wx is None or wy is None) else (wx, wy)
        self._solve((vx, vy), sphi, sv, lphi, ws, tol)
