"""Drop-in for the reference's ``solver/CGSolverBuffer.py`` (:3-8): four fp64 cell-centred scratch grids
``d, r, q, b`` of shape ``gres`` shared by the pressure (and, in the reference, density) solvers.
Here they are torch CUDA tensors; the pressure kernels work on them in place."""
import torch

from . import _arrays as A


class CGSolverBuffer:
    def __init__(self, gres):
        g = A.to_host_ints(gres)
        dev = A.device()
        self.d = torch.zeros(g, dtype=torch.float64, device=dev)
        self.r = torch.zeros(g, dtype=torch.float64, device=dev)
        self.q = torch.zeros(g, dtype=torch.float64, device=dev)
        self.b = torch.zeros(g, dtype=torch.float64, device=dev)
