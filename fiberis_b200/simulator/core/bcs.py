"""``BoundaryCondition`` helper (API parity only — the solver never calls it).

Row-overwrite helper on a dense (A, b) pair, as in
/root/reference/src/fiberis/simulator/core/bcs.py:7-60.  Note its Neumann row is
``diag=+1, neighbour=-1``, the opposite sign of the inline Neumann row the assembly uses
(/root/reference/src/fiberis/simulator/solver/matbuilder.py:53-55); both conventions are kept
exactly as the reference has them.  Host-side NumPy on caller-owned arrays.
"""
import numpy as np


class BoundaryCondition:
    def __init__(self, type=None, value=None, idx=None):
        self.type, self.value, self.idx = type, value, idx
        self.matA = None
        self.vecB = None

    def set_matrix(self, matA, vecB):
        self.matA, self.vecB = matA, vecB
        rows_cols, vec_shape = np.shape(matA), np.shape(vecB)
        if rows_cols[0] != rows_cols[1]:
            raise ValueError("Matrix A is not square")
        if len(vec_shape) != 1:
            raise ValueError("Vector B is not a vector")
        if rows_cols[0] != vec_shape[0]:
            raise ValueError("The size of matrix A and vector B does not match")

    def apply(self):
        """Overwrite row ``idx`` of A and entry ``idx`` of b in place."""
        k = self.idx
        if self.type == "PML":
            return None  # placeholder in the reference as well
        if self.type not in ("Dirichlet", "Neumann"):
            raise ValueError("Unknown BC type")
        self.matA[k, :] = 0
        self.matA[k, k] = 1
        if self.type == "Neumann":
            self.matA[k, k + 1 if k == 0 else k - 1] = -1
        self.vecB[k] = self.value
        return None
