"""adobo.irlbpy drop-in: lanczos() (adobo/irlbpy/irlb.py:95-258) on the GPU."""
import numpy as np
import scipy.sparse as sp

from . import engine as _engine
from .engine import LanczosResult  # noqa: F401


class MatrixShapeException(Exception):
    pass


def lanczos(A, nval, tol=0.0001, maxit=50, center=None, scale=None, L=None, seed=None, eng=None):
    """Truncated SVD of a 2-D matrix by augmented implicitly restarted Lanczos bidiagonalization.
    Returns an object with U, s, V, steps, nmult like the reference.  `center` / `scale` are not
    supported (adobo never passes them, dr.py:163; the reference's handling of them is incomplete,
    irlb.py:182-189); 1-D (Hankel) input is the reference's SSA path and is out of scope."""
    if center is not None or scale is not None:
        raise Exception('adobo_b200.irlbpy.lanczos: center / scale are not supported; scale the matrix first')
    if hasattr(A, 'values') and not sp.issparse(A):
        A = A.values
    if not sp.issparse(A):
        A = np.asarray(A)
        if A.ndim != 2:
            raise MatrixShapeException('The input matrix must be 2D array')
    if min(A.shape) < 2:
        raise MatrixShapeException('The input matrix must be at least 2x2.')
    eng = eng or _engine.default_engine()
    eng.resident_key = None
    return eng.lanczos_csr(sp.csr_matrix(A), nval, tol=tol, maxit=maxit, seed=seed)
