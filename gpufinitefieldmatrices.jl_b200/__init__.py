"""gpufinitefieldmatrices.jl_b200 -- B200-native (sm_100a) hot path of GPUFiniteFieldMatrices.jl behind the
reference's CuModMatrix / KaratsubaMatrix interface.  Import as `gffm_b200` (see gffm_b200.py at the repo root;
the directory name contains a dot and cannot be imported directly)."""
from . import capi
from .capi import (CuModArrayModulusMismatchException, CuModArraySizeMismatchException, CuModMatrixModulusNotPrimeException, CuModMatrixNotSquareException,
                   GffmError, InexactError, InverseNotDefinedException, MatrixNotInvertibleException)
from .cumodmatrix import *  # noqa: F401,F403
from .cumodmatrix import Context, CuModMatrix, CuModVector, default_context
from . import karatsuba
from . import multigpu
from .karatsuba import (KaratsubaArray, KaratsubaMatrix, KaratsubaVector, KaratsubaZeros, Karatsubacopy, KMatMul_, KMatMul_gemv_, KMatToMat,
                        MatToKMat, initialize_plan_)

__all__ = [n for n in dir() if not n.startswith("_")]
