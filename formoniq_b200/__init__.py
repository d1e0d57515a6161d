"""formoniq_b200 — B200-native Galerkin assembly + CSR SpMV behind formoniq's interfaces."""
from ._lib import (FQ_DIF_BOTH, FQ_DIF_TEST, FQ_DIF_TRIAL, FQ_LUMPED, FQ_MASS, LIB_PATH, FormoniqError)
from .api import (BilinearForm, Context, DeviceCsr, DeviceVector, ElementOperator, HodgeBlocks, LinearFormPlan, SourceForm, WeightedHodgeMass, Mesh, Report, ScalarLumpedMass,
                  StopCriterion, WhitneyComplex, WhitneyPairing, cg, cg_op, kuhn_cell_faces_host, kuhn_counts, kuhn_slab_ranges, minres,
                  minres_blockdiag, minres_op, nlocal)
from .eigen import CsrPencil, EigenError, shift_invert_lanczos, sparse_shift_invert_eigen

__all__ = [
    "FQ_MASS", "FQ_DIF_TRIAL", "FQ_DIF_TEST", "FQ_DIF_BOTH", "FQ_LUMPED", "LIB_PATH", "FormoniqError", "BilinearForm",
    "Context", "DeviceCsr", "DeviceVector", "ElementOperator", "HodgeBlocks", "LinearFormPlan", "SourceForm", "WeightedHodgeMass", "Mesh", "Report", "ScalarLumpedMass", "StopCriterion",
    "WhitneyPairing", "WhitneyComplex", "cg", "minres", "cg_op", "minres_op", "minres_blockdiag", "sparse_shift_invert_eigen",
    "shift_invert_lanczos", "CsrPencil", "EigenError", "kuhn_cell_faces_host", "kuhn_counts", "kuhn_slab_ranges", "nlocal",
]
