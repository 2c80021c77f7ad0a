"""ikarus_b200: B200-native drop-in for Ikarus' global FEM assembly hot path.

Host-side mirror of the reference's assembler interface over the C-ABI in include/ikb200.h.
"""
from .assembler import (AffordanceCollection, DBCOption, DenseFlatAssembler, DeviceMatrix, FERequirements,  # noqa: F401
                        MatrixAffordance, ResultTypes, ScalarAffordance, SparseFlatAssembler, VectorAffordance,
                        elastoStatics,
                        makeDenseFlatAssembler, makeSparseFlatAssembler)
from .fe import (DirichletValues, FEContainer, Materials, eas, linearElastic, makeFE, neumannBoundaryLoad,  # noqa: F401
                 nonLinearElastic, planeStrain, planeStress, skills, toLamesFirstParameterAndShearModulus, volumeLoad)
from .io import ResultFunction, makeResultFunction  # noqa: F401
from .solvers import (ControlInformation, DeviceLinearSolver, DeviceTruncatedCG, LoadControl,  # noqa: F401
                      LoadControlConfig, NewtonRaphson, NewtonRaphsonConfig, NonLinearSolverInformation, NRSettings,
                      PreConditioner, TrustRegion, TRSettings, obtainForcesDueToIDBC)

__all__ = [n for n in dir() if not n.startswith("_")]
