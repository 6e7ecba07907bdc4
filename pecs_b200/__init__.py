"""pecs_b200 -- B200-native per-step IMEX path of the PhotoElectrochemical Cell Simulator (mdh266/PECS).

Python is only the binding layer: the host-side mirror of the reference's class API is C++
(pecs_b200/csrc/host), the per-step arithmetic is hand-written CUDA for sm_100a (pecs_b200/csrc/cuda) behind the
C ABI of include/pecs_b200.h.  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import PecsError, LIB_PATH, load  # noqa: F401
from .solarcell import (SolarCellProblem, KIND_PRODUCTION, KIND_TEST_STEADY, KIND_TEST_TRANSIENT,  # noqa: F401
                        KIND_TEST_DD_POISSON, ELECTRONS, HOLES, REDUCTANTS, OXIDANTS, POISSON, PARAM_NAMES,
                        default_input_file, device_count, device_warmup)
