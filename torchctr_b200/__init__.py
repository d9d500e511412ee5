"""torchctr_b200 -- B200-native (sm_100a) embedding / feature-interaction hot path behind the
torchctr module API.  Kernels live in ``csrc/`` behind the C ABI of ``include/ctr_b200.h``;
``ops`` wraps them on tensors, ``nn`` / ``models`` mirror ``torchctr.nn`` / ``torchctr.models``."""
__version__ = "0.1.0"

from . import _lib, ops  # noqa: F401
from . import nn, models  # noqa: F401
from .models import DNN, DeepFM, DCNv2  # noqa: F401
