"""B200-native residual/Jacobian evaluation engine for Ceres (host-side Python glue).

The product is C++/CUDA (``csrc/``, ``include/``); Python here only generates
workloads (``problems``) and binds the test driver / C ABI for pytest and bench.py
(``binding``)."""
from . import problems  # noqa: F401
from . import binding  # noqa: F401
from . import lm  # noqa: F401
