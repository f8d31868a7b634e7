"""B200-native batch-SOM training epoch behind the DBGSOM scikit-learn API.

    from dbgsom_b200 import SomVQ, SomClassifier

The estimators need `libdbgsom_b200.so` (built by `__graft_entry__.build()` or
`python -m dbgsom_b200.build`) and a CUDA device; there is no CPU fallback.
"""

from .SomClassifier import SomClassifier
from .SomVQ import SomVQ

__all__ = ["SomVQ", "SomClassifier"]
__version__ = "0.1.0"
