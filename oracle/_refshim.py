"""Import the UNMODIFIED reference (SandroMartens/DBGSOM) in the build container.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that
runs there (`-m gpu` tests, smoke(), bench.py) may import this module; it is used by
`tests/golden/make_golden.py` to generate the committed fixtures and by the
`needs_reference` tests that re-validate the oracle when the reference is present.

Three shims are needed:
  1. `dbgsom/BaseSom.py:14-36` imports seaborn inside a try that sys.exit()s -> stub modules;
  2. `dbgsom/SomClassifier.py:8-14` imports `check_X_y` from `sklearn.base`, which
     scikit-learn 1.9 no longer re-exports -> alias it;
  3. `numba_quantization_error` (`dbgsom/BaseSom.py:1058-1073`) races under >1 numba thread
     (quirk Q2) -> NUMBA_NUM_THREADS=1 must be set before numba is imported.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DBGSOM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dbgsom"))


def load():
    """Return (SomVQ, SomClassifier, BaseSom_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "numba" in sys.modules and os.environ.get("NUMBA_NUM_THREADS") != "1":
        raise RuntimeError("set NUMBA_NUM_THREADS=1 before numba is imported (quirk Q2)")
    os.environ["NUMBA_NUM_THREADS"] = "1"
    for m in ("seaborn", "seaborn.objects"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["seaborn"].objects = sys.modules["seaborn.objects"]
    import sklearn.base
    import sklearn.utils

    if not hasattr(sklearn.base, "check_X_y"):
        sklearn.base.check_X_y = sklearn.utils.check_X_y
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from dbgsom import BaseSom as ref_base
    from dbgsom.SomClassifier import SomClassifier
    from dbgsom.SomVQ import SomVQ

    return SomVQ, SomClassifier, ref_base
