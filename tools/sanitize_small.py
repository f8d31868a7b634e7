"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbgsom_b200 import SomClassifier  # noqa: E402
from dbgsom_b200.engine import DeviceEngine  # noqa: E402

rng = np.random.default_rng(0)
X = rng.normal(size=(700, 40)).astype(np.float32)
y = rng.integers(0, 3, 700)
for backend in ("simt", "tensor", "tensor1"):
    est = SomClassifier(n_iter=6, max_neurons=12, random_state=0, bmu_backend=backend).fit(X, y)
    print(backend, len(est.neurons_), est.quantization_error_)
side = 16
hop = np.abs(np.arange(side * side)[:, None] // side - np.arange(side * side)[None, :] // side) + np.abs(
    np.arange(side * side)[:, None] % side - np.arange(side * side)[None, :] % side)
for d in (128, 320):
    Xb = rng.normal(size=(1500, d)).astype(np.float32)
    e = DeviceEngine(bmu_backend="tensor")
    e.load_data(Xb, None, 0)
    e.set_map(Xb[: side * side].astype(np.float64))
    e.set_hops(hop.astype(np.uint16))
    print(d, e.epoch(2.0, True, False)["change"])
    e.close()
