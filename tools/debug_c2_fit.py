"""Where does the device's config-2 fit leave the reference trajectory (tests/golden/fitstep_c2_gmm784_clf.npz)?"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _datasets  # noqa: E402
from dbgsom_b200 import SomClassifier  # noqa: E402
from dbgsom_b200.engine import DeviceEngine  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "fitstep_c2_gmm784_clf.npz"))
meta = json.loads(str(g["meta"]))
X, y = _datasets.load(meta["data"])
X = np.ascontiguousarray(X.astype(meta["cast"]))
log = dict(M=[], E=[], n=[], change=[])


class Rec(DeviceEngine):
    def epoch(self, sigma, pack_rows, entropy_error, **kw):
        log["M"].append(self.M)
        r = super().epoch(sigma, pack_rows, entropy_error, **kw)
        log["E"].append(np.array(r["error"]))
        log["n"].append(np.array(r["counts"]))
        log["change"].append(r["change"])
        return r


class Est(SomClassifier):
    def _make_engine(self, distributed=None):
        return Rec(device=self.device, bmu_backend=self.bmu_backend, strict_ties=self.strict_ties, bound_scale=self.bound_scale)


for kw in (dict(strict_ties=True), dict(strict_ties=True, bmu_backend="simt")):
    for k in log:
        log[k].clear()
    est = Est(**meta["params"], **kw)
    est.fit(X, y)
    refM = g["epoch_M"]
    print(kw, "device M:", log["M"])
    print("   ref M:", refM.tolist())
    off_e = off_n = 0
    for e, m in enumerate(log["M"]):
        rm = int(refM[e])
        Er, nr = g["E_flat"][off_e:off_e + rm], g["n_flat"][off_n:off_n + rm]
        off_e += rm
        off_n += rm
        if m != rm:
            print("   first map difference at epoch", e)
            break
        dn = np.abs(log["n"][e] - nr)
        de = np.abs(log["E"][e] - Er) / np.maximum(np.abs(Er), 1e-9)
        print(f"   epoch {e}: M {m} count diffs {int(dn.sum())} max rel E diff {de.max():.2e} change {log['change'][e]:.6g} ref {g['epoch_change'][e]:.6g}"
              f" GT {est.growing_threshold_:.6g} maxE {log['E'][e].max():.6g} ref maxE {Er.max():.6g}")
    print("   final neurons", len(est.neurons_), "ref", len(g["neurons"]))
