"""Vector-quantisation / clustering estimator (drop-in for `dbgsom.SomVQ.SomVQ`)."""

from __future__ import annotations

import numpy as np
from sklearn.base import ClusterMixin, TransformerMixin
from sklearn.utils import check_array
from sklearn.utils.validation import check_is_fitted

from .BaseSom import BaseSom


class SomVQ(BaseSom, ClusterMixin, TransformerMixin):
    """Directed batch growing SOM used as a vector quantiser, trained on a B200.

    Same hyper-parameters, methods and fitted attributes as the reference class
    (dbgsom/SomVQ.py:16-152): `labels_`, `som_`, `weights_`, `neurons_`,
    `topographic_error_`, `quantization_error_`.  `predict` returns the index of the best
    matching prototype (row of `weights_`).
    """

    def _check_input_data(self, X, y=None):
        # dbgsom/SomVQ.py:121-124 -- any y is dropped
        X = check_array(array=X, ensure_min_samples=4, dtype=[np.float64, np.float32])
        return X, None

    def _label_prototypes(self, y, engine) -> None:
        # dbgsom/SomVQ.py:126-128 -- label == ordinal of the node
        for i, node in enumerate(self.som_):
            self.som_.nodes[node]["label"] = i

    def predict(self, X) -> np.ndarray:
        """Index of the closest prototype per sample (dbgsom/SomVQ.py:130-148)."""
        check_is_fitted(self)
        X = check_array(X)
        _, labels = self._get_winning_neurons(X, n_bmu=1)
        return labels

    def _fit(self, engine) -> None:
        # dbgsom/SomVQ.py:150-152: labels_ = predict(X); removed neurons never won a
        # sample, so the training winners on the reduced map are that result
        self.labels_ = np.asarray(engine.winners_host(), dtype=np.int64)
