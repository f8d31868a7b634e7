"""Classifier estimator (drop-in for `dbgsom.SomClassifier.SomClassifier`)."""

from __future__ import annotations

import numpy as np
from sklearn.base import ClassifierMixin, TransformerMixin
from sklearn.utils import check_array, check_X_y
from sklearn.utils.validation import check_is_fitted

from .BaseSom import BaseSom


class SomClassifier(BaseSom, TransformerMixin, ClassifierMixin):
    """Directed batch growing SOM used as a classifier, trained on a B200.

    Same hyper-parameters, methods and fitted attributes as the reference class
    (dbgsom/SomClassifier.py:19-220).  Training is the same device epoch as `SomVQ`; `y`
    only enters through prototype labelling and the optional entropy growth criterion.
    `predict_proba` follows the reference: sparse non-negative code of the sample over the
    prototypes times the per-prototype class frequencies, row-normalised (host path).
    """

    def _check_input_data(self, X, y):
        # dbgsom/SomClassifier.py:124-128
        X, y = check_X_y(X=X, y=y, ensure_min_samples=4, dtype=[np.float64, np.float32])
        return X, y

    def _label_prototypes(self, y, engine) -> None:
        """Majority class and class frequencies per prototype (dbgsom/SomClassifier.py:130-152).

        `statistics.mode` returns, among the most frequent classes, the one met first in
        sample order; the per-(prototype, class) first-occurrence index reproduces that.
        The class histogram and first occurrences are reduced on the device (`dbgsom_label_hist`).
        """
        m, c = len(self.neurons_), len(self.classes_)
        counts, first = engine.label_histogram(c)
        counts, first = counts.reshape(m, c), first.reshape(m, c)
        for j, node in enumerate(self.neurons_):
            row = counts[j]
            hit = row.sum()
            attrs = self.som_.nodes[node]
            prob = np.zeros(shape=self.classes_.shape)
            if hit == 0:
                # a survivor of the dead-neuron removal (hit_count > 0 on the PRE-update prototypes) that wins no
                # sample on the updated map: the reference writes probabilities[-1] = 0 / hit_count, i.e. an
                # all-zero row (dbgsom/SomClassifier.py:136-152)
                attrs["label"] = -1
                prob[-1] = 0.0 if attrs["hit_count"] > 0 else 1
            else:
                best = np.flatnonzero(row == row.max())
                attrs["label"] = int(best[np.argmin(first[j, best])])
                hit_count = attrs["hit_count"]
                nz = row > 0
                prob[nz] = row[nz] / hit_count if hit_count > 0 else 1
            attrs["probabilities"] = prob

    def predict(self, X) -> np.ndarray:
        """Predicted class per sample (dbgsom/SomClassifier.py:159-176)."""
        check_is_fitted(self)
        X = check_array(X)
        labels = np.argmax(self.predict_proba(X=X), axis=1)
        return self.classes_[labels]

    def predict_proba(self, X) -> np.ndarray:
        """Class probabilities per sample (dbgsom/SomClassifier.py:178-220, flat map)."""
        check_is_fitted(self)
        X = check_array(X)
        code = self.transform(X)
        node_prob = self._extract_values_from_graph("probabilities")
        raw = code @ node_prob
        return raw / raw.sum(axis=1)[np.newaxis].T
