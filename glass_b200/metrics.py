"""impl/metrics.py equivalents (host-side numpy scores; not on the hot path)."""
import numpy as np


def _micro_f1(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """Micro-averaged F1 for multi-class (== accuracy) or multi-label indicator arrays."""
    if y_true.ndim == 1:
        return float(np.mean(y_true == y_pred))
    tp = float(np.sum((y_true == 1) & (y_pred == 1)))
    fp = float(np.sum((y_true == 0) & (y_pred == 1)))
    fn = float(np.sum((y_true == 1) & (y_pred == 0)))
    return 0.0 if tp == 0 else 2 * tp / (2 * tp + fp + fn)


def binaryf1(pred, label):
    """impl/metrics.py:5-12: threshold logits at 0, then sklearn `f1_score(label[n, -1], pred, average="micro")`.
    sklearn treats a single-column target as a BINARY problem and micro-averages over both classes, which is
    the accuracy; only a real multi-label indicator matrix (more than one column) gets the tp/fp/fn formula."""
    pred_i = (np.asarray(pred) > 0).astype(np.int64)
    label_i = np.asarray(label).reshape(pred_i.shape[0], -1).astype(np.int64)
    if label_i.shape[1] == 1:
        return float(np.mean(label_i.reshape(-1) == pred_i.reshape(-1)))
    return _micro_f1(label_i, pred_i.reshape(label_i.shape))


def microf1(pred, label):
    """impl/metrics.py:15-20."""
    return _micro_f1(np.asarray(label), np.argmax(pred, axis=1))


def auroc(pred, label):
    """impl/metrics.py:23-27."""
    from sklearn.metrics import roc_auc_score
    return roc_auc_score(label, pred)
