"""VOC detection metrics -- the API of the reference's ``utility/metrics.py`` (:102-261): streaming TP/FP arrays,
precision / recall, AP (VOC07 11-point and VOC12 area).  The reference accumulates these in TF local variables on the
CPU; they are a few thousand booleans per class, so this is host-side numpy in float64 like the reference's own
``dtype=tf.float64`` -- not a kernel.
"""
import numpy as np


class StreamingTpFpArrays(object):
    """``streaming_tp_fp_arrays`` (utility/metrics.py:135-199) for one class: keeps (n_objects, n_detections, tp, fp,
    scores) over batches; detections that are neither TP nor FP, or whose score is <= 1e-4, are dropped (:170-176)."""

    def __init__(self, remove_zero_scores=True):
        self.remove_zero_scores = remove_zero_scores
        self.nobjects = 0
        self.ndetections = 0
        self.tp = np.zeros((0,), bool)
        self.fp = np.zeros((0,), bool)
        self.scores = np.zeros((0,), np.float32)

    def update(self, num_gbboxes, tp, fp, scores):
        tp = np.asarray(tp, bool).reshape(-1)
        fp = np.asarray(fp, bool).reshape(-1)
        scores = np.asarray(scores, np.float32).reshape(-1)
        mask = tp | fp
        if self.remove_zero_scores:
            mask &= scores > np.float32(1e-4)
            scores, tp, fp = scores[mask], tp[mask], fp[mask]
        self.nobjects += int(np.sum(np.asarray(num_gbboxes, np.int64)))
        self.ndetections += int(scores.shape[0])
        self.scores = np.concatenate([self.scores, scores])
        self.tp = np.concatenate([self.tp, tp])
        self.fp = np.concatenate([self.fp, fp])

    def value(self):
        return self.nobjects, self.ndetections, self.tp, self.fp, self.scores


def streaming_tp_fp_arrays(num_gbboxes, tp, fp, scores, remove_zero_scores=True, metrics_collections=None,
                           updates_collections=None, name=None, state=None):
    """Dictionary form of the reference (:143-157): one ``StreamingTpFpArrays`` per class, updated in place.
    ``state`` is the dict returned by a previous call (None starts a new accumulation); the TF collection / name
    arguments of the reference signature are accepted and unused."""
    state = {} if state is None else state
    for c in num_gbboxes:
        state.setdefault(c, StreamingTpFpArrays(remove_zero_scores)).update(
            _np(num_gbboxes[c]), _np(tp[c]), _np(fp[c]), _np(scores[c]))
    return state


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def precision_recall(num_gbboxes, num_detections, tp, fp, scores, dtype=np.float64, scope=None):
    """utility/metrics.py:102-132: sort by score (tf.nn.top_k: descending, ties -> lower index), cumulative TP / FP."""
    scores = np.asarray(scores, np.float32)[:num_detections]
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    tp = np.cumsum(np.asarray(tp, bool)[order].astype(dtype))
    fp = np.cumsum(np.asarray(fp, bool)[order].astype(dtype))
    recall = np.where(num_gbboxes > 0, tp / max(float(num_gbboxes), 1e-300), 0.0).astype(dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = np.where(tp + fp > 0, tp / (tp + fp), 0.0).astype(dtype)
    return precision, recall


def average_precision_voc12(precision, recall, name=None):
    """utility/metrics.py:205-227."""
    precision = np.concatenate([[0.], np.asarray(precision, np.float64), [0.]])
    recall = np.concatenate([[0.], np.asarray(recall, np.float64), [1.]])
    precision = np.maximum.accumulate(precision[::-1])[::-1]  # cummax(reverse=True)
    return float(np.sum(precision[1:] * (recall[1:] - recall[:-1])))


def average_precision_voc07(precision, recall, name=None):
    """utility/metrics.py:230-252: 11-point interpolation."""
    precision = np.concatenate([np.asarray(precision, np.float64), [0.]])
    recall = np.concatenate([np.asarray(recall, np.float64), [np.inf]])
    ap = 0.0
    for t in np.arange(0., 1.1, 0.1):
        ap += float(np.max(precision[recall >= t])) / 11.
    return ap


def voc_map(state, use_07_metric=True):
    """mAP over the classes of a ``streaming_tp_fp_arrays`` state, as the eval script reports it
    (light_head_rfcn_eval.py:300-333: per-class AP, then their mean)."""
    aps = {}
    for c, acc in state.items():
        nobj, ndet, tp, fp, scores = acc.value()
        p, r = precision_recall(nobj, ndet, tp, fp, scores)
        aps[c] = average_precision_voc07(p, r) if use_07_metric else average_precision_voc12(p, r)
    return (float(np.mean(list(aps.values()))) if aps else 0.0), aps
