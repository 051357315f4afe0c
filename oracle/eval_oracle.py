"""CPU ORACLE of the evaluation metrics (test infrastructure, not product code): numpy restatement of the reference's
lib/evaluate/eval.py (compute_recall_at_k :73-99, compute_ap :20-70, eval_svol :102-117) and lib/evaluate/utils.py
(compute_iou_batch_paired :36-73, compute_average_precision_detection :118-202, interpolated_precision_recall
:98-115) on flat arrays.  Pinned to tests/golden/eval_*.npz, which tests/golden/make_golden_eval.py records by running
the reference's own eval_svol on synthetic results.  Nothing here is copied from /root/reference.

Flat inputs (what the CUDA path takes):
  pred      (F, q_f, 5) float32  per evaluated frame, rows sorted by score (descending, stable): x0, y0, x1, y1, score
  gt        (S, 4) float32       ground-truth boxes, xyxy, frame-major
  gt_off    (F + 1) int          frame f owns gt[gt_off[f]:gt_off[f + 1]]
  frame_off (V + 1) int          evaluation unit v (video + sketch) owns frames frame_off[v]:frame_off[v + 1]
"""
from __future__ import annotations

import numpy as np

IOU_THDS_AP = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]       # eval.py:20-22
IOU_THDS_RECALL = [float(f"{e:.2f}") for e in np.linspace(0.1, 0.9, 9)]     # eval.py:73,93


def round4(x: np.ndarray) -> np.ndarray:
    """test.py:161 writes float(f'{e:.4f}'): the value rounded to 4 decimals, as a double."""
    return np.rint(np.asarray(x, np.float64) * 1e4) / 1e4


def iou_cross(box1: np.ndarray, box2: np.ndarray) -> np.ndarray:
    """utils.py:36-96 in float64, operation order kept: (N,4) x (M,4) -> (N,M)."""
    b1, b2 = box1[:, None, :], box2[None, :, :]
    xmin, ymin = np.maximum(b1[..., 0], b2[..., 0]), np.maximum(b1[..., 1], b2[..., 1])
    xmax, ymax = np.minimum(b1[..., 2], b2[..., 2]), np.minimum(b1[..., 3], b2[..., 3])
    inter = (xmax - xmin) * (ymax - ymin)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    union = (a1 + a2) - inter
    valid = np.logical_and(xmin <= xmax, ymin <= ymax)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(valid, inter / union, 0.0)


def iou_cross_reference_layout(box1: np.ndarray, box2: np.ndarray) -> np.ndarray:
    """compute_iou_batch_cross AS THE REFERENCE LAYS IT OUT (utils.py:76-96): it pairs np.tile(box1, (M,1)) with
    np.repeat(box2, N, axis=0) -- flat entry i is the pair (box1[i % N], box2[i // N]) -- and then reshapes the flat
    vector to (N, M) row-major, so entry [n, m] holds the IoU of pair (box1[(n*M+m) % N], box2[(n*M+m) // N]).  For
    N == 1 (recall@1, the AP matching) that is the plain cross IoU; for N > 1 and M > 1 the columns mix ground-truth
    boxes.  Reproduced literally: the metric values the reference reports depend on it."""
    N, M = box1.shape[0], box2.shape[0]
    c = iou_cross(box1, box2)
    i = np.arange(N * M).reshape(N, M)
    return c[i % N, i // N]


def max_ious(pred, gt, gt_off, k: int) -> np.ndarray:
    """eval.py:75-90: per ground-truth column, the maximum over the frame's first k predictions of the reference's
    (N, M) IoU array."""
    out = []
    for f in range(pred.shape[0]):
        g = np.asarray(gt[gt_off[f]:gt_off[f + 1]], np.float64)
        if g.shape[0] == 0:
            continue
        out.extend(iou_cross_reference_layout(round4(pred[f, :k, :4]), g).max(axis=0))
    return np.asarray(out, np.float64)


def recall_at_k(pred, gt, gt_off, k: int):
    m = max_ious(pred, gt, gt_off, k)
    rec = {str(t): float(f"{np.mean(m >= t) * 100:.2f}") for t in IOU_THDS_RECALL}     # eval.py:92-97
    return rec, float(f"{np.mean(m) * 100:.2f}")


def interpolated_ap(precision, recall) -> float:
    """utils.py:98-115 (VOC 2011 interpolation)."""
    mp = np.hstack([[0], precision, [0]])
    mr = np.hstack([[0], recall, [1]])
    for i in range(len(mp) - 1)[::-1]:
        mp[i] = max(mp[i], mp[i + 1])
    idx = np.where(mr[1::] != mr[0:-1])[0] + 1
    return float(np.sum((mr[idx] - mr[idx - 1]) * mp[idx]))


def average_precision_unit(pred, gt, gt_off, f0: int, f1: int) -> np.ndarray:
    """utils.py:118-202 for one evaluation unit (frames f0:f1): predictions of all its frames sorted by score
    (descending, stable), greedy matching per IoU threshold, AP per threshold."""
    q = pred.shape[1]
    rows = round4(pred[f0:f1].reshape(-1, 5))
    frame_of = np.repeat(np.arange(f0, f1), q)
    order = np.argsort(-rows[:, 4], kind="stable")                      # list.sort(key=-score) is stable
    n_gt = int(gt_off[f1] - gt_off[f0])
    K = len(IOU_THDS_AP)
    tp, fp = np.zeros((K, len(order))), np.zeros((K, len(order)))
    lock = -np.ones((K, n_gt))
    for i, p in enumerate(order):
        f = frame_of[p]
        g0, g1 = int(gt_off[f]), int(gt_off[f + 1])
        if g1 == g0:
            fp[:, i] = 1
            continue
        iou = iou_cross(rows[p:p + 1, :4], np.asarray(gt[g0:g1], np.float64)).reshape(-1)
        cand = iou.argsort()[::-1]
        for t, thd in enumerate(IOU_THDS_AP):
            for j in cand:
                if iou[j] < thd:
                    fp[t, i] = 1
                    break
                if lock[t, g0 - int(gt_off[f0]) + j] >= 0:
                    continue
                tp[t, i] = 1
                lock[t, g0 - int(gt_off[f0]) + j] = i
                break
            if fp[t, i] == 0 and tp[t, i] == 0:
                fp[t, i] = 1
    tpc, fpc = np.cumsum(tp, axis=1), np.cumsum(fp, axis=1)
    rec = tpc / float(n_gt)
    prec = tpc / (tpc + fpc)
    return np.array([interpolated_ap(prec[t], rec[t]) for t in range(K)])


def eval_svol(pred, gt, gt_off, frame_off) -> dict:
    """eval.py:102-117: mAP over evaluation units + recall@1 / recall@5 / mIoU over ground-truth boxes."""
    aps = np.stack([average_precision_unit(pred, gt, gt_off, int(frame_off[v]), int(frame_off[v + 1]))
                    for v in range(len(frame_off) - 1)])
    ap_thds = aps.mean(0)
    m_ap = {str(t): float(f"{100 * v:.2f}") for t, v in zip(IOU_THDS_AP, ap_thds)}
    m_ap["average"] = float(f"{100 * np.mean(ap_thds):.2f}")
    r1, miou1 = recall_at_k(pred, gt, gt_off, 1)
    r5, miou5 = recall_at_k(pred, gt, gt_off, 5)
    return {"SVOL-mAP": m_ap, "SVOL-R1": r1, "SVOL-R5": r5, "mIoU@R1": miou1, "mIoU@R5": miou5}
