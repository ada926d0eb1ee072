"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's post-forward decode and temporal NMS.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(univtg_b200/postproc.py -> univtg_decode_mr / univtg_temporal_nms) never does.

Follows, line by line in behaviour (not in code):
  * main/inference_mr.py:112-120  scores = prob[..., 0]; pred_spans = timestamp + pred_spans; scores[~timestamp_mask.bool()] = 0
  * main/inference_mr.py:146-157  spans * duration; clamp(0, duration); rows [st, ed, score]; sorted(key=score, reverse=True)
                                  (skipped with --no_sort_results); every number -> float(f"{e:.4f}")
  * main/inference_mr.py:126-136  saliency scores: .half(), cut at the number of valid clips
  * main_gradio.py:100-106        windows = (pred_spans + timestamp) * ctx_l * clip_len; argmax / topk(5) of pred_logits
  * utils/temporal_nms.py:7-22,25-74 and main/inference_mr.py:31-40  greedy temporal NMS, "IoU" = intersection / convex hull
Pinning: temporal_nms is checked against the live reference function (tests/test_oracle_vs_reference.py) and against
fixtures produced from it (tests/golden/postproc_nms.json, tests/golden/make_golden_postproc.py).  The decode block is inline code
of compute_mr_results (not callable without the dataset stack): its restatement is pinned by construction only - "parity
unpinned" for that block, as for the rest of the path.
"""
import torch


def decode_mr(pred_logits, pred_spans, timestamp, timestamp_mask, durations, sort=True):
    """-> list over samples of [[st, ed, score], ...] (Python floats, rounded to 4 decimals like the reference)."""
    prob = pred_logits.detach().to("cpu", torch.float32).clone()
    scores = prob[..., 0]
    spans = timestamp.detach().to("cpu", torch.float32) + pred_spans.detach().to("cpu", torch.float32)
    mask = timestamp_mask.detach().to("cpu").bool()
    scores[~mask] = 0
    out = []
    for b in range(spans.shape[0]):
        dur = float(durations[b])
        sp = spans[b] * dur
        sp = torch.clamp(sp, 0, dur)
        rows = torch.cat([sp, scores[b][:, None]], dim=1).tolist()
        if sort:
            rows = sorted(rows, key=lambda r: r[2], reverse=True)
        out.append([[float(f"{e:.4f}") for e in row] for row in rows])
    return out


def saliency_lists(saliency_scores, src_vid_mask):
    """-> list over samples of the fp16-rounded saliency scores of the valid clips (inference_mr.py:126-136)."""
    sal = saliency_scores.detach().to("cpu").half()
    lens = src_vid_mask.detach().to("cpu").sum(1).tolist()
    return [sal[j, :int(lens[j])].tolist() for j in range(len(lens))]


def gradio_decode(pred_logits, pred_spans, timestamp, ctx_l, clip_len, k=5):
    """main_gradio.py:100-106 for one sample: (top-1 window, top-k windows) in seconds."""
    logits = pred_logits.detach().to("cpu", torch.float32)
    windows = (pred_spans.detach().to("cpu", torch.float32) + timestamp.detach().to("cpu", torch.float32)) * ctx_l * clip_len
    top1 = windows[torch.argmax(logits)].tolist()
    _, idx = torch.topk(logits.flatten(), k=min(k, logits.numel()))
    return top1, windows[idx].tolist()


def hull_iou(a, b):
    inter = max(0, min(a[1], b[1]) - max(a[0], b[0]))
    hull = max(a[1], b[1]) - min(a[0], b[0])
    if hull == 0:
        return 0
    return 1.0 * inter / hull


def temporal_nms(predictions, nms_thd, max_after_nms=100):
    """Greedy NMS over rows [st, ed, score]: a row is kept unless an earlier KEPT row overlaps it by more than nms_thd."""
    if len(predictions) == 1:
        return predictions
    rows = sorted(predictions, key=lambda r: r[2], reverse=True)
    alive = [True] * len(rows)
    kept = []
    for i, r in enumerate(rows):
        if len(kept) >= max_after_nms:
            break
        if not alive[i]:
            continue
        kept.append([r[0], r[1], r[2]])
        for j in range(i + 1, len(rows)):
            if alive[j] and hull_iou(r, rows[j]) > nms_thd:
                alive[j] = False
    return kept


def post_processing_mr_nms(windows_per_sample, nms_thd, max_before_nms, max_after_nms):
    return [temporal_nms(w[:max_before_nms], nms_thd=nms_thd, max_after_nms=max_after_nms) for w in windows_per_sample]
