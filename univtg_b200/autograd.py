"""Training-mode forward/backward glue: one torch.autograd.Function around univtg_forward_train / univtg_backward.

Autograd sees a single node whose inputs are the model parameters (in C-ABI order) and whose outputs are the four
differentiable tensors of the reference's output dict; gradients come back through one flat fp32 buffer (one view per
parameter), which is also what the data-parallel all-reduce operates on (univtg_b200/ddp.py).
"""
import ctypes

import torch

from . import _lib


def _draw_randomness(model, B, Lv, Lt, dev):
    """DropPath scales [2 * enc_layers, B] and input-dropout multipliers (one per projector layer and modality), drawn with
    torch's generator (reference model/univtg.py:107-108,394; transformer_encoder_droppath.py:154-167).

    Default: batched draws - one uniform tensor for all DropPath scales and a Bernoulli fill per dropout mask - a handful of
    launches per step.  `model.reference_rng_order = True` issues the reference's own calls in the reference's order instead
    (F.dropout on ones per projector layer, then one torch.rand((B, 1, 1)) per DropPath site: ~45 tiny launches per step)."""
    n = model.n_input_proj
    masks = [None] * (2 * n)
    p = model.input_dropout
    ref_order = bool(getattr(model, "reference_rng_order", False))
    if p > 0.0:
        dims_v = [model.vid_dim] + [model.hidden_dim] * 3
        dims_t = [model.txt_dim] + [model.hidden_dim] * 3
        shapes = [(B, Lv, dims_v[i]) for i in range(n)] + [(B, Lt, dims_t[i]) for i in range(n)]
        for i, shp in enumerate(shapes):
            if ref_order:
                masks[i] = torch.nn.functional.dropout(torch.ones(shp, device=dev), p, True).contiguous()
            else:
                masks[i] = torch.empty(shp, dtype=torch.float32, device=dev).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))
    scales = None
    if model.droppath > 0.0:
        keep = 1.0 - model.droppath
        if ref_order:
            rows = []
            for _ in range(2 * model.enc_layers):
                m = keep + torch.rand((B, 1, 1), dtype=torch.float32, device=dev)
                rows.append(m.floor_().flatten() / keep)
            scales = torch.stack(rows).contiguous()
        else:
            scales = torch.rand((2 * model.enc_layers, B), dtype=torch.float32, device=dev).add_(keep).floor_().div_(keep)
    if model.attn_dropout > 0.0:
        raise NotImplementedError("attention dropout > 0 is not supported (every reference script sets --dropout 0)")
    return scales, masks


class _UniVTGFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, src_txt, src_txt_mask, src_vid, src_vid_mask, *params):
        lib = _lib.load_library()
        dev = model._device()
        B, Lv, _ = src_vid.shape
        Lt = src_txt.shape[1]
        d = model.hidden_dim
        with torch.cuda.device(dev):
            model._ensure_packed(training=True)
            plan = model._get_plan(B, Lv, Lt, True)
            ws = model._get_train_ws(B, Lv, Lt)
            txt = src_txt.detach().to(torch.float32).contiguous()
            vid = src_vid.detach().to(torch.float32).contiguous()
            tmask = src_txt_mask.detach().to(torch.float32).contiguous()
            vmask = src_vid_mask.detach().to(torch.float32).contiguous()
            scales, masks = _draw_randomness(model, B, Lv, Lt, dev)
            if getattr(model, "keep_last_draw", False):  # parity tests hand the same multipliers to the oracle
                model.__dict__["_last_draw"] = (scales, masks)
            mask_arr = None
            if any(m is not None for m in masks):
                mask_arr = (ctypes.c_void_p * len(masks))(*[m.data_ptr() if m is not None else None for m in masks])
            pred_logits = torch.empty(B, Lv, 1, device=dev)
            pred_spans = torch.empty(B, Lv, 2, device=dev)
            vid_mem_proj = torch.empty(B, Lv, d, device=dev)
            txt_mem_proj = torch.empty(B, 1, d, device=dev)
            saliency = torch.empty(B, Lv, device=dev)
            _lib.check(lib.univtg_forward_train(plan.handle, _lib.ptr(ws), _lib.ptr(txt), _lib.ptr(tmask), _lib.ptr(vid),
                                                _lib.ptr(vmask), _lib.ptr(scales), mask_arr, _lib.ptr(pred_logits),
                                                _lib.ptr(pred_spans), _lib.ptr(vid_mem_proj), _lib.ptr(txt_mem_proj),
                                                _lib.ptr(saliency), _lib.stream_ptr()), "univtg_forward_train")
        ctx.model = model
        ctx.plan = plan
        ctx.ws = ws
        ctx.shape = (B, Lv, Lt)
        ctx.saved = (txt, vid, scales, masks, mask_arr)
        ctx.mark_non_differentiable(saliency)
        return pred_logits, pred_spans, vid_mem_proj, txt_mem_proj, saliency

    @staticmethod
    def backward(ctx, g_logits, g_spans, g_vmp, g_tmp, _g_sal):
        lib = _lib.load_library()
        model = ctx.model
        dev = model._device()
        txt, vid, scales, masks, mask_arr = ctx.saved
        params = model._abi_params()
        with torch.cuda.device(dev):
            flat, views = model._grad_buffer()
            flat.zero_()

            def prep(g):
                return None if g is None else g.detach().to(torch.float32).contiguous()

            g_logits, g_spans, g_vmp, g_tmp = prep(g_logits), prep(g_spans), prep(g_vmp), prep(g_tmp)
            if (g_logits is None) != (g_spans is None):  # the heads' backward consumes both together
                B, Lv, _ = ctx.shape
                g_logits = g_logits if g_logits is not None else torch.zeros(B, Lv, 1, device=dev)
                g_spans = g_spans if g_spans is not None else torch.zeros(B, Lv, 2, device=dev)
            arr = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
            sync = getattr(model, "_grad_sync", None)  # univtg_b200.ddp.OverlappedGradExchange
            if sync is not None:
                sync.before_backward(ctx.plan)
            _lib.check(lib.univtg_backward(ctx.plan.handle, _lib.ptr(ctx.ws), _lib.ptr(txt), _lib.ptr(vid), _lib.ptr(scales),
                                           mask_arr, _lib.ptr(g_logits), _lib.ptr(g_spans), _lib.ptr(g_vmp), _lib.ptr(g_tmp),
                                           float(model.grad_scale), arr, len(views), _lib.stream_ptr()), "univtg_backward")
            hook = getattr(model, "_flat_grad_hook", None)
            if sync is not None:
                sync.after_backward(flat)  # stage-wise all-reduce on a side stream, chained behind the stage events
            elif hook is not None:
                hook(flat)  # e.g. the single NCCL all-reduce of univtg_b200.ddp
        if getattr(model, "direct_grad", False):
            # Hand the flat buffer's views to param.grad directly: no per-parameter AccumulateGrad copies (77 memcpys / step).
            # Used with univtg_b200.ddp (one flat all-reduce); torch DDP needs the autograd route below.
            for v, p in zip(views, params):
                if not p.requires_grad:
                    continue
                if p.grad is None:
                    p.grad = v
                elif p.grad.data_ptr() != v.data_ptr():
                    p.grad.add_(v)
            return (None,) * (5 + len(params))
        grads = tuple(v if p.requires_grad else None for v, p in zip(views, params))
        return (None, None, None, None, None) + grads


def forward_train(model, src_txt, src_txt_mask, src_vid, src_vid_mask):
    params = model._abi_params()
    pred_logits, pred_spans, vmp, tmp, sal = _UniVTGFunction.apply(model, src_txt, src_txt_mask, src_vid, src_vid_mask, *params)
    return {"pred_logits": pred_logits, "pred_spans": pred_spans, "src_vid_mask": src_vid_mask, "vid_mem_proj": vmp,
            "txt_mem_proj": tmp, "saliency_scores": sal}
