"""Deterministic synthetic weights / inputs / targets of the shapes the reference feeds this path.

Input conventions follow the reference data pipeline: row-L2-normalised features + TEF columns
(main/dataset.py:534-540, 689-695), zero right-padding with float masks (utils/tensor_utils.py:36-53), dense per-clip
targets (main/dataset.py:501, 525-556, 1078-1098).  Seeds: weights torch.Generator(seed), data Generator(seed+1).
"""
import math
from argparse import Namespace

import torch

CONFIGS = {
    # BASELINE.json configs[0]: gradio demo shapes (tmp/vid.npz 15x512(+2 TEF), tmp/txt.npz 12x512), d=256, 2 layers
    "cfg1": dict(hidden_dim=256, nheads=8, dim_feedforward=1024, enc_layers=2, n_input_proj=2, v_feat_dim=514, t_feat_dim=512,
                 batch=1, l_vid=15, l_txt=12),
    # configs[1] / [2]: QVHighlights-shaped
    "cfg2": dict(hidden_dim=1024, nheads=8, dim_feedforward=1024, enc_layers=4, n_input_proj=2, v_feat_dim=2818, t_feat_dim=512,
                 batch=32, l_vid=75, l_txt=32),
    # configs[3]: per-rank shard of the vlp_ddp pre-training batch
    "cfg4": dict(hidden_dim=1024, nheads=8, dim_feedforward=1024, enc_layers=4, n_input_proj=2, v_feat_dim=2818, t_feat_dim=512,
                 batch=32, l_vid=150, l_txt=32),
    # configs[4]: long-video stress
    "cfg5": dict(hidden_dim=1024, nheads=8, dim_feedforward=1024, enc_layers=6, n_input_proj=2, v_feat_dim=2818, t_feat_dim=512,
                 batch=8, l_vid=1200, l_txt=77),
    # small ragged parity case the oracle finishes in < 1 s
    "tiny": dict(hidden_dim=256, nheads=2, dim_feedforward=256, enc_layers=2, n_input_proj=2, v_feat_dim=194, t_feat_dim=128,
                 batch=3, l_vid=21, l_txt=9),
}


def reference_args(cfg, **over):
    """argparse.Namespace with every field reference build_model(args) reads (model/univtg.py:409-448)."""
    ns = Namespace(
        device="cpu", hidden_dim=cfg["hidden_dim"], dropout=0.0, droppath=0.1, nheads=cfg["nheads"],
        dim_feedforward=cfg["dim_feedforward"], enc_layers=cfg["enc_layers"], dec_layers=2, pre_norm=False,
        position_embedding="sine", max_q_l=max(75, cfg.get("l_txt", 32)), input_dropout=0.5, t_feat_dim=cfg["t_feat_dim"],
        v_feat_dim=cfg["v_feat_dim"], span_loss_type="l1", use_txt_pos=False, n_input_proj=cfg["n_input_proj"],
        set_cost_span=10, set_cost_giou=1, set_cost_class=4, max_v_l=max(75, cfg.get("l_vid", 75)), b_loss_coef=10.0,
        g_loss_coef=1.0, f_loss_coef=10.0, s_loss_intra_coef=0.1, s_loss_inter_coef=0.1, dset_type="vlp",
        train_path=["synthetic"], eos_coef=0.1, temperature=0.07, saliency_margin=0.2)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def state_dict_shapes(cfg, max_q_l=None):
    """Reference state_dict keys -> shapes, in the reference's registration order."""
    d, ff, N, n = cfg["hidden_dim"], cfg["dim_feedforward"], cfg["enc_layers"], cfg["n_input_proj"]
    max_q_l = max_q_l or max(75, cfg.get("l_txt", 32))
    out = {}
    for l in range(N):
        p = f"transformer.encoder.layers.{l}."
        out[p + "self_attn.in_proj_weight"] = (3 * d, d)
        out[p + "self_attn.in_proj_bias"] = (3 * d,)
        out[p + "self_attn.out_proj.weight"] = (d, d)
        out[p + "self_attn.out_proj.bias"] = (d,)
        out[p + "linear1.weight"] = (ff, d)
        out[p + "linear1.bias"] = (ff,)
        out[p + "linear2.weight"] = (d, ff)
        out[p + "linear2.bias"] = (d,)
        out[p + "norm1.weight"] = (d,)
        out[p + "norm1.bias"] = (d,)
        out[p + "norm2.weight"] = (d,)
        out[p + "norm2.bias"] = (d,)
    out["txt_position_embed.position_embeddings.weight"] = (max_q_l, d)
    out["txt_position_embed.LayerNorm.weight"] = (d,)
    out["txt_position_embed.LayerNorm.bias"] = (d,)
    out["token_type_embeddings.weight"] = (2, d)
    for head, od in (("span_embed", 2), ("class_embed", 1)):
        out[f"{head}.layers.0.weight"] = (d, d, 3)
        out[f"{head}.layers.0.bias"] = (d,)
        out[f"{head}.layers.1.weight"] = (d, d, 3)
        out[f"{head}.layers.1.bias"] = (d,)
        out[f"{head}.layers.2.weight"] = (od, d, 3)
        out[f"{head}.layers.2.bias"] = (od,)
    for name, din in (("input_txt_proj", cfg["t_feat_dim"]), ("input_vid_proj", cfg["v_feat_dim"])):
        k = din
        for i in range(n):
            out[f"{name}.{i}.LayerNorm.weight"] = (k,)
            out[f"{name}.{i}.LayerNorm.bias"] = (k,)
            out[f"{name}.{i}.net.1.weight"] = (d, k)
            out[f"{name}.{i}.net.1.bias"] = (d,)
            k = d
    out["weightedpool.weight"] = (d, 1)
    return out


def make_state_dict(cfg, seed=0, head_gain=1.0, dtype=torch.float32):
    """Seeded weights with the reference's init scales (xavier-uniform encoder matrices, U(+-1/sqrt(fan_in)) elsewhere)
    but non-trivial LayerNorm affine terms and biases, so that every parameter matters in a parity test."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k.endswith("LayerNorm.weight") or ".norm1.weight" in k or ".norm2.weight" in k:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("LayerNorm.bias") or ".norm1.bias" in k or ".norm2.bias" in k:
            t = 0.05 * torch.randn(shp, generator=g)
        elif k == "token_type_embeddings.weight":
            t = 0.02 * torch.randn(shp, generator=g)
        elif k == "txt_position_embed.position_embeddings.weight":
            t = torch.randn(shp, generator=g)
        elif k.startswith("transformer.") and len(shp) == 2:
            bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif k == "weightedpool.weight":
            bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif len(shp) >= 2:
            fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
            if k.endswith("layers.2.weight"):
                t = t * head_gain
        else:  # biases
            t = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        sd[k] = t.to(dtype)
    return sd


def make_inputs(cfg, seed=1, ragged=False, batch=None, l_vid=None, l_txt=None):
    """src_txt [B,Lt,Dt], src_txt_mask [B,Lt], src_vid [B,Lv,Dv], src_vid_mask [B,Lv] (float32, zero right-padded)."""
    B = batch or cfg["batch"]
    Lv = l_vid or cfg["l_vid"]
    Lt = l_txt or cfg["l_txt"]
    Dv, Dt = cfg["v_feat_dim"], cfg["t_feat_dim"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw = torch.randn(B, Lv, Dv - 2, generator=g)
    split = (Dv - 2) * 9 // 11 if Dv - 2 >= 11 else (Dv - 2) // 2  # 2304 / 512 style two-backbone split
    if Dv == 2818:
        split = 2304
    parts = [raw[..., :split], raw[..., split:]] if 0 < split < Dv - 2 else [raw]
    parts = [p / (p.norm(dim=-1, keepdim=True) + 1e-5) for p in parts]
    tef = torch.stack([torch.arange(Lv) / Lv, (torch.arange(Lv) + 1) / Lv], dim=1)[None].expand(B, Lv, 2)
    src_vid = torch.cat(parts + [tef], dim=-1).float()
    src_txt = torch.randn(B, Lt, Dt, generator=g)
    src_txt = src_txt / (src_txt.norm(dim=-1, keepdim=True) + 1e-5)
    vmask = torch.ones(B, Lv)
    tmask = torch.ones(B, Lt)
    if ragged:
        lens_v = torch.randint(max(2, Lv // 5), Lv + 1, (B,), generator=g)
        lens_t = torch.randint(max(2, Lt // 5), Lt + 1, (B,), generator=g)
        lens_v[0] = Lv  # the collate pads to the longest sample
        lens_t[-1] = Lt
        vmask = (torch.arange(Lv)[None] < lens_v[:, None]).float()
        tmask = (torch.arange(Lt)[None] < lens_t[:, None]).float()
        # TEF is computed per sample over its own length before padding
        for b in range(B):
            n = int(lens_v[b])
            src_vid[b, :n, -2] = torch.arange(n) / n
            src_vid[b, :n, -1] = (torch.arange(n) + 1) / n
        src_vid = src_vid * vmask[..., None]
        src_txt = src_txt * tmask[..., None]
    return dict(src_txt=src_txt.float().contiguous(), src_txt_mask=tmask, src_vid=src_vid.float().contiguous(),
                src_vid_mask=vmask)


def make_targets(inputs, seed=2, clip_len=2.0):
    """Dense per-clip targets as DatasetMR/DatasetVLP build them: one random ground-truth window per sample."""
    vmask = inputs["src_vid_mask"]
    B, Lv = vmask.shape
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = vmask.sum(1).long()
    timestamp = torch.zeros(B, Lv, 2)
    window = torch.zeros(B, Lv)
    span_nn = torch.zeros(B, Lv, 2)
    sal = torch.zeros(B, Lv)
    pos = torch.zeros(B, 1, dtype=torch.long)
    neg = torch.zeros(B, 1, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        centers = (torch.arange(n) + 0.5) / n  # ((l + clip_len/2) / L) with duration normalised to 1 (dataset.py:501)
        timestamp[b, :n, 0] = centers
        timestamp[b, :n, 1] = centers
        a = int(torch.randint(0, n, (1,), generator=g))
        e = int(torch.randint(a, n, (1,), generator=g))
        st, ed = a / n, (e + 1) / n
        inside = (centers >= st) & (centers <= ed)
        window[b, :n] = inside.float()
        span_nn[b, :n, 0] = st
        span_nn[b, :n, 1] = ed
        sal[b, :n] = inside.float() * (0.5 + 0.5 * torch.rand(n, generator=g))
        fg = inside.nonzero().flatten()
        pos[b, 0] = fg[int(torch.randint(0, len(fg), (1,), generator=g))]
        bg = (~inside).nonzero().flatten()
        neg[b, 0] = bg[0] if len(bg) else 0
    return dict(timestamp=timestamp, timestamp_mask=vmask.clone(), timestamp_window=window, span_labels_nn=span_nn,
                saliency_scores=sal, saliency_pos_labels=pos, saliency_neg_labels=neg)


def flops_forward(cfg, batch=None, l_vid=None, l_txt=None):
    """Algorithmic forward FLOPs (SURVEY.md 8d formulas), returns (total, encoder_only)."""
    d, ff, N = cfg["hidden_dim"], cfg["dim_feedforward"], cfg["enc_layers"]
    B = batch or cfg["batch"]
    Lv = l_vid or cfg["l_vid"]
    Lt = l_txt or cfg["l_txt"]
    L = Lv + Lt
    enc = N * (8 * L * d * d + 4 * L * d * ff + 4 * L * L * d)
    proj = 2 * Lv * (cfg["v_feat_dim"] * d + d * d) + 2 * Lt * (cfg["t_feat_dim"] * d + d * d)
    heads = 8 * Lv * 3 * d * d + 18 * Lv * d
    return B * (enc + proj + heads), B * enc
