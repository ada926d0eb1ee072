"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from univtg_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OUT_KEYS = ("pred_logits", "pred_spans", "saliency_scores", "vid_mem_proj", "txt_mem_proj")
GOLDEN_CASES = ("tiny_ragged", "tiny_full", "cfg1_demo", "cfg2_b4_ragged", "cfg2_full", "cfg4_b4_ragged")


def load_golden(name):
    """Returns (cfg, state_dict, inputs, targets, golden npz dict) for a fixture written by tests/golden/make_golden.py."""
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    cfg = synth.CONFIGS[str(z["meta_cfg"])]
    seed = int(z["meta_seed"])
    sd = synth.make_state_dict(cfg, seed=seed, head_gain=float(z["meta_head_gain"]))
    if int(z["meta_demo"]):
        inp = {k: torch.from_numpy(z["in_" + k]) for k in ("src_txt", "src_txt_mask", "src_vid", "src_vid_mask")}
    else:
        inp = synth.make_inputs(cfg, seed=seed + 1, ragged=bool(int(z["meta_ragged"])), batch=int(z["meta_batch"]))
    tgt = synth.make_targets(inp, seed=seed + 2)
    return cfg, sd, inp, tgt, z


def golden_out(z, key):
    return torch.from_numpy(z["out_" + key])


def subsample(key, t, z):
    if key == "vid_mem_proj":
        return t[:, ::int(z["meta_vid_stride"])]
    return t
