"""Pin the oracle against the LIVE reference (only where /root/reference exists, i.e. the build container)."""
import sys

import pytest
import torch

from tests.conftest import REFERENCE, has_reference
from univtg_b200 import synth

pytestmark = pytest.mark.skipif(not has_reference(), reason="/root/reference not present on this box")


def _ref_model(cfg, sd, **over):
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from model.univtg import build_model  # the unmodified reference

    model, crit = build_model(synth.reference_args(cfg, **over))
    model.load_state_dict(sd, strict=True)
    return model, crit


@pytest.mark.parametrize("cfg_name,ragged,batch", [("tiny", True, None), ("tiny", False, 5), ("cfg1", True, 3)])
def test_forward_and_losses(cfg_name, ragged, batch):
    from oracle import univtg_oracle as O

    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg, seed=123)
    model, crit = _ref_model(cfg, sd)
    model.eval()
    inp = synth.make_inputs(cfg, seed=7, ragged=ragged, batch=batch)
    tgt = synth.make_targets(inp, seed=8)
    with torch.no_grad():
        ref = model(**inp)
        ref_loss = crit(ref, tgt)
    out = O.forward(sd, cfg, **inp)
    for k in ("pred_logits", "pred_spans", "saliency_scores", "vid_mem_proj", "txt_mem_proj"):
        torch.testing.assert_close(out[k], ref[k].double(), rtol=2e-5, atol=2e-5)
    loss = O.criterion(out, tgt)
    for k, v in ref_loss.items():
        assert abs(float(loss[k]) - float(v)) < 5e-6 * max(1.0, abs(float(v))), k


def test_droppath_scales_match_reference_train_mode():
    """Train mode with droppath: the reference draws floor(keep + U) per sample, per residual branch, in layer order;
    feeding the same draws to the oracle as scales reproduces its output."""
    from oracle import univtg_oracle as O

    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, seed=5)
    model, _ = _ref_model(cfg, sd, droppath=0.3, input_dropout=0.0)
    model.train()
    inp = synth.make_inputs(cfg, seed=9, ragged=True, batch=6)
    B = inp["src_vid"].shape[0]
    torch.manual_seed(77)
    ref = model(**inp)
    torch.manual_seed(77)
    keep = 0.7
    scales = torch.stack([torch.floor(keep + torch.rand((B, 1, 1))).flatten() / keep for _ in range(2 * cfg["enc_layers"])])
    out = O.forward(sd, cfg, **inp, dp_scale=scales)
    torch.testing.assert_close(out["pred_spans"], ref["pred_spans"].double(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(out["pred_logits"], ref["pred_logits"].double(), rtol=2e-5, atol=2e-5)


def test_state_dict_keys_and_shapes_match_reference():
    for name in ("tiny", "cfg1"):
        cfg = synth.CONFIGS[name]
        model, _ = _ref_model(cfg, synth.make_state_dict(cfg))
        ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert list(ref.items()) == list(synth.state_dict_shapes(cfg).items())
