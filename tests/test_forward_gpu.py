"""GPU parity of Model.forward (through the plugin -> C-ABI -> CUDA kernels) against golden vectors and the oracle."""
import pytest
import torch

from tests.helpers import GOLDEN_CASES, golden_out, load_golden, subsample
from univtg_b200 import build_model, synth

pytestmark = pytest.mark.gpu


def _model(cfg, sd, **over):
    model, _ = build_model(synth.reference_args(cfg, device="cuda:0", **over))
    model.load_state_dict(sd, strict=True)
    return model.to("cuda:0").eval()


def _run(model, inp):
    with torch.no_grad():
        out = model(**{k: v.cuda() for k, v in inp.items()})
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_reference_golden(name):
    """North-star tolerance: rtol 1e-3 / atol 1e-4 on saliency, spans, logits vs the fp32 reference; argmax bit-exact."""
    cfg, sd, inp, tgt, z = load_golden(name)
    out = _run(_model(cfg, sd), inp)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(out[k].cpu(), golden_out(z, k), rtol=1e-3, atol=1e-4, msg=lambda m: f"{name}/{k}: {m}")
    # pre-encoder projections feed the saliency losses; fp16 operands over K=2818 leave ~1e-3 absolute error
    torch.testing.assert_close(subsample("vid_mem_proj", out["vid_mem_proj"].cpu(), z), golden_out(z, "vid_mem_proj"), rtol=3e-3,
                               atol=3e-3)
    torch.testing.assert_close(out["txt_mem_proj"].cpu(), golden_out(z, "txt_mem_proj"), rtol=3e-3, atol=3e-3)
    ref_logits = golden_out(z, "pred_logits").squeeze(-1)
    top2 = ref_logits.topk(2, dim=1).values
    decisive = (top2[:, 0] - top2[:, 1]) > 2e-4  # moment index must be bit-exact wherever the reference's margin is real
    got = out["pred_logits"].cpu().squeeze(-1).argmax(1)
    assert torch.equal(got[decisive], ref_logits.argmax(1)[decisive])
    sal = golden_out(z, "saliency_scores")
    t2 = sal.topk(2, dim=1).values
    dec = (t2[:, 0] - t2[:, 1]) > 5e-4
    assert torch.equal(out["saliency_scores"].cpu().argmax(1)[dec], sal.argmax(1)[dec])
    assert out["src_vid_mask"] is not None and out["pred_spans"].shape[-1] == 2
    assert (out["pred_spans"][..., 0] <= 0).all() and (out["pred_spans"][..., 1] >= 0).all()


@pytest.mark.parametrize("name", ["tiny_ragged", "cfg1_demo", "cfg2_b4_ragged"])
def test_forward_matches_fp16_emulating_oracle(name):
    """With the oracle applying the same fp16 operand rounding, only fp32 accumulation order differs."""
    from oracle import univtg_oracle as O

    cfg, sd, inp, tgt, z = load_golden(name)
    out = _run(_model(cfg, sd), inp)
    emu = O.forward(sd, cfg, **inp, opq=O.round_fp16)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(out[k].double().cpu(), emu[k], rtol=2e-4, atol=5e-5, msg=lambda m: f"{name}/{k}: {m}")
    # projector outputs: K = 2818 products accumulated in fp32 (tensor core) vs fp64 (oracle)
    for k in ("vid_mem_proj", "txt_mem_proj"):
        torch.testing.assert_close(out[k].double().cpu(), emu[k], rtol=5e-4, atol=5e-4, msg=lambda m: f"{name}/{k}: {m}")


def test_padded_clips_get_reference_saliency_offset():
    """log(mask + 1e-45) must survive (no flush-to-zero): padded clips carry cos - 103.2789 (model/univtg.py:147)."""
    cfg, sd, inp, tgt, z = load_golden("tiny_ragged")
    out = _run(_model(cfg, sd), inp)
    pad = inp["src_vid_mask"] == 0
    assert pad.any()
    s = out["saliency_scores"].cpu()
    assert torch.isfinite(s).all()
    assert (s[pad] < -102.0).all() and (s[pad] > -104.5).all()
    assert (s[~pad].abs() <= 1.0 + 1e-5).all()


def test_bf16_operand_mode_runs_and_is_coarser():
    cfg, sd, inp, tgt, z = load_golden("cfg2_b4_ragged")
    out = _run(_model(cfg, sd, operand_format="bf16"), inp)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(out[k].cpu(), golden_out(z, k), rtol=1e-2, atol=3e-3)


def test_batch_invariance_and_determinism_full_size():
    """Size-independent properties at BASELINE cfg2 size: a sample's outputs do not depend on its batch neighbours, and the
    forward is bit-reproducible run to run (no atomics on the forward path)."""
    cfg = synth.CONFIGS["cfg2"]
    sd = synth.make_state_dict(cfg, seed=50)
    model = _model(cfg, sd)
    inp = synth.make_inputs(cfg, seed=51, ragged=True)
    a = _run(model, inp)
    b = _run(model, inp)
    for k in ("pred_logits", "pred_spans", "saliency_scores", "vid_mem_proj"):
        assert torch.equal(a[k], b[k]), k
    one = {k: v[5:6].contiguous() for k, v in inp.items()}
    c = _run(model, one)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(c[k][0], a[k][5], rtol=1e-5, atol=2e-6)


def test_padding_invariance():
    """Right-padding a batch with extra masked clips / tokens leaves the valid outputs unchanged (key padding mask)."""
    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, seed=11)
    model = _model(cfg, sd)
    inp = synth.make_inputs(cfg, seed=12, ragged=False, batch=2)
    a = _run(model, inp)
    B, Lt = inp["src_txt_mask"].shape
    padded = dict(inp)
    padded["src_txt"] = torch.cat([inp["src_txt"], torch.zeros(B, 5, cfg["t_feat_dim"])], 1)
    padded["src_txt_mask"] = torch.cat([inp["src_txt_mask"], torch.zeros(B, 5)], 1)
    b = _run(model, padded)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(a[k], b[k], rtol=1e-4, atol=1e-5)


def test_long_video_config5_shape_runs():
    """cfg5 (L = 1277, 6 layers): multi-tile online-softmax attention path; checked against the fp16-emulating oracle on a
    1-sample slice (the oracle needs a few seconds for it)."""
    from oracle import univtg_oracle as O

    cfg = synth.CONFIGS["cfg5"]
    sd = synth.make_state_dict(cfg, seed=60)
    inp = synth.make_inputs(cfg, seed=61, ragged=False, batch=1)
    inp["src_vid_mask"][0, 1100:] = 0  # some padded clips
    inp["src_vid"][0, 1100:] = 0
    out = _run(_model(cfg, sd), inp)
    emu = O.forward(sd, cfg, **inp, opq=O.round_fp16, dtype=torch.float32)
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        torch.testing.assert_close(out[k].double().cpu(), emu[k].double(), rtol=3e-4, atol=1e-4, msg=lambda m: f"cfg5/{k}: {m}")


def test_state_dict_roundtrip_and_repack_on_update():
    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, seed=1)
    model = _model(cfg, sd)
    assert list(model.state_dict().keys()) == list(synth.state_dict_shapes(cfg).keys())
    inp = synth.make_inputs(cfg, seed=2)
    a = _run(model, inp)
    with torch.no_grad():
        model.class_embed.layers[2].bias.add_(1.0)  # in-place update bumps the version counter -> repack
    b = _run(model, inp)
    assert (b["pred_logits"] > a["pred_logits"]).all()
    assert torch.equal(a["pred_spans"], b["pred_spans"])


def test_missing_gpu_tensor_raises():
    cfg = synth.CONFIGS["tiny"]
    model = _model(cfg, synth.make_state_dict(cfg, seed=1))
    inp = synth.make_inputs(cfg, seed=2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            model(**inp)  # CPU tensors: there is no CPU path


def test_cuda_graph_replay_matches_eager_launches():
    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, seed=21)
    model = _model(cfg, sd)
    inp = synth.make_inputs(cfg, seed=22, ragged=True, batch=4)
    ref = _run(model, inp)
    model.use_cuda_graphs = True
    a = _run(model, inp)
    inp2 = synth.make_inputs(cfg, seed=23, ragged=True, batch=4)
    b = _run(model, inp2)
    model.use_cuda_graphs = False
    ref2 = _run(model, inp2)
    for k in ("pred_logits", "pred_spans", "saliency_scores", "vid_mem_proj", "txt_mem_proj"):
        assert torch.equal(a[k], ref[k]), k
        assert torch.equal(b[k], ref2[k]), k
