"""16-bit feature inputs (packed shards) through the CUDA path: the first LayerNorm reads fp16 features;
parity against the oracle fed the same fp16-rounded inputs, inference and training."""
import numpy as np
import pytest
import torch

from univtg_b200 import build_model, synth
from univtg_b200 import data as D

pytestmark = pytest.mark.gpu
WD = {"loss_b": 10.0, "loss_g": 1.0, "loss_f": 10.0, "loss_s_intra": 0.1, "loss_s_inter": 0.1}


@pytest.mark.parametrize("direct", [True, False])
def test_shard_loader_feeds_fp16_features_to_the_model(tmp_path, direct):
    from oracle import univtg_oracle as O

    cfg = dict(synth.CONFIGS["tiny"], nheads=2)
    sd = synth.make_state_dict(cfg, seed=3)
    # a shard whose matrices are synthetic "already prepared" features of the tiny config's widths
    rng = np.random.default_rng(0)
    vids = [D.l2_normalize(rng.standard_normal((int(n), cfg["v_feat_dim"])).astype(np.float32)) for n in rng.integers(9, 30, 6)]
    qs = [D.l2_normalize(rng.standard_normal((int(n), cfg["t_feat_dim"])).astype(np.float32)) for n in rng.integers(3, 10, 11)]
    samples = [(int(rng.integers(0, 6)), q) for q in range(11)]
    path = str(tmp_path / "t.uvshard")
    D.write_shard(path, vids, qs, samples)
    model, crit = build_model(synth.reference_args(cfg, device="cuda:0", droppath=0.0, input_dropout=0.0))
    model.load_state_dict(sd, strict=True)
    model.to("cuda:0")
    crit.to("cuda:0")
    # direct: the copy engines read the page-locked shard mapping; else: native gather into pinned staging + one H2D per tensor
    loader = D.ShardLoader(path, batch_size=4, device="cuda:0", slots=3, workers=2, direct=direct)
    assert loader.direct in (direct, False)
    n = 0
    for batch, idx in loader:
        assert batch["src_vid"].dtype == torch.float16 and batch["src_vid"].is_cuda
        host = {k: v.detach().cpu() for k, v in batch.items()}
        oin = {"src_vid": host["src_vid"].float(), "src_txt": host["src_txt"].float(), "src_vid_mask": host["src_vid_mask"],
               "src_txt_mask": host["src_txt_mask"]}
        # inference
        model.eval()
        with torch.no_grad():
            out = model(**batch)
        emu = O.forward(sd, cfg, **oin, opq=O.round_fp16)
        for k in ("pred_logits", "pred_spans", "saliency_scores"):
            torch.testing.assert_close(out[k].double().cpu(), emu[k], rtol=2e-4, atol=5e-5, msg=lambda m: f"{k}: {m}")
        # the fp32 route on the same (fp16-representable) values gives bit-identical results: only the load instruction differs
        with torch.no_grad():
            out32 = model(**{k: (v.float() if v.dtype == torch.float16 else v) for k, v in batch.items()})
        for k in ("pred_logits", "pred_spans", "saliency_scores", "vid_mem_proj"):
            assert torch.equal(out[k], out32[k]), k
        # training step: the first projector LayerNorm's backward reads the 16-bit features too
        model.train()
        for p in model.parameters():
            p.grad = None
        tgt = synth.make_targets(oin, seed=40 + n)
        o = model(**batch)
        ld = crit(o, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in tgt.items()})
        sum(ld[k] * crit.weight_dict[k] for k in ld).backward()
        leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
        ol = O.criterion(O.forward(leaves, cfg, **oin), tgt)
        O.weighted_total(ol, WD).backward()
        named = dict(model.named_parameters())
        for n_ in ("input_vid_proj.0.LayerNorm.weight", "input_vid_proj.0.LayerNorm.bias", "input_txt_proj.0.LayerNorm.weight",
                   "input_vid_proj.0.net.1.weight", "input_txt_proj.0.net.1.weight"):
            a, b = named[n_].grad.double().cpu(), leaves[n_].grad
            assert float((a - b).norm() / b.norm()) < 6e-2, (n_, float((a - b).norm() / b.norm()))
        n += 1
    assert n == 3
