#!/usr/bin/env python
"""Diagnostic: per-parameter gradient error of the CUDA training step vs the fp64 oracle (and vs the fp16-emulating oracle)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
torch.set_num_threads(16)
from tests.helpers import load_golden
from univtg_b200 import build_model, synth
from oracle import univtg_oracle as O

WD = {"loss_b": 10.0, "loss_g": 1.0, "loss_f": 10.0, "loss_s_intra": 0.1, "loss_s_inter": 0.1}
name = sys.argv[1] if len(sys.argv) > 1 else "tiny_ragged"
cfg, sd, inp, tgt, z = load_golden(name)
model, crit = build_model(synth.reference_args(cfg, device="cuda:0", droppath=0.0, input_dropout=0.0))
model.load_state_dict(sd, strict=True)
model.to("cuda:0").train(); crit.to("cuda:0")
out = model(**{k: v.cuda() for k, v in inp.items()})
loss = crit(out, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in tgt.items()})
total = sum(loss[k] * crit.weight_dict[k] for k in loss)
total.backward(); torch.cuda.synchronize()
res = {}
for tag, opq in (("exact", None), ("fp16emu", O.round_fp16)):
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    o = O.forward(leaves, cfg, **inp, opq=opq)
    l = O.criterion(o, tgt)
    O.weighted_total(l, WD).backward()
    res[tag] = (l, {k: v.grad for k, v in leaves.items()})
print("losses", {k: (float(loss[k]), float(res["exact"][0][k]), float(res["fp16emu"][0][k])) for k in loss})
rows = []
for n_, p in model.named_parameters():
    og = res["exact"][1][n_]
    if og is None or float(og.abs().max()) == 0: 
        continue
    g = p.grad.double().cpu()
    eg = res["fp16emu"][1][n_]
    rel = float((g - og).norm() / og.norm()); rel_e = float((g - eg).norm() / eg.norm()); emu_vs_exact = float((eg - og).norm() / og.norm())
    rows.append((rel, rel_e, emu_vs_exact, n_, float(og.norm())))
rows.sort(reverse=True)
for r in rows[:14]:
    print("rel_exact %.4f rel_emu %.4f emu_vs_exact %.4f  %s  |g|=%.3e" % r)
print("median rel_exact %.4f" % sorted(r[0] for r in rows)[len(rows) // 2])
