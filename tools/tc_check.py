#!/usr/bin/env python
"""Parity of the 3xTF32 tensor-core generator path vs the CPU oracle, stage by stage."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import vc_oracle as O
from openvoice_b200.api import NativeSynthesizer
from openvoice_b200.utils import HParams


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / (np.sqrt((b ** 2).mean()) + 1e-30))


sd = O.synthetic_state_dict(1234)
m = NativeSynthesizer(HParams(**O.DEFAULT_HPARAMS), "cuda:0")
m.load_state_dict(sd)
B, T, lens = 2, 70, [70, 41]
spec, lengths, gs, gt, noise = O.synthetic_inputs(B, T, 3, lengths=lens)
taps = {}
with torch.no_grad():
    ro, _, (rz, rzp, rzh) = O.voice_conversion(sd, spec, lengths, gs, gt, noise, 0.3, taps=taps)
for mode in ("fp32", "f16x3", "f16"):
    m.native.set_precision(mode)
    m.native.debug_enable(True)
    o, _, lat = m.voice_conversion(spec.cuda(), lengths.cuda(), gs.cuda(), gt.cuda(), tau=0.3, noise=noise.cuda(), ragged=False)
    torch.cuda.synchronize()
    print(mode, "o_hat", f"{rel(o.cpu().numpy(), ro.numpy()):.3e}", "z_hat", f"{rel(lat[2].cpu().numpy(), rzh.numpy()):.3e}")
    for name in ["dec.pre", "dec.ups0", "dec.stage0", "dec.ups1", "dec.stage1", "dec.ups2", "dec.stage2", "dec.ups3", "dec.stage3"]:
        print("   ", name, f"{rel(m.native.debug_fetch(name), taps[name].numpy()):.3e}")
    m.native.debug_enable(False)
    o2, _, _ = m.voice_conversion(spec.cuda(), lengths.cuda(), gs.cuda(), gt.cuda(), tau=0.3, noise=noise.cuda(), ragged=True)
    with torch.no_grad():
        qo, _, _ = O.voice_conversion_ragged(sd, spec, lengths, gs, gt, noise, 0.3)
    print("    ragged o_hat", f"{rel(o2.cpu().numpy(), qo.numpy()):.3e}")
print("tc_check done")
