#!/usr/bin/env python
"""Bitwise run-to-run determinism of the CUDA path: sha1 of o_hat / z_hat over repeated calls (and, run twice, over
processes), next to the CPU oracle's own sha1 (the oracle is torch-CPU code: its bits may depend on the host)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vc_oracle as O
from openvoice_b200.api import NativeSynthesizer
from openvoice_b200.utils import HParams

sd = O.synthetic_state_dict(1234)
m = NativeSynthesizer(HParams(**O.DEFAULT_HPARAMS), "cuda:0")
m.load_state_dict(sd)
spec, lengths, gs, gt, noise = O.synthetic_inputs(2, 48, 3, lengths=[48, 31])
h = lambda t: hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:12]
for mode in ("fp32", "f16x3"):
    m.native.set_precision(mode)
    seen = set()
    for i in range(5):
        o, _, (z, zp, zh) = m.voice_conversion(spec.cuda(), lengths.cuda(), gs.cuda(), gt.cuda(), tau=0.3, noise=noise.cuda(), ragged=True)
        torch.cuda.synchronize()
        seen.add((h(o), h(zh)))
    print(mode, "distinct (o_hat, z_hat) hashes over 5 calls:", len(seen), sorted(seen))
with torch.no_grad():
    ro, _, (rz, rzp, rzh) = O.voice_conversion_ragged(sd, spec, lengths, gs, gt, noise, 0.3)
print("oracle hashes:", h(ro), h(rzh), "threads", torch.get_num_threads())
