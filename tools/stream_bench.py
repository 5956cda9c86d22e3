#!/usr/bin/env python
"""Small-request serving: N independent 3 s clips, batch 1 each -- sequential convert() vs convert_concurrent()
(one utterance per CUDA stream) vs one convert_batch() call.  python tools/stream_bench.py [--n 32] [--secs 3]"""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import vc_oracle as O
from openvoice_b200.api import ToneColorConverter

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--secs", type=float, default=3.0)
args = ap.parse_args()
with tempfile.TemporaryDirectory() as td:
    cfg = os.path.join(td, "c.json")
    json.dump(O.DEFAULT_HPARAMS, open(cfg, "w"))
    conv = ToneColorConverter(cfg, device="cuda:0", enable_watermark=False)
conv.model.load_state_dict(O.synthetic_state_dict(1234))
rng = np.random.default_rng(0)
L = int(args.secs * 22050)
wavs = [(0.5 * (2 * rng.random(L, dtype=np.float32) - 1)).astype(np.float32) for _ in range(args.n)]
g = 0.1 * torch.randn(1, 256, 1)
audio_s = args.n * (L // 256) * 256 / 22050


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


res = {"n": args.n, "secs": args.secs}
res["sequential_convert_ms"] = 1e3 * timed(lambda: [conv.convert(w, g, g, tau=0.3) for w in wavs])
for s in (2, 4, 8):
    res[f"concurrent_{s}_streams_ms"] = 1e3 * timed(lambda: conv.convert_concurrent(wavs, g, g, tau=0.3, streams=s))
res["one_batch_ms"] = 1e3 * timed(lambda: conv.convert_batch(wavs, g, g, tau=0.3, max_batch=args.n))
for k in list(res):
    if k.endswith("_ms"):
        res[k.replace("_ms", "_audio_s_per_s")] = audio_s / (res[k] * 1e-3)
print(json.dumps(res))
