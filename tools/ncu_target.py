#!/usr/bin/env python
"""Target process for ncu captures: `calls` convert_waveform calls of batch x secs synthetic clips (default f16x3 mode).

Tensor-core conv launches (tcconv_* and tcpair_* kernels) of ONE call, in order (167): enc WN 32, flow fwd 32, flow rev 32,
then the generator: conv_pre [96], ups0 [97], stage0 [98..115], ups1 [116], stage1 [117..134], ups2 [135], stage2 [136..150],
ups3 [151], stage3 [152..166].  Within stages 0-1: k=3 | 7 | 11, each c1(d1) c2 c1(d3) c2 c1(d5) c2; within stages 2-3 the three
k = 3 pairs are ONE fused launch each (d1, d3, d5), then k = 7 and k = 11 as above.  A call launches 208 kernels in all
(207 + the call-parameter kernel).  See tools/gpu_ncu.sh for the -s / -c arithmetic."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--secs", type=float, default=10.0)
    ap.add_argument("--calls", type=int, default=2)
    ap.add_argument("--precision", default="f16x3")
    args = ap.parse_args()
    import torch
    from oracle import vc_oracle as O
    from openvoice_b200.api import ToneColorConverter

    with tempfile.TemporaryDirectory() as td:
        cfg = os.path.join(td, "config.json")
        json.dump(O.DEFAULT_HPARAMS, open(cfg, "w"))
        conv = ToneColorConverter(cfg, device="cuda:0", enable_watermark=False, precision=args.precision)
    conv.model.load_state_dict(O.synthetic_state_dict(1234))
    B, L = args.batch, int(round(args.secs * 22050))
    wav = (torch.rand(B, L, generator=torch.Generator().manual_seed(0)) - 0.5).cuda()
    wlen = torch.full((B,), L, dtype=torch.int64, device="cuda")
    g = 0.1 * torch.randn(B, 256, generator=torch.Generator().manual_seed(1)).cuda()
    conv.model.native.set_option("graph", 0)      # every call launches its kernels directly: the -s / -c arithmetic stays valid
    for i in range(args.calls):
        conv.model.native.convert_waveform(wav, wlen, g, g, tau=0.3, seed=i)
    torch.cuda.synchronize()
    print("launches per call", conv.model.native.last_launch_count)


if __name__ == "__main__":
    main()
