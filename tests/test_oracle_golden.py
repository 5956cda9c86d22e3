"""The CPU oracle against vectors produced by the REAL reference (oracle/make_golden.py).

The reference ships no tests or golden vectors; these fixtures are outputs of its own
SynthesizerTrn.voice_conversion / ToneColorConverter.convert / spectrogram_torch /
ReferenceEncoder on the seeded synthetic checkpoint.  Tolerance: the reference's own fp32
noise floor (fp32 vs fp64 twin) is ~8e-7 on latents and ~2e-7 on audio; the oracle must sit
inside 4x of that."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vc_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["vc_b1_t24", "vc_b1_t67", "vc_b2_padded", "vc_b1_t24_v2", "vc_b1_t24_tau0"]


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return d, json.loads(str(d["meta"]))


@pytest.mark.parametrize("name", CASES)
def test_voice_conversion_matches_reference(name, synthetic_sd):
    d, c = load(name)
    spec, lengths, gs, gt, noise = O.synthetic_inputs(c["B"], c["T"], c["seed"], lengths=c["lengths"])
    with torch.no_grad():
        o, mask, (z, zp, zh) = O.voice_conversion(synthetic_sd, spec, lengths, gs, gt, noise, c["tau"], c["zero_g"])
    assert np.abs(o.numpy() - d["o_hat"]).max() < 1e-6
    for got, key in ((z, "z"), (zp, "z_p"), (zh, "z_hat")):
        assert np.abs(got.numpy() - d[key]).max() < 4e-6, key
    assert np.array_equal(mask.numpy(), d["mask"])


def test_ragged_equals_solo(synthetic_sd):
    """convert() is batch 1; the ragged oracle must equal the reference run on each item alone."""
    d, c = load("vc_b1_t24")
    spec, lengths, gs, gt, noise = O.synthetic_inputs(1, 24, c["seed"])
    pad = torch.zeros(1, 513, 40)
    pad[:, :, :24] = spec
    nz = torch.zeros(1, 192, 40)
    nz[:, :, :24] = noise
    with torch.no_grad():
        o, _, (z, zp, zh) = O.voice_conversion_ragged(synthetic_sd, pad, lengths, gs, gt, nz, c["tau"])
    assert np.abs(o.numpy()[:, :, : 24 * 256] - d["o_hat"]).max() < 1e-6
    assert np.abs(o.numpy()[:, :, 24 * 256:]).max() == 0
    assert np.abs(zh.numpy()[:, :, :24] - d["z_hat"]).max() < 4e-6


def test_convert_waveform_matches_reference(synthetic_sd):
    d = np.load(os.path.join(GOLD, "convert_wave.npz"))
    L = int(d["L"])
    rng = np.random.default_rng(1000)
    wav = (0.5 * (2 * rng.random(L, dtype=np.float32) - 1)).astype(np.float32)
    gen = torch.Generator().manual_seed(2000)
    src = 0.1 * torch.randn(1, 256, 1, generator=gen)
    tgt = 0.1 * torch.randn(1, 256, 1, generator=gen)
    spec = O.spectrogram(torch.from_numpy(wav)[None])
    assert np.abs(spec.numpy() - d["spec"]).max() < 1e-5
    noise = torch.randn(1, 192, L // 256, generator=torch.Generator().manual_seed(4000))
    with torch.no_grad():
        a0 = O.convert_waveform(synthetic_sd, torch.from_numpy(wav), src, tgt, None, 0.0)
        a1 = O.convert_waveform(synthetic_sd, torch.from_numpy(wav), src, tgt, noise, 0.3)
    assert a0.shape[0] == 256 * (L // 256)
    assert np.abs(a0.numpy() - d["audio_tau0"]).max() < 1e-6
    assert np.abs(a1.numpy() - d["audio_tau03"]).max() < 1e-6


def test_reference_encoder_matches_reference(synthetic_sd):
    d = np.load(os.path.join(GOLD, "ref_enc.npz"))
    spec = O.synthetic_inputs(2, 140, 7)[0]
    with torch.no_grad():
        g = O.reference_encoder(synthetic_sd, spec.transpose(1, 2))
    assert np.abs(g.numpy() - d["g"]).max() < 1e-6


def test_flow_is_invertible(synthetic_sd):
    """flow(reverse) o flow(forward) with the same g is the identity."""
    spec, lengths, gs, gt, noise = O.synthetic_inputs(1, 20, 11)
    mask = O.sequence_mask(lengths, 20, torch.float32)
    z = noise * mask
    with torch.no_grad():
        back = O.flow(synthetic_sd, O.flow(synthetic_sd, z, mask, gs, False), mask, gs, True)
    assert (back - z).abs().max() < 1e-5


def test_state_dict_schema_counts():
    s = O.state_dict_schema()
    assert len(s) == 486
    n = lambda p: sum(int(np.prod(v)) for k, v in s.items() if k.startswith(p) and not k.endswith("weight_g"))  # noqa: E731
    assert abs(n("dec.") - 14468608) < 20000 and abs(n("enc_q.") - 8823168) < 20000
