"""CPU check of the TTS element functions (openvoice_b200/csrc/ovc_tts_ops.h).

The CUDA kernels of the text side are one-thread-per-element wrappers around these functions; here the SAME header
is compiled with g++ (tests/hostcheck/tts_ops_host.cpp) and every function is compared with the oracle, then the whole
text front half is replayed in the order the device code runs it (dense convs done by torch here, by the tensor-core
conv kernel on the GPU) and compared with the vectors of the real reference."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tts_oracle as T
from oracle import vc_oracle as V

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostcheck") / "tts_ops_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "hostcheck", "tts_ops_host.cpp")])
    return C.CDLL(so)


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def f32(t):
    return np.ascontiguousarray(t.detach().numpy().astype(np.float32))


def cl(t):   # [B,C,T] -> channels-last [B,T,C]
    return np.ascontiguousarray(t.detach().transpose(1, 2).numpy().astype(np.float32))


def uncl(a):
    return torch.from_numpy(a).transpose(1, 2)


class Ops:
    """numpy front-ends of the hostcheck entry points (channels-last arrays)."""

    def __init__(self, lib):
        self.lib = lib

    def embed(self, tokens, lens, emb):
        B, Tn = tokens.shape
        out = np.empty((B, Tn, emb.shape[1]), np.float32)
        tk, ln, e = np.ascontiguousarray(tokens.numpy()), np.ascontiguousarray(lens.numpy()), f32(emb)
        self.lib.hc_embed(fp(tk), fp(ln), fp(e), B, Tn, emb.shape[1], fp(out))
        return out

    def ln(self, a, r, res, gamma, beta, pre=0, post=0):
        out = np.empty_like(a)
        g, b = f32(gamma), f32(beta)
        self.lib.hc_layer_norm(fp(a), fp(r) if r is not None else None, fp(res) if res is not None else None, fp(g), fp(b),
                               a.size // a.shape[-1], a.shape[-1], pre, post, fp(out))
        return out

    def attention(self, qkv, lens, rel_k, rel_v, heads, window):
        B, Tn, H3 = qkv.shape
        H = H3 // 3
        out = np.empty((B, Tn, H), np.float32)
        scores = np.zeros((heads, Tn, Tn), np.float32)
        ln, rk, rv = np.ascontiguousarray(lens.numpy()), f32(rel_k), f32(rel_v)
        self.lib.hc_attention(fp(qkv), fp(ln), fp(rk), fp(rv), B, Tn, H, heads, window, fp(scores), fp(out))
        return out

    def dwconv(self, x, lens, w, bias, dil):
        out = np.empty_like(x)
        ln, ww, bb = np.ascontiguousarray(lens.numpy()), f32(w.reshape(-1, 3)), f32(bias)
        self.lib.hc_dwconv(fp(x), fp(ln), fp(ww), fp(bb), x.shape[0], x.shape[1], x.shape[2], dil, fp(out))
        return out

    def dense(self, x, lens, w, bias, relu_in=False):
        B, Tn, Cin = x.shape
        N, _, K = w.shape
        out = np.empty((B, Tn, N), np.float32)
        ln, wt, bb = np.ascontiguousarray(lens.numpy()), f32(w.permute(2, 1, 0)), f32(bias)     # [K][Cin][N]
        self.lib.hc_dense(fp(x), fp(ln), fp(wt), fp(bb), B, Tn, Cin, K, N, 1 if relu_in else 0, fp(out))
        return out

    def convflow_tail(self, h, lens, pw, pb, x1, bound):
        out = np.empty_like(x1)
        ln, w, b = np.ascontiguousarray(lens.numpy()), f32(pw.reshape(pw.shape[0], -1)), f32(pb)
        self.lib.hc_convflow_tail(fp(h), fp(ln), fp(w), fp(b), fp(x1), h.shape[0], h.shape[1], h.shape[2],
                                  C.c_float(bound), fp(out))
        return out

    def durations(self, ls, ld, lens, ratio, length_scale):
        B, Tn = ls.shape
        logw, wc = np.empty((B, Tn), np.float32), np.empty((B, Tn), np.float32)
        cum, yl = np.empty((B, Tn), np.int32), np.empty((B,), np.int64)
        ln = np.ascontiguousarray(lens.numpy())
        self.lib.hc_durations(fp(ls), fp(ld), fp(ln), C.c_float(ratio), C.c_float(length_scale), B, Tn, fp(logw), fp(wc),
                              fp(cum), fp(yl))
        return logw, wc, cum, yl

    def frame_tokens(self, cum, yl):
        B, Tn = cum.shape
        Ty = int(yl.max())
        tok = np.empty((B, Ty), np.int32)
        self.lib.hc_frame_tokens(fp(cum), fp(yl), B, Tn, Ty, fp(tok))
        return tok


@pytest.fixture(scope="module")
def ops(hc):
    return Ops(hc)


def test_layer_norm_modes(ops):
    g = torch.Generator().manual_seed(1)
    a, r, res = (torch.randn(2, 256, 19, generator=g) for _ in range(3))
    gamma, beta = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    ref = T.layer_norm_c(a + r, gamma, beta)
    assert np.abs(ops.ln(cl(a), cl(r), None, gamma, beta) - cl(ref)).max() < 2e-6
    ref = T.layer_norm_c(torch.relu(a), gamma, beta)
    assert np.abs(ops.ln(cl(a), None, None, gamma, beta, pre=1) - cl(ref)).max() < 2e-6
    ref = res + F.gelu(T.layer_norm_c(a, gamma, beta))
    assert np.abs(ops.ln(cl(a), None, cl(res), gamma, beta, post=1) - cl(ref)).max() < 2e-6


def test_rel_attention(ops):
    g = torch.Generator().manual_seed(2)
    B, H, Tn, heads, win = 2, 192, 23, 2, 4
    q, k, v = (torch.randn(B, H, Tn, generator=g) for _ in range(3))
    rk, rv = (torch.randn(1, 2 * win + 1, H // heads, generator=g) * 0.1 for _ in range(2))
    lens = torch.tensor([Tn, 3])           # 3 < window + 1: the embedding-slice branch (attentions.py:349-362)
    mask = V.sequence_mask(lens, Tn, torch.float32)
    ref = T.rel_attention(q, k, v, mask, rk, rv, heads, win) * mask
    qkv = np.ascontiguousarray(np.concatenate([cl(q), cl(k), cl(v)], axis=2))
    got = ops.attention(qkv, lens, rk[0], rv[0], heads, win)
    assert np.abs(got - cl(ref)).max() < 3e-6


def test_dwconv(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 192, 40, generator=g)
    w, b = torch.randn(192, 1, 3, generator=g), torch.randn(192, generator=g)
    lens = torch.tensor([40, 17])
    mask = V.sequence_mask(lens, 40, torch.float32)
    for dil in (1, 3, 9):
        ref = F.conv1d(x * mask, w, b, padding=dil, dilation=dil, groups=192) * mask
        assert np.abs(ops.dwconv(cl(x), lens, w, b, dil) - cl(ref)).max() < 2e-6


def test_dense_same_conv(ops):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 96, 21, generator=g)
    w, b = torch.randn(40, 96, 3, generator=g) / 17, torch.randn(40, generator=g)
    lens = torch.tensor([21, 8])
    mask = V.sequence_mask(lens, 21, torch.float32)
    ref = F.conv1d(torch.relu(x * mask), w, b, padding=1) * mask
    assert np.abs(ops.dense(cl(x), lens, w, b, relu_in=True) - cl(ref)).max() < 3e-6
    ref = F.conv1d(x * mask, w, b, padding=1) * mask
    assert np.abs(ops.dense(cl(x), lens, w, b) - cl(ref)).max() < 3e-6


def test_spline_inverse(hc):
    g = torch.Generator().manual_seed(4)
    n = 5000
    x = 7 * (2 * torch.rand(n, generator=g) - 1)
    x[:4] = torch.tensor([-5.0, 5.0, 0.0, 4.999999])
    p = 3 * torch.randn(n, 29, generator=g)
    scale = math.sqrt(192.0)
    ref = T.rq_spline(x.double(), p[:, :10].double() / scale, p[:, 10:20].double() / scale, p[:, 20:].double(), True, 5.0)
    out = np.empty(n, np.float32)
    xa, pa = f32(x), f32(p)
    hc.hc_spline_inverse(fp(xa), fp(pa), n, C.c_float(scale), C.c_float(5.0), fp(out))
    assert np.abs(out - ref.numpy()).max() < 2e-5
    ref32 = T.rq_spline(x, p[:, :10] / scale, p[:, 10:20] / scale, p[:, 20:], True, 5.0)
    assert np.abs(out - ref32.numpy()).max() < 2e-5


def test_durations_and_path(ops):
    g = torch.Generator().manual_seed(5)
    B, Tn = 3, 31
    ls, ld = torch.randn(B, Tn, generator=g), 0.5 + 0.3 * torch.randn(B, Tn, generator=g)
    lens = torch.tensor([31, 12, 1])
    mask = V.sequence_mask(lens, Tn, torch.float32)
    logw = ls[:, None] * 0.2 + ld[:, None] * 0.8
    w_ceil = torch.ceil(torch.exp(logw) * mask * 1.1)
    yl_ref = torch.clamp_min(w_ceil.sum([1, 2]), 1).long()
    _, wc, cum, yl = ops.durations(f32(ls), f32(ld), lens, 0.2, 1.1)
    assert np.array_equal(wc, w_ceil[:, 0].numpy()) and np.array_equal(yl, yl_ref.numpy())
    tok = ops.frame_tokens(cum, yl)
    ref_tok = T.generate_path(w_ceil, yl_ref).numpy()
    for b in range(B):
        assert np.array_equal(tok[b, :yl[b]], ref_tok[b, :yl[b]])


def conv_cl(x_cl, w, b, pad=0, relu_in=False, lens=None):
    """Dense conv on channels-last rows, with the conv kernel's conventions: rows >= len read as zero, rows >= len
    not produced (left zero).  On the GPU this is launch_tc (3xTF32 tensor cores)."""
    x = uncl(x_cl)
    if lens is not None:
        x = x * V.sequence_mask(lens, x.shape[2], x.dtype)
    if relu_in:
        x = torch.relu(x)
    y = F.conv1d(x, w, b, padding=pad)
    if lens is not None:
        y = y * V.sequence_mask(lens, x.shape[2], x.dtype)
    return cl(y)


def host_front(ops, sd, tokens, lens, sid, noise_w, noise_scale_w, length_scale, sdp_ratio):
    """The device schedule of ovc_tts_encode, replayed on the CPU."""
    hp, tts = V.DEFAULT_HPARAMS["model"], T.TTS_HPARAMS
    H, heads, win = hp["hidden_channels"], hp["n_heads"], tts["window_size"]
    x = ops.embed(tokens, lens, sd["enc_p.emb.weight"])
    for i in range(hp["n_layers"]):
        a = f"enc_p.encoder.attn_layers.{i}"
        wqkv = torch.cat([sd[f"{a}.conv_{n}.weight"] for n in "qkv"], 0)
        bqkv = torch.cat([sd[f"{a}.conv_{n}.bias"] for n in "qkv"], 0)
        qkv = conv_cl(x, wqkv, bqkv, lens=lens)
        att = ops.attention(qkv, lens, sd[f"{a}.emb_rel_k"][0], sd[f"{a}.emb_rel_v"][0], heads, win)
        y = conv_cl(att, sd[f"{a}.conv_o.weight"], sd[f"{a}.conv_o.bias"], lens=lens)
        e = "enc_p.encoder"
        x = ops.ln(x, y, None, sd[f"{e}.norm_layers_1.{i}.gamma"], sd[f"{e}.norm_layers_1.{i}.beta"])
        f = f"{e}.ffn_layers.{i}"
        h1 = ops.dense(x, lens, sd[f"{f}.conv_1.weight"], sd[f"{f}.conv_1.bias"])
        y = ops.dense(h1, lens, sd[f"{f}.conv_2.weight"], sd[f"{f}.conv_2.bias"], relu_in=True)
        x = ops.ln(x, y, None, sd[f"{e}.norm_layers_2.{i}.gamma"], sd[f"{e}.norm_layers_2.{i}.beta"])
    stats = conv_cl(x, sd["enc_p.proj.weight"], sd["enc_p.proj.bias"], lens=lens)
    g = sd["emb_g.weight"][sid]                                                  # [B,gin]
    # duration predictor
    cv = g @ sd["dp.cond.weight"][:, :, 0].t() + sd["dp.cond.bias"]
    d = x + cv[:, None, :].numpy()
    d = ops.dense(np.ascontiguousarray(d, dtype=np.float32), lens, sd["dp.conv_1.weight"], sd["dp.conv_1.bias"])
    d = ops.ln(d, None, None, sd["dp.norm_1.gamma"], sd["dp.norm_1.beta"], pre=1)
    d = ops.dense(d, lens, sd["dp.conv_2.weight"], sd["dp.conv_2.bias"])
    d = ops.ln(d, None, None, sd["dp.norm_2.gamma"], sd["dp.norm_2.beta"], pre=1)
    logw_d = (torch.from_numpy(d) @ sd["dp.proj.weight"][0, :, 0] + sd["dp.proj.bias"]).numpy()

    def dds(p, h):
        for i in range(3):
            y = ops.dwconv(h, lens, sd[f"{p}.convs_sep.{i}.weight"], sd[f"{p}.convs_sep.{i}.bias"], 3 ** i)
            y = ops.ln(y, None, None, sd[f"{p}.norms_1.{i}.gamma"], sd[f"{p}.norms_1.{i}.beta"], post=1)
            y = conv_cl(y, sd[f"{p}.convs_1x1.{i}.weight"], sd[f"{p}.convs_1x1.{i}.bias"], lens=lens)
            h = ops.ln(y, None, h, sd[f"{p}.norms_2.{i}.gamma"], sd[f"{p}.norms_2.{i}.beta"], post=1)
        return h

    # stochastic duration predictor, reverse
    s = conv_cl(x, sd["sdp.pre.weight"], sd["sdp.pre.bias"], lens=lens)
    cv = g @ sd["sdp.cond.weight"][:, :, 0].t() + sd["sdp.cond.bias"]
    s = dds("sdp.convs", s + cv[:, None, :].numpy())
    xc = conv_cl(s, sd["sdp.proj.weight"], sd["sdp.proj.bias"], lens=lens)
    za, zb = f32(noise_w[:, 0] * noise_scale_w), f32(noise_w[:, 1] * noise_scale_w)
    for j in (3, 2, 1):
        za, zb = zb, za                                                          # Flip
        p = f"sdp.flows.{2 * j + 1}"
        h = za[:, :, None] * f32(sd[f"{p}.pre.weight"][:, 0, 0]) + f32(sd[f"{p}.pre.bias"]) + xc
        h = dds(f"{p}.convs", h)
        zb = ops.convflow_tail(h, lens, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"], zb, 5.0)
    za, zb = zb, za
    m, lg = sd["sdp.flows.0.m"][:, 0], sd["sdp.flows.0.logs"][:, 0]
    logw_s = (za - float(m[0])) * math.exp(-float(lg[0]))
    logw, wc, cum, yl = ops.durations(f32(torch.from_numpy(logw_s)), f32(torch.from_numpy(logw_d)), lens, sdp_ratio,
                                      length_scale)
    return dict(x=x, stats=stats, logw_sdp=logw_s, logw_dp=logw_d, w_ceil=wc, cum=cum, y_lengths=yl)


@pytest.mark.parametrize("name", ["tts_b1_t37", "tts_b2_padded", "tts_b1_t121_tails"])
def test_device_schedule_on_host_matches_reference(name, ops):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    c = json.loads(str(d["meta"]))
    sd = T.synthetic_tts_state_dict()
    tokens, lens, sid, noise_w = T.synthetic_tts_inputs(c["B"], c["T"], c["seed"], c["lengths"])
    with torch.no_grad():
        r = host_front(ops, sd, tokens, lens, sid, noise_w, c["noise_scale_w"], c["length_scale"], c["sdp_ratio"])
    mask = V.sequence_mask(lens, c["T"], torch.float32)[:, 0].numpy()
    m3 = mask[:, :, None]      # rows past the length are don't-care on the device (never read), zero in the reference
    assert np.abs(r["x"] * m3 - cl(torch.from_numpy(d["x"]))).max() < 2e-5
    assert np.abs(r["stats"][..., :192] - cl(torch.from_numpy(d["m_p"]))).max() < 2e-5
    assert np.abs(r["stats"][..., 192:] - cl(torch.from_numpy(d["logs_p"]))).max() < 2e-5
    assert np.abs((r["logw_dp"] - d["logw_dp"][:, 0]) * mask).max() < 2e-5
    assert np.abs((r["logw_sdp"] - d["logw_sdp"][:, 0]) * mask).max() < 1e-4
    assert np.array_equal(r["w_ceil"], d["w_ceil"])
    assert np.array_equal(r["y_lengths"], d["y_lengths"])
