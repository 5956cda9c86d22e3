"""CPU oracle for the V1 base-speaker TTS front half.

TEST INFRASTRUCTURE ONLY (same rules as ``vc_oracle``): only ``tests/``, ``smoke()`` and the
bench's CPU legs may import this file; the product never does.

Functional torch-CPU restatement of ``SynthesizerTrn.infer`` (openvoice/models.py:467-490):
``TextEncoder`` (models.py:16-57) with the windowed relative-position transformer
(attentions.py:37-121, 210-465), ``DurationPredictor`` (models.py:60-100),
``StochasticDurationPredictor`` in reverse (models.py:102-180) with ``DDSConv`` / ``ConvFlow``
(modules.py:84-130, 459-516) and the rational-quadratic spline (transforms.py:12-209),
``generate_path`` (commons.py:128-142); flow-reverse and the generator come from ``vc_oracle``.

Pinned by ``oracle/make_golden_tts.py`` on outputs of the real reference (tests/golden/tts_*.npz).
The relative-position terms are written as explicit index sums here (the reference uses
pad/reshape skewing, attentions.py:364-397); the two agree to rounding, which the goldens check.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

try:
    from . import vc_oracle as V
except ImportError:   # run as a script from oracle/
    import vc_oracle as V

StateDict = Dict[str, torch.Tensor]

TTS_HPARAMS = {
    "n_vocab": 40,        # len(hps.symbols); api.py:27
    "n_speakers": 4,      # hps.data.n_speakers; api.py:28
    "window_size": 4,     # attentions.py:46
    "dp_filter": 256,     # models.py:463
    "sdp_flows": 4,       # models.py:462
    "dds_layers": 3,      # models.py:116,121,129
    "num_bins": 10,       # modules.py:466
    "tail_bound": 5.0,    # modules.py:467
}


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def tts_state_dict_schema(hp: Optional[dict] = None, tts: Optional[dict] = None) -> Dict[str, tuple]:
    """Names/shapes of the TTS-only members of a V1 base-speaker checkpoint
    (constructors: models.py:16-45, 60-84, 102-133, 458-465; attentions.py:37-101, 210-260,
    410-433; modules.py:17-24, 84-113, 384-389, 459-482).  ``sdp.post_*`` (training only,
    models.py:118-125) is not listed."""
    hp = hp or V.DEFAULT_HPARAMS
    tts = tts or TTS_HPARAMS
    m = hp["model"]
    H, C, Fc, gin = m["hidden_channels"], m["inter_channels"], m["filter_channels"], m["gin_channels"]
    nh, nl, k = m["n_heads"], m["n_layers"], m["kernel_size"]
    win = tts["window_size"]
    out: Dict[str, tuple] = {}
    out["enc_p.emb.weight"] = (tts["n_vocab"], H)
    for i in range(nl):
        a = f"enc_p.encoder.attn_layers.{i}"
        for n in "qkvo":
            out[f"{a}.conv_{n}.weight"] = (H, H, 1)
            out[f"{a}.conv_{n}.bias"] = (H,)
        out[f"{a}.emb_rel_k"] = (1, 2 * win + 1, H // nh)
        out[f"{a}.emb_rel_v"] = (1, 2 * win + 1, H // nh)
        for n in ("norm_layers_1", "norm_layers_2"):
            out[f"enc_p.encoder.{n}.{i}.gamma"] = (H,)
            out[f"enc_p.encoder.{n}.{i}.beta"] = (H,)
        f = f"enc_p.encoder.ffn_layers.{i}"
        out[f"{f}.conv_1.weight"] = (Fc, H, k)
        out[f"{f}.conv_1.bias"] = (Fc,)
        out[f"{f}.conv_2.weight"] = (H, Fc, k)
        out[f"{f}.conv_2.bias"] = (H,)
    out["enc_p.proj.weight"] = (2 * C, H, 1)
    out["enc_p.proj.bias"] = (2 * C,)

    D = tts["dp_filter"]
    out["dp.conv_1.weight"] = (D, H, 3)
    out["dp.conv_1.bias"] = (D,)
    out["dp.conv_2.weight"] = (D, D, 3)
    out["dp.conv_2.bias"] = (D,)
    for n in ("norm_1", "norm_2"):
        out[f"dp.{n}.gamma"] = (D,)
        out[f"dp.{n}.beta"] = (D,)
    out["dp.proj.weight"] = (1, D, 1)
    out["dp.proj.bias"] = (1,)
    out["dp.cond.weight"] = (H, gin, 1)
    out["dp.cond.bias"] = (H,)

    def dds(p: str) -> None:
        for i in range(tts["dds_layers"]):
            out[f"{p}.convs_sep.{i}.weight"] = (H, 1, 3)
            out[f"{p}.convs_sep.{i}.bias"] = (H,)
            out[f"{p}.convs_1x1.{i}.weight"] = (H, H, 1)
            out[f"{p}.convs_1x1.{i}.bias"] = (H,)
            for n in ("norms_1", "norms_2"):
                out[f"{p}.{n}.{i}.gamma"] = (H,)
                out[f"{p}.{n}.{i}.beta"] = (H,)

    for n in ("pre", "proj"):
        out[f"sdp.{n}.weight"] = (H, H, 1)
        out[f"sdp.{n}.bias"] = (H,)
    out["sdp.cond.weight"] = (H, gin, 1)
    out["sdp.cond.bias"] = (H,)
    dds("sdp.convs")
    out["sdp.flows.0.m"] = (2, 1)
    out["sdp.flows.0.logs"] = (2, 1)
    nb = tts["num_bins"]
    for j in range(tts["sdp_flows"]):
        p = f"sdp.flows.{2 * j + 1}"
        out[f"{p}.pre.weight"] = (H, 1, 1)
        out[f"{p}.pre.bias"] = (H,)
        dds(f"{p}.convs")
        out[f"{p}.proj.weight"] = (3 * nb - 1, H, 1)
        out[f"{p}.proj.bias"] = (3 * nb - 1,)
    out["emb_g.weight"] = (tts["n_speakers"], gin)
    return out


def synthetic_tts_state_dict(seed: int = 4321, hp: Optional[dict] = None, tts: Optional[dict] = None) -> StateDict:
    """Converter-side synthetic checkpoint of ``vc_oracle`` (enc_q / flow / dec; ref_enc dropped, V1 base
    speakers have n_speakers > 0, models.py:451-465) plus seeded TTS members.  Gains keep every block
    O(1): the spline parameters span a few units (so bins differ), durations come out at ~1-6 frames."""
    hp = hp or V.DEFAULT_HPARAMS
    sd = {k: v for k, v in V.synthetic_state_dict(1234, hp).items() if not k.startswith("ref_enc.")}
    gen = torch.Generator().manual_seed(seed)
    H = hp["model"]["hidden_channels"]
    for name, shape in sorted(tts_state_dict_schema(hp, tts).items()):
        if name.endswith(".gamma"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif name.endswith(".beta") or name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=gen)
        elif name == "enc_p.emb.weight":
            sd[name] = torch.randn(shape, generator=gen) * H ** -0.5
        elif "emb_rel_" in name:
            sd[name] = torch.randn(shape, generator=gen) * shape[-1] ** -0.5
        elif name == "emb_g.weight":
            sd[name] = 0.1 * torch.randn(shape, generator=gen)
        elif name.endswith(".m") or name.endswith(".logs"):
            sd[name] = 0.3 * torch.randn(shape, generator=gen)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 1.0
            if name.startswith("sdp.flows.") and ".proj." in name:
                gain = 6.0        # h / sqrt(192) (modules.py:497-499) must still spread the bins
            elif name == "enc_p.proj.weight":
                gain = 0.5
            elif name == "dp.proj.weight":
                gain = 0.4
            sd[name] = torch.randn(shape, generator=gen) * (gain / math.sqrt(fan_in))
    sd["dp.proj.bias"] = torch.tensor([0.7])
    return sd


def synthetic_tts_inputs(B: int, T: int, seed: int = 0, lengths=None, tts: Optional[dict] = None):
    """Seeded (tokens [B,T] int64, lengths [B], sid [B], noise_w [B,2,T])."""
    tts = tts or TTS_HPARAMS
    gen = torch.Generator().manual_seed(20_000 + seed)
    tokens = torch.randint(0, tts["n_vocab"], (B, T), generator=gen)
    sid = torch.randint(0, tts["n_speakers"], (B,), generator=gen)
    noise_w = torch.randn(B, 2, T, generator=gen)
    if lengths is None:
        lengths = [T] * B
    lengths = torch.tensor(list(lengths), dtype=torch.int64)
    for b in range(B):
        tokens[b, int(lengths[b]):] = 0
    return tokens, lengths, sid, noise_w


# --------------------------------------------------------------------------------------
# transformer text encoder
# --------------------------------------------------------------------------------------
def layer_norm_c(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the channel axis of [B,C,T] (modules.py:26-29)."""
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), gamma, beta, eps).transpose(1, 2)


def rel_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor,
                  rel_k: torch.Tensor, rel_v: torch.Tensor, n_heads: int, window: int) -> torch.Tensor:
    """Self-attention core with windowed relative-position keys/values (attentions.py:272-324).

    q,k,v [B,C,T]; mask [B,1,T] (1 = valid); rel_* [1,2w+1,dk].  Score(i,j) = q_i.k_j/sqrt(dk)
    + [|j-i|<=w] q_i.Ek[j-i+w]/sqrt(dk); masked pairs -> -1e4 (:296); softmax over j; out_i =
    sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w] (:312-319)."""
    B, C, T = q.shape
    dk = C // n_heads
    qh = q.view(B, n_heads, dk, T).transpose(2, 3) / math.sqrt(dk)
    kh = k.view(B, n_heads, dk, T).transpose(2, 3)
    vh = v.view(B, n_heads, dk, T).transpose(2, 3)
    scores = qh @ kh.transpose(-1, -2)                                  # [B,h,T,T]
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]                                   # j - i
    inside = rel.abs() <= window
    slot = (rel + window).clamp(0, 2 * window)
    qe = qh @ rel_k[0].t()                                              # [B,h,T,2w+1]
    local = torch.gather(qe, 3, slot[None, None].expand(B, n_heads, T, T))
    scores = scores + local * inside
    pair = mask.unsqueeze(2) * mask.unsqueeze(-1)                       # [B,1,T,T] (attentions.py:106)
    scores = scores.masked_fill(pair == 0, -1e4)
    p = torch.softmax(scores, dim=-1)
    out = p @ vh
    pw = torch.zeros(B, n_heads, T, 2 * window + 1, dtype=p.dtype)
    pw.scatter_add_(3, slot[None, None].expand(B, n_heads, T, T), p * inside)
    out = out + pw @ rel_v[0]
    return out.transpose(2, 3).reshape(B, C, T)


def mha(sd: StateDict, p: str, x: torch.Tensor, mask: torch.Tensor, n_heads: int, window: int) -> torch.Tensor:
    """MultiHeadAttention.forward on (x, x) (attentions.py:262-270)."""
    q = F.conv1d(x, sd[f"{p}.conv_q.weight"], sd[f"{p}.conv_q.bias"])
    k = F.conv1d(x, sd[f"{p}.conv_k.weight"], sd[f"{p}.conv_k.bias"])
    v = F.conv1d(x, sd[f"{p}.conv_v.weight"], sd[f"{p}.conv_v.bias"])
    o = rel_attention(q, k, v, mask, sd[f"{p}.emb_rel_k"], sd[f"{p}.emb_rel_v"], n_heads, window)
    return F.conv1d(o, sd[f"{p}.conv_o.weight"], sd[f"{p}.conv_o.bias"])


def ffn(sd: StateDict, p: str, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """FFN.forward, relu activation, 'same' padding (attentions.py:439-465)."""
    w1, w2 = sd[f"{p}.conv_1.weight"], sd[f"{p}.conv_2.weight"]
    k = w1.shape[2]
    pad = ((k - 1) // 2, k // 2)
    h = F.conv1d(F.pad(x * mask, pad), w1, sd[f"{p}.conv_1.bias"])
    h = torch.relu(h)
    h = F.conv1d(F.pad(h * mask, pad), w2, sd[f"{p}.conv_2.bias"])
    return h * mask


def text_encoder(sd: StateDict, tokens: torch.Tensor, lengths: torch.Tensor, hp: Optional[dict] = None,
                 tts: Optional[dict] = None):
    """TextEncoder.forward (models.py:47-57) + Encoder.forward (attentions.py:105-121).
    Returns x [B,H,T], m [B,C,T], logs [B,C,T], mask [B,1,T]."""
    hp = hp or V.DEFAULT_HPARAMS
    tts = tts or TTS_HPARAMS
    m = hp["model"]
    H, C = m["hidden_channels"], m["inter_channels"]
    x = F.embedding(tokens, sd["enc_p.emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, 2)
    mask = V.sequence_mask(lengths, x.shape[2], x.dtype)
    x = x * mask
    for i in range(m["n_layers"]):
        e = "enc_p.encoder"
        y = mha(sd, f"{e}.attn_layers.{i}", x, mask, m["n_heads"], tts["window_size"])
        x = layer_norm_c(x + y, sd[f"{e}.norm_layers_1.{i}.gamma"], sd[f"{e}.norm_layers_1.{i}.beta"])
        y = ffn(sd, f"{e}.ffn_layers.{i}", x, mask)
        x = layer_norm_c(x + y, sd[f"{e}.norm_layers_2.{i}.gamma"], sd[f"{e}.norm_layers_2.{i}.beta"])
    x = x * mask
    stats = F.conv1d(x, sd["enc_p.proj.weight"], sd["enc_p.proj.bias"]) * mask
    return x, stats[:, :C], stats[:, C:], mask


# --------------------------------------------------------------------------------------
# duration predictors
# --------------------------------------------------------------------------------------
def duration_predictor(sd: StateDict, x: torch.Tensor, mask: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """DurationPredictor.forward (models.py:86-100): conv-relu-LN twice, 1x1 proj."""
    x = x + F.conv1d(g, sd["dp.cond.weight"], sd["dp.cond.bias"])
    for n in ("1", "2"):
        x = F.conv1d(x * mask, sd[f"dp.conv_{n}.weight"], sd[f"dp.conv_{n}.bias"], padding=1)
        x = layer_norm_c(torch.relu(x), sd[f"dp.norm_{n}.gamma"], sd[f"dp.norm_{n}.beta"])
    return F.conv1d(x * mask, sd["dp.proj.weight"], sd["dp.proj.bias"]) * mask


def dds_conv(sd: StateDict, p: str, x: torch.Tensor, mask: torch.Tensor, g: Optional[torch.Tensor] = None,
             n_layers: int = 3, k: int = 3) -> torch.Tensor:
    """DDSConv.forward (modules.py:115-130): depthwise dilated conv (dilation k^i), LN, GELU(erf),
    1x1, LN, GELU, residual."""
    if g is not None:
        x = x + g
    C = x.shape[1]
    for i in range(n_layers):
        d = k ** i
        y = F.conv1d(x * mask, sd[f"{p}.convs_sep.{i}.weight"], sd[f"{p}.convs_sep.{i}.bias"],
                     padding=(k * d - d) // 2, dilation=d, groups=C)
        y = F.gelu(layer_norm_c(y, sd[f"{p}.norms_1.{i}.gamma"], sd[f"{p}.norms_1.{i}.beta"]))
        y = F.conv1d(y, sd[f"{p}.convs_1x1.{i}.weight"], sd[f"{p}.convs_1x1.{i}.bias"])
        y = F.gelu(layer_norm_c(y, sd[f"{p}.norms_2.{i}.gamma"], sd[f"{p}.norms_2.{i}.beta"]))
        x = x + y
    return x * mask


MIN_BIN = 1e-3      # transforms.py:7-9 (width, height and derivative floors)


def rq_spline(x: torch.Tensor, uw: torch.Tensor, uh: torch.Tensor, ud: torch.Tensor, inverse: bool,
              tail_bound: float) -> torch.Tensor:
    """Monotone rational-quadratic spline with linear tails (transforms.py:50-209), values only.

    x [...]; uw, uh [..., nb]; ud [..., nb-1].  Outside [-B, B]: identity (:75-76).  Inside: knots from
    softmax widths/heights floored at 1e-3 (:120-140), derivatives 1e-3 + softplus with the two edge
    derivatives pinned so the tails join with slope 1 (:70-73), bin = #edges <= x - 1 with the last edge
    nudged by 1e-6 (:45-47), inverse = root 2c / (-b - sqrt(b^2 - 4ac)) (:161-176)."""
    nb = uw.shape[-1]
    const = math.log(math.exp(1 - MIN_BIN) - 1)
    edge = torch.full_like(ud[..., :1], const)
    d = MIN_BIN + F.softplus(torch.cat([edge, ud, edge], dim=-1))

    def knots(u: torch.Tensor):
        w = MIN_BIN + (1 - MIN_BIN * nb) * torch.softmax(u, dim=-1)
        c = F.pad(torch.cumsum(w, dim=-1), (1, 0))
        c = 2 * tail_bound * c - tail_bound
        c[..., 0] = -tail_bound
        c[..., -1] = tail_bound
        return c, c[..., 1:] - c[..., :-1]

    cw, w = knots(uw)
    ch, h = knots(uh)
    inside = (x >= -tail_bound) & (x <= tail_bound)
    xc = x.clamp(-tail_bound, tail_bound)
    edges = (ch if inverse else cw).clone()
    edges[..., -1] += 1e-6
    b = ((xc[..., None] >= edges).sum(-1) - 1).clamp(0, nb - 1)[..., None]
    pick = lambda t: t.gather(-1, b)[..., 0]  # noqa: E731
    in_cw, in_w, in_ch, in_h = pick(cw), pick(w), pick(ch), pick(h)
    delta = in_h / in_w
    d0, d1 = pick(d), pick(d[..., 1:])
    if inverse:
        y = xc - in_ch
        s = d0 + d1 - 2 * delta
        qa = y * s + in_h * (delta - d0)
        qb = in_h * d0 - y * s
        qc = -delta * y
        root = (2 * qc) / (-qb - torch.sqrt(qb * qb - 4 * qa * qc))
        out = root * in_w + in_cw
    else:
        th = (xc - in_cw) / in_w
        t1 = th * (1 - th)
        out = in_ch + in_h * (delta * th * th + d0 * t1) / (delta + (d0 + d1 - 2 * delta) * t1)
    return torch.where(inside, out, x)


def conv_flow_reverse(sd: StateDict, p: str, z: torch.Tensor, mask: torch.Tensor, g: torch.Tensor,
                      tts: dict) -> torch.Tensor:
    """ConvFlow.forward(reverse=True) on z [B,2,T] (modules.py:484-516)."""
    x0, x1 = z[:, :1], z[:, 1:]
    h = F.conv1d(x0, sd[f"{p}.pre.weight"], sd[f"{p}.pre.bias"])
    h = dds_conv(sd, f"{p}.convs", h, mask, g=g, n_layers=tts["dds_layers"])
    h = F.conv1d(h, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"]) * mask      # [B,29,T]
    nb = tts["num_bins"]
    hp_ = h.transpose(1, 2)                                                   # [B,T,29] (half_channels = 1)
    scale = math.sqrt(sd[f"{p}.pre.weight"].shape[0])
    y1 = rq_spline(x1[:, 0], hp_[..., :nb] / scale, hp_[..., nb:2 * nb] / scale, hp_[..., 2 * nb:],
                   inverse=True, tail_bound=tts["tail_bound"])
    return torch.cat([x0, y1[:, None]], 1) * mask


def sdp_reverse(sd: StateDict, x: torch.Tensor, mask: torch.Tensor, g: torch.Tensor, noise_w: torch.Tensor,
                noise_scale_w: float, tts: Optional[dict] = None) -> torch.Tensor:
    """StochasticDurationPredictor.forward(reverse=True) (models.py:135-143, 170-180); ``noise_w`` [B,2,T]
    replaces torch.randn at :173.  Flow order: Flip, CF3, Flip, CF2, Flip, CF1, Flip, ElementwiseAffine^-1
    (the first ConvFlow is dropped, :172)."""
    tts = tts or TTS_HPARAMS
    x = F.conv1d(x, sd["sdp.pre.weight"], sd["sdp.pre.bias"])
    x = x + F.conv1d(g, sd["sdp.cond.weight"], sd["sdp.cond.bias"])
    x = dds_conv(sd, "sdp.convs", x, mask, n_layers=tts["dds_layers"])
    x = F.conv1d(x, sd["sdp.proj.weight"], sd["sdp.proj.bias"]) * mask
    z = noise_w * noise_scale_w
    for j in range(tts["sdp_flows"] - 1, 0, -1):
        z = torch.flip(z, [1])                                               # modules.py:375-376
        z = conv_flow_reverse(sd, f"sdp.flows.{2 * j + 1}", z, mask, x, tts)
    z = torch.flip(z, [1])
    z = (z - sd["sdp.flows.0.m"]) * torch.exp(-sd["sdp.flows.0.logs"]) * mask  # modules.py:398-399
    return z[:, :1]


def generate_path(w_ceil: torch.Tensor, y_lengths: torch.Tensor) -> torch.Tensor:
    """Frame -> token index [B, Ty] (the argmax of commons.generate_path's 0/1 matrix, commons.py:128-142):
    frame y belongs to the first token whose cumulative duration exceeds y."""
    cum = torch.cumsum(w_ceil[:, 0], -1)                                     # [B,T]
    Ty = int(y_lengths.max())
    y = torch.arange(Ty)[None, :, None].to(cum.dtype)
    return (cum[:, None, :] <= y).sum(-1).clamp(max=w_ceil.shape[-1] - 1)


def tts_infer(sd: StateDict, tokens: torch.Tensor, lengths: torch.Tensor, sid: torch.Tensor,
              noise_w: torch.Tensor, noise: Optional[torch.Tensor] = None, noise_scale: float = 0.667,
              length_scale: float = 1.0, noise_scale_w: float = 0.6, sdp_ratio: float = 0.2,
              hp: Optional[dict] = None, tts: Optional[dict] = None, ragged: bool = False,
              max_len: Optional[int] = None) -> dict:
    """SynthesizerTrn.infer (models.py:467-490).  ``noise`` [B,C,Ty] replaces torch.randn_like at :487
    (None -> zeros).  ``ragged``: decode every utterance on its own length (what a B=1 call gives)."""
    hp = hp or V.DEFAULT_HPARAMS
    tts = tts or TTS_HPARAMS
    x, m_p, logs_p, x_mask = text_encoder(sd, tokens, lengths, hp, tts)
    g = F.embedding(sid, sd["emb_g.weight"]).unsqueeze(-1)
    logw_s = sdp_reverse(sd, x, x_mask, g, noise_w, noise_scale_w, tts)
    logw_d = duration_predictor(sd, x, x_mask, g)
    logw = logw_s * sdp_ratio + logw_d * (1 - sdp_ratio)
    w = torch.exp(logw) * x_mask * length_scale
    w_ceil = torch.ceil(w)
    y_lengths = torch.clamp_min(w_ceil.sum([1, 2]), 1).long()
    Ty = int(y_lengths.max())
    y_mask = V.sequence_mask(y_lengths, Ty, x.dtype)
    tok = generate_path(w_ceil, y_lengths)                                    # [B,Ty]
    gi = tok[:, None, :].expand(-1, m_p.shape[1], -1)
    m_y = torch.gather(m_p, 2, gi) * y_mask
    logs_y = torch.gather(logs_p, 2, gi) * y_mask
    if noise is None:
        noise = torch.zeros_like(m_y)
    z_p = m_y + noise[:, :, :Ty] * torch.exp(logs_y) * noise_scale
    res = {"x": x, "m_p": m_p, "logs_p": logs_p, "logw_sdp": logw_s, "logw_dp": logw_d, "logw": logw,
           "w_ceil": w_ceil, "y_lengths": y_lengths, "tok": tok, "z_p": z_p, "g": g}
    if ragged:
        zs, os_ = torch.zeros_like(z_p), torch.zeros(z_p.shape[0], 1, Ty * hp["data"]["hop_length"])
        for b in range(z_p.shape[0]):
            n = int(y_lengths[b])
            mb = torch.ones(1, 1, n)
            zb = V.flow(sd, z_p[b:b + 1, :, :n], mb, g[b:b + 1], reverse=True)
            zs[b:b + 1, :, :n] = zb
            ob = V.generator(sd, zb, g[b:b + 1], hp)
            os_[b:b + 1, :, :ob.shape[-1]] = ob
        res["z"], res["o"] = zs, os_
    else:
        z = V.flow(sd, z_p, y_mask, g, reverse=True)
        res["z"] = z
        res["o"] = V.generator(sd, (z * y_mask)[:, :, :max_len], g, hp)          # models.py:489
    return res
