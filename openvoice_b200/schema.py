"""State-dict key schema of the converter checkpoint's hot-path tensors:
what ``load_state_dict`` reports as missing / unexpected, like the reference's
``load_state_dict(strict=False)`` (openvoice/api.py:35-39)."""

ENC_Q_LAYERS = 16   # openvoice/models.py:438-446
FLOW_LAYERS = 4     # openvoice/models.py:448
N_FLOWS = 4


def _wn_keys(prefix, n_layers):
    ks = [f"{prefix}.cond_layer.{s}" for s in ("bias", "weight_g", "weight_v")]
    for i in range(n_layers):
        for part in ("in_layers", "res_skip_layers"):
            ks += [f"{prefix}.{part}.{i}.{s}" for s in ("bias", "weight_g", "weight_v")]
    return ks


def hot_path_keys(hps):
    """Keys the CUDA path needs (enc_q.*, flow.*, dec.*)."""
    m = hps.model if hasattr(hps, "model") else hps["model"]
    get = (lambda k: getattr(m, k)) if not isinstance(m, dict) else (lambda k: m[k])
    keys = ["enc_q.pre.weight", "enc_q.pre.bias", "enc_q.proj.weight", "enc_q.proj.bias"]
    keys += _wn_keys("enc_q.enc", ENC_Q_LAYERS)
    for f in range(N_FLOWS):
        p = f"flow.flows.{2 * f}"
        keys += [f"{p}.pre.weight", f"{p}.pre.bias", f"{p}.post.weight", f"{p}.post.bias"]
        keys += _wn_keys(f"{p}.enc", FLOW_LAYERS)
    keys += ["dec.conv_pre.weight", "dec.conv_pre.bias", "dec.cond.weight", "dec.cond.bias", "dec.conv_post.weight"]
    n_up = len(get("upsample_rates"))
    n_k = len(get("resblock_kernel_sizes"))
    for i in range(n_up):
        keys += [f"dec.ups.{i}.{s}" for s in ("bias", "weight_g", "weight_v")]
    for r in range(n_up * n_k):
        for cv in ("convs1", "convs2"):
            for d in range(3):
                keys += [f"dec.resblocks.{r}.{cv}.{d}.{s}" for s in ("bias", "weight_g", "weight_v")]
    return keys


def ref_enc_keys():
    keys = []
    for i in range(6):
        keys += [f"ref_enc.convs.{i}.{s}" for s in ("bias", "weight_g", "weight_v")]
    keys += [f"ref_enc.gru.{s}" for s in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    keys += ["ref_enc.proj.weight", "ref_enc.proj.bias", "ref_enc.layernorm.weight", "ref_enc.layernorm.bias"]
    return keys


def tts_keys(hps):
    """TTS-only members of a V1 base-speaker checkpoint that ``SynthesizerTrn.infer`` reads
    (openvoice/models.py:451-465; ``sdp.post_*`` and ``sdp.flows.1`` are training-only / skipped in reverse)."""
    m = hps.model if hasattr(hps, "model") else hps["model"]
    get = (lambda k: getattr(m, k)) if not isinstance(m, dict) else (lambda k: m[k])
    keys = ["enc_p.emb.weight", "enc_p.proj.weight", "enc_p.proj.bias", "emb_g.weight"]
    for i in range(int(get("n_layers"))):
        a = f"enc_p.encoder.attn_layers.{i}"
        for n in "qkvo":
            keys += [f"{a}.conv_{n}.weight", f"{a}.conv_{n}.bias"]
        keys += [f"{a}.emb_rel_k", f"{a}.emb_rel_v"]
        for n in ("norm_layers_1", "norm_layers_2"):
            keys += [f"enc_p.encoder.{n}.{i}.gamma", f"enc_p.encoder.{n}.{i}.beta"]
        for n in ("conv_1", "conv_2"):
            keys += [f"enc_p.encoder.ffn_layers.{i}.{n}.weight", f"enc_p.encoder.ffn_layers.{i}.{n}.bias"]
    for n in ("conv_1", "conv_2", "proj", "cond"):
        keys += [f"dp.{n}.weight", f"dp.{n}.bias"]
    for n in ("norm_1", "norm_2"):
        keys += [f"dp.{n}.gamma", f"dp.{n}.beta"]

    def dds(p):
        out = []
        for i in range(3):
            out += [f"{p}.convs_sep.{i}.weight", f"{p}.convs_sep.{i}.bias", f"{p}.convs_1x1.{i}.weight",
                    f"{p}.convs_1x1.{i}.bias"]
            for n in ("norms_1", "norms_2"):
                out += [f"{p}.{n}.{i}.gamma", f"{p}.{n}.{i}.beta"]
        return out

    for n in ("pre", "proj", "cond"):
        keys += [f"sdp.{n}.weight", f"sdp.{n}.bias"]
    keys += dds("sdp.convs") + ["sdp.flows.0.m", "sdp.flows.0.logs"]
    for j in (1, 2, 3):
        p = f"sdp.flows.{2 * j + 1}"
        keys += [f"{p}.pre.weight", f"{p}.pre.bias", f"{p}.proj.weight", f"{p}.proj.bias"] + dds(f"{p}.convs")
    return keys
