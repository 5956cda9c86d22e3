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
