"""Tone-colour (speaker) embedding extractor used by ``extract_se``.

SURVEY.md section 8 row f2 ("next"): runs once per reference speaker, ~0.03 s for 30 s of audio in
the reference, so it is host-side torch plumbing here (cuDNN conv2d + a GRU recurrence) and not
one of the hand-written kernels.  Functional restatement of ReferenceEncoder.forward
(openvoice/models.py:339-359) on the reference's state-dict keys.
"""
import torch
import torch.nn.functional as F


def _weight(sd, prefix):
    if f"{prefix}.weight" in sd:
        return sd[f"{prefix}.weight"]
    v, g = sd[f"{prefix}.weight_v"], sd[f"{prefix}.weight_g"]
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


class ReferenceEncoder:
    """callable(spec_t [N, T, spec_channels]) -> [N, gin]."""

    def __init__(self, spec_channels, gin_channels):
        self.spec_channels = spec_channels
        self.gin_channels = gin_channels
        self.p = None

    def load_state_dict(self, sd, device):
        keys = [k for k in sd if k.startswith("ref_enc.")]
        if not keys:
            self.p = None
            return False
        p = {k: sd[k].detach().to(device=device, dtype=torch.float32) for k in keys}
        self.p = {"convs": [(_weight(p, f"ref_enc.convs.{i}"), p[f"ref_enc.convs.{i}.bias"]) for i in range(6)],
                  "w_ih": p["ref_enc.gru.weight_ih_l0"], "w_hh": p["ref_enc.gru.weight_hh_l0"],
                  "b_ih": p["ref_enc.gru.bias_ih_l0"], "b_hh": p["ref_enc.gru.bias_hh_l0"],
                  "proj_w": p["ref_enc.proj.weight"], "proj_b": p["ref_enc.proj.bias"],
                  "ln_w": p.get("ref_enc.layernorm.weight"), "ln_b": p.get("ref_enc.layernorm.bias")}
        return True

    @torch.no_grad()
    def __call__(self, inputs, mask=None):
        if self.p is None:
            raise RuntimeError("ref_enc.* weights were not in the checkpoint")
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):   # fp32 like the reference on CPU
            return self._forward(inputs)

    def _forward(self, inputs):
        p = self.p
        N = inputs.size(0)
        x = inputs.reshape(N, 1, -1, self.spec_channels)
        if p["ln_w"] is not None:
            x = F.layer_norm(x, (self.spec_channels,), p["ln_w"], p["ln_b"])
        for w, b in p["convs"]:
            x = F.relu(F.conv2d(x, w, b, stride=2, padding=1))
        x = x.transpose(1, 2).contiguous().view(N, x.size(2), -1)
        gi_all = x @ p["w_ih"].T + p["b_ih"]          # input projections of every step at once
        h = torch.zeros(N, 128, dtype=x.dtype, device=x.device)
        for t in range(x.size(1)):
            gi = gi_all[:, t]
            gh = h @ p["w_hh"].T + p["b_hh"]
            r = torch.sigmoid(gi[:, :128] + gh[:, :128])
            u = torch.sigmoid(gi[:, 128:256] + gh[:, 128:256])
            n = torch.tanh(gi[:, 256:] + r * gh[:, 256:])
            h = (1 - u) * n + u * h
        return h @ p["proj_w"].T + p["proj_b"]
