"""Tone-colour (speaker) embedding extractor used by ``extract_se``.

``converter.model.ref_enc(spec_t)`` keeps the reference's call shape -- ``spec_t`` is ``[N, T, spec_channels]``
(``y.transpose(1, 2)``, openvoice/api.py:130) and the result ``[N, gin]`` -- but the arithmetic of
ReferenceEncoder.forward (openvoice/models.py:339-359) runs in libovc_b200.so (``ovc_reference_encoder``:
LayerNorm, 6 x Conv2d 3x3 s2 + ReLU, GRU, Linear as CUDA kernels in csrc/ovc_refenc.cuh).
"""
import torch


class ReferenceEncoder:
    def __init__(self, native, spec_channels, gin_channels):
        self.native = native
        self.spec_channels = spec_channels
        self.gin_channels = gin_channels

    @torch.no_grad()
    def __call__(self, inputs, mask=None):
        """inputs [N, T, spec_channels] (any device) -> [N, gin] on the converter's device."""
        dev = torch.device("cuda", self.native.device_index)
        spec = inputs.to(dev, torch.float32).transpose(1, 2).contiguous()   # [N, F, T], the kernels' layout
        return self.native.reference_encoder(spec)
