"""Stateful streaming front of the tone-colour converter.

``ToneColorConverter.convert`` (openvoice/api.py:141-160) needs the whole utterance.  The path is not causal -- the
posterior encoder, both flow passes and the generator together see +-110 spectrogram frames
(``ToneColorConverter.HALO_FRAMES`` = 128 with margin) -- so a stream can emit frame t once the audio of frame
t + 128 has arrived.  ``StreamingConverter`` keeps exactly that much state between calls:

* the spectrogram frames of the last window's right halo plus the left halo of the next one (computed on the device
  from the audio tail, frames whose STFT support is still incomplete are left for the next push),
* the noise columns drawn for those frames (the draw of ``models.py:220`` is per frame: a frame keeps its noise no
  matter which window it is converted in),
* the audio samples not yet covered by a complete STFT frame.

Every ``push`` converts as many ``window_frames``-frame windows as the new audio completes -- window + halo on both
sides as ONE batch-1 ragged call of the same kernels ``convert`` runs -- and returns their interiors; ``flush`` ends
the stream (right reflect padding of ``spectrogram_torch``, mel_processing.py:62-63) and returns the rest.  The
concatenated output equals ``convert`` on the whole clip with the same noise (``tests/test_gpu_parity.py``:
<= 2e-6 * rms, the fp32 reordering between tile geometries), whatever the chunking of the input.
Algorithmic latency: (128 + window_frames) * 256 samples; compute per emitted frame: (window + 256) / window of the
offline cost.
"""
from typing import Callable, Optional

import numpy as np
import torch


class StreamingConverter:
    def __init__(self, converter, src_se, tgt_se, tau: float = 0.3, window_frames: int = 256,
                 noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None, seed: Optional[int] = None):
        """``converter``: a ToneColorConverter.  ``noise_fn(t0, t1) -> [inter_channels, t1 - t0]`` supplies the noise of
        absolute frames [t0, t1) (tests pass slices of one tensor); default: a seeded device generator."""
        self.conv = converter
        hp = converter.hps
        self.hop = hp.data.hop_length
        self.nfft = hp.data.filter_length
        self.pad = (self.nfft - self.hop) // 2            # reflect padding of spectrogram_torch
        self.C = hp.model.inter_channels
        self.S = hp.data.filter_length // 2 + 1
        self.H = converter.HALO_FRAMES
        self.W = int(window_frames)
        assert self.W >= 1
        self.tau = float(tau)
        self.dev = converter.device
        self.src = converter._stack_se(src_se, 1)
        self.tgt = converter._stack_se(tgt_se, 1)
        if noise_fn is None:
            gen = torch.Generator(device=self.dev)
            gen.manual_seed(int(seed if seed is not None else torch.randint(0, 2 ** 62, (1,)).item()))
            noise_fn = lambda t0, t1: torch.randn(self.C, t1 - t0, device=self.dev, generator=gen)  # noqa: E731
        self.noise_fn = noise_fn
        # ---- state
        self.audio = np.zeros(0, dtype=np.float32)        # samples from absolute index a0 on
        self.a0 = 0
        self.n_in = 0                                     # samples received so far
        self.spec = torch.zeros(1, self.S, 0, device=self.dev)    # frames [f0, f0 + spec.shape[2])
        self.noise = torch.zeros(self.C, 0, device=self.dev)
        self.f0 = 0
        self.emitted = 0                                  # frames whose samples have been returned
        self.closed = False

    # ------------------------------------------------------------------ state size (tests: bounded)
    @property
    def state_frames(self) -> int:
        return int(self.spec.shape[2])

    @property
    def state_samples(self) -> int:
        return int(len(self.audio))

    # ------------------------------------------------------------------ spectrogram frames as audio arrives
    def _extend_spec(self, final: bool):
        """Append every frame whose STFT support [t*hop - pad, t*hop - pad + nfft) is complete (all of them, with the
        right reflect padding, when the stream ends)."""
        hop, pad = self.hop, self.pad
        have = self.f0 + self.spec.shape[2]               # next frame to compute
        if final:
            upto = self.n_in // hop
        else:
            upto = max(0, (self.n_in + pad - self.nfft) // hop + 1)
            upto = min(upto, self.n_in // hop)
        if upto <= have:
            return
        # segment of audio that gives frames [have, upto) away from its own reflect-padded ends: the native STFT pads
        # the segment it is given, so start 2 frames early (unless at the stream start) and require the support of
        # frame upto-1 inside the segment (unless the stream has ended)
        lead = 2 if have >= 2 else have
        s_lo = (have - lead) * hop
        s_hi = self.n_in
        seg = self.audio[s_lo - self.a0: s_hi - self.a0]
        wav = torch.from_numpy(np.ascontiguousarray(seg)).to(self.dev)[None]
        wlen = torch.tensor([wav.shape[1]], dtype=torch.int64, device=self.dev)
        sp, _ = self.conv.model.native.spectrogram(wav.contiguous(), wlen)
        new = sp[:, :, lead: lead + (upto - have)]
        assert new.shape[2] == upto - have, (new.shape, upto, have, lead)
        self.spec = torch.cat([self.spec, new], 2)
        self.noise = torch.cat([self.noise, self.noise_fn(have, upto).to(self.dev, torch.float32).reshape(self.C, -1)], 1)
        # audio before the support of the next frame (and its 2 lead frames) is no longer needed
        keep_from = max(0, (upto - 2) * hop - pad)
        if keep_from > self.a0:
            self.audio = self.audio[keep_from - self.a0:]
            self.a0 = keep_from

    def _convert_window(self, e0: int, e1: int, t_end: Optional[int]) -> np.ndarray:
        """Samples of frames [e0, e1): one ragged batch-1 call over [e0 - H, e1 + H) clipped to the stream."""
        lo = max(0, e0 - self.H)
        hi = e1 + self.H if t_end is None else min(t_end, e1 + self.H)
        sp = self.spec[:, :, lo - self.f0: hi - self.f0].contiguous()
        nz = self.noise[None, :, lo - self.f0: hi - self.f0].contiguous()
        lens = torch.tensor([hi - lo], dtype=torch.int64, device=self.dev)
        o, _, _ = self.conv.model.voice_conversion(sp, lens, self.src, self.tgt, tau=self.tau, noise=nz, ragged=True,
                                                   latents=False)
        out = o[0, 0, (e0 - lo) * self.hop: (e1 - lo) * self.hop].cpu().numpy().copy()
        self.emitted = e1
        # frames before the next window's left halo can go
        drop = max(0, e1 - self.H) - self.f0
        if drop > 0:
            self.spec = self.spec[:, :, drop:]
            self.noise = self.noise[:, drop:]
            self.f0 += drop
        return out

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def push(self, samples) -> np.ndarray:
        """Feed float32 samples at the model's sampling rate; returns the converted samples that became final."""
        assert not self.closed, "the stream has been flushed"
        x = np.asarray(samples, dtype=np.float32).reshape(-1)
        self.audio = np.concatenate([self.audio, x])
        self.n_in += len(x)
        self._extend_spec(final=False)
        outs = []
        have = self.f0 + self.spec.shape[2]
        while self.emitted + self.W + self.H <= have:
            outs.append(self._convert_window(self.emitted, self.emitted + self.W, None))
        return np.concatenate(outs) if outs else np.zeros(0, dtype=np.float32)

    @torch.no_grad()
    def flush(self) -> np.ndarray:
        """End of the stream: converts what is left (the last frames use the right reflect padding, exactly like the
        whole-clip spectrogram).  Total output = hop * (samples_in // hop), as ``convert`` returns."""
        assert not self.closed
        self.closed = True
        T = self.n_in // self.hop
        if T < 1 or self.n_in <= self.pad:
            raise ValueError("audio too short")       # shorter than one hop / the STFT reflect padding, like convert
        self._extend_spec(final=True)
        outs = []
        while self.emitted < T:
            e1 = min(T, self.emitted + self.W)
            outs.append(self._convert_window(self.emitted, e1, T))
        return np.concatenate(outs) if outs else np.zeros(0, dtype=np.float32)
