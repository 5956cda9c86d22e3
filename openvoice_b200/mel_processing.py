"""Linear-spectrogram front end of ``ToneColorConverter.convert`` (openvoice/mel_processing.py:40-75).

Same name and arguments as the reference's ``spectrogram_torch``.  Arithmetic: reflect-pad
(n_fft-hop)/2 on both sides, 1024-point periodic-hann STFT with centre=False, one-sided,
magnitude sqrt(re^2 + im^2 + 1e-6).  On a CUDA tensor the FFT is cuFFT via torch.stft (library
call; SURVEY.md section 8 row a2 / f1: < 0.3 % of the path's time).  Unlike the reference this
does not print min/max range warnings (two device syncs, mel_processing.py:41-44).
"""
import torch

hann_window = {}


def spectrogram_torch(y, n_fft, sampling_rate, hop_size, win_size, center=False):
    """y: [B, L] float waveform -> [B, n_fft//2+1, L//hop] magnitude spectrogram."""
    key = f"{win_size}_{y.dtype}_{y.device}"
    if key not in hann_window:
        hann_window[key] = torch.hann_window(win_size, dtype=y.dtype, device=y.device)
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=hann_window[key],
                      center=center, pad_mode="reflect", normalized=False, onesided=True,
                      return_complex=True)
    spec = torch.view_as_real(spec)
    return torch.sqrt(spec.pow(2).sum(-1) + 1e-6)
