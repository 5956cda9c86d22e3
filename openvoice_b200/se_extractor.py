"""``se_extractor.get_se`` surface of the reference (openvoice/se_extractor.py:129-152).

The reference splits the clip with third-party VAD / ASR models (silero via whisper-timestamped,
faster-whisper; pydub for slicing) that are outside this build's scope
and are not installed here.  When they are importable the reference's own splitter modules can be
passed in; otherwise the clip is cut into ~10 s pieces directly.  Only ``vc_model.device``,
``.version`` and ``.extract_se`` are used, exactly like the reference.
"""
import base64
import hashlib
import os

import numpy as np

from .api import _load_audio


def hash_numpy_array(audio_path):
    """openvoice/se_extractor.py:118-127: sha256 of the decoded samples, base64, 16 chars."""
    array = _load_audio(audio_path, None) if not isinstance(audio_path, np.ndarray) else audio_path
    digest = hashlib.sha256(np.ascontiguousarray(array).tobytes()).digest()
    return base64.b64encode(digest).decode("utf-8")[:16].replace("/", "_^")


def split_audio_fixed(audio, sr, seg_seconds=10.0, min_seconds=1.5):
    """Fallback splitter: consecutive ~10 s pieces (the VAD splitter also targets ~10 s,
    openvoice/se_extractor.py:99-112)."""
    n = int(seg_seconds * sr)
    segs = [audio[i: i + n] for i in range(0, len(audio), n)]
    segs = [s for s in segs if len(s) >= int(min_seconds * sr)]
    return segs or [audio]


def get_se(audio_path, vc_model, target_dir="processed", vad=True, splitter=None):
    """Returns (se [1, gin, 1], audio_name) like the reference.  ``splitter(audio, sr) -> list of
    waveforms`` overrides the segmentation."""
    version = vc_model.version
    print("OpenVoice version:", version)
    sr = vc_model.hps.data.sampling_rate
    audio = _load_audio(audio_path, sr)
    base = os.path.basename(audio_path).rsplit(".", 1)[0] if isinstance(audio_path, str) else "array"
    audio_name = f"{base}_{version}_{hash_numpy_array(audio)}"
    se_path = os.path.join(target_dir, audio_name, "se.pth")
    if os.path.isfile(se_path):   # openvoice/se_extractor.py:139-142: reuse the cached embedding
        import torch
        return torch.load(se_path).to(vc_model.device), audio_name
    segs = (splitter or split_audio_fixed)(audio, sr)
    if len(segs) == 0:
        raise NotImplementedError("No audio segments found!")
    return vc_model.extract_se(segs, se_save_path=se_path), audio_name
