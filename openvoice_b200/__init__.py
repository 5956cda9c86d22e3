"""openvoice_b200 -- B200-native (sm_100a) tone-colour-converter hot path of OpenVoice.

Drop-in surface of the reference's ``openvoice.api`` / ``openvoice.se_extractor`` for the
``ToneColorConverter.convert -> SynthesizerTrn.voice_conversion`` path; the arithmetic runs in
``libovc_b200.so`` (hand-written CUDA, C ABI in ``include/ovc.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import utils  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA
    import importlib
    if name in ("ToneColorConverter", "OpenVoiceBaseClass"):
        return getattr(importlib.import_module(".api", __name__), name)
    if name in ("api", "se_extractor", "mel_processing", "schema", "ref_enc"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
