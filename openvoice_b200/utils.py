"""Config loading: the JSON -> attribute-tree loader of the reference (openvoice/utils.py:6-43),
plus the watermark payload codec (openvoice/utils.py:46-75).  Same names and behaviour."""
import json

import numpy as np


class HParams(dict):
    """Nested config tree with attribute *and* item access (``hps.data.hop_length``, ``hps["model"]``,
    ``"zero_g" in hps.model``, ``len``, ``keys/items/values``) -- the behaviour callers of the reference's
    ``utils.HParams`` (openvoice/utils.py:14-43) rely on, implemented as a dict subclass."""

    def __init__(self, **entries):
        super().__init__()
        for key, value in entries.items():
            self[key] = HParams(**value) if isinstance(value, dict) else value

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, HParams) else v) for k, v in self.items()}


def get_hparams_from_file(config_path):
    """openvoice/utils.py:6-12."""
    with open(config_path, "r", encoding="utf-8") as f:
        return HParams(**json.loads(f.read()))


def string_to_bits(string, pad_len=8):
    """ASCII string -> [pad_len, 8] bit matrix (openvoice/utils.py:44-60): one row per
    character (MSB first), truncated to ``pad_len`` rows and padded with the bits of an ASCII
    space (only bit 2 set)."""
    rows = [[(ord(ch) >> (7 - i)) & 1 for i in range(8)] for ch in string[:pad_len]]
    full = np.zeros((pad_len, 8), dtype=np.int64)
    full[:, 2] = 1
    if rows:
        full[: len(rows)] = np.array(rows, dtype=np.int64)
    return full


def bits_to_string(bits_array):
    """Inverse of string_to_bits (openvoice/utils.py:65-75)."""
    return "".join(chr(int("".join(str(int(b)) for b in row), 2)) for row in bits_array)
