"""Config loading: the JSON -> attribute-tree loader of the reference (openvoice/utils.py:6-43),
plus the watermark payload codec (openvoice/utils.py:46-75).  Same names and behaviour."""
import json

import numpy as np


class HParams:
    """Attribute tree over a (nested) dict; behaves like the reference's HParams
    (openvoice/utils.py:14-43): item and attribute access, ``in``, ``len``, keys/items/values."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            self[k] = HParams(**v) if isinstance(v, dict) else v

    def keys(self):
        return self.__dict__.keys()

    def items(self):
        return self.__dict__.items()

    def values(self):
        return self.__dict__.values()

    def __len__(self):
        return len(self.__dict__)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__

    def __repr__(self):
        return repr(self.__dict__)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, HParams) else v) for k, v in self.__dict__.items()}


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
