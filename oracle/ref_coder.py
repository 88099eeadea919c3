"""ORACLE (test infrastructure): the reference's OWN compiled coder (oracle/_ref/ans*.so, built
by oracle/Makefile from /root/reference sources) behind the same call shapes as oracle/coder.py."""
import numpy as np

from .coder import Tables  # noqa: F401
from .ref_loader import load_ref_ext


def _lists(t):
    return t.cdf.tolist(), t.lengths.tolist(), t.offsets.tolist()


def encode_with_indexes(symbols, indexes, t):
    ans = load_ref_ext("ans")
    cdf, ln, off = _lists(t)
    return ans.RansEncoder().encode_with_indexes(np.asarray(symbols).reshape(-1).tolist(),
                                                 np.asarray(indexes).reshape(-1).tolist(), cdf, ln, off)


class Decoder:
    def __init__(self, stream):
        self._d = load_ref_ext("ans").RansDecoder()
        self._d.set_stream(bytes(stream))
        self._lists = None

    def decode_stream(self, indexes, t):
        if self._lists is None:
            self._lists = _lists(t)
        cdf, ln, off = self._lists
        return np.asarray(self._d.decode_stream(np.asarray(indexes).reshape(-1).tolist(), cdf, ln, off), dtype=np.int32)


def decode_with_indexes(stream, indexes, t):
    return Decoder(stream).decode_stream(indexes, t)


def pmf_to_quantized_cdf(pmf, precision=16):
    cxx = load_ref_ext("_CXX")
    return np.asarray(cxx.pmf_to_quantized_cdf(np.asarray(pmf, dtype=np.float32).tolist(), precision), dtype=np.int64)
