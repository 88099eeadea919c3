"""ORACLE (test infrastructure): ctypes wrapper of the plain-C rANS / pmf->cdf restatement
(oracle/rans_oracle.c).  Mirrors the call shapes of compressai.ans
(reference rans_interface.cpp:353-373) so tests read like the reference's call sites."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "librgbd_oracle.so")
        src = os.path.join(HERE, "rans_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", path, src])
        _LIB = C.CDLL(path)
        _LIB.rgbd_oracle_rans_encode.restype = C.c_int64
        _LIB.rgbd_oracle_rans_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int64]
        _LIB.rgbd_oracle_rans_decode_init.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.rgbd_oracle_rans_decode_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.rgbd_oracle_pmf_to_quantized_cdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return _LIB


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class Tables:
    """(cdf int32 [n, stride], cdf_length [n], offset [n]) as the reference passes them."""

    def __init__(self, cdf, lengths, offsets):
        self.cdf = _i32(cdf)
        self.lengths = _i32(lengths).reshape(-1)
        self.offsets = _i32(offsets).reshape(-1)
        assert self.cdf.ndim == 2 and self.cdf.shape[0] == self.lengths.size == self.offsets.size


def encode_with_indexes(symbols, indexes, t):
    """RansEncoder.encode_with_indexes / BufferedRansEncoder + flush -> bytes"""
    sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    assert sym.size == idx.size
    cap = 8 * (sym.size + 16) * 4
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().rgbd_oracle_rans_encode(sym.ctypes.data, idx.ctypes.data, sym.size, t.cdf.ctypes.data,
                                      t.cdf.shape[1], t.lengths.ctypes.data, t.offsets.ctypes.data,
                                      out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"oracle encode failed ({n})")
    return out[:n].tobytes()


class Decoder:
    """RansDecoder: set_stream + resumable decode_stream."""

    def __init__(self, stream):
        self.stream = np.frombuffer(bytes(stream) + b"\0" * 64, dtype=np.uint8).copy()
        self.state = np.zeros(2, dtype=np.int64)
        lib().rgbd_oracle_rans_decode_init(self.state.ctypes.data, self.stream.ctypes.data)

    def decode_stream(self, indexes, t):
        idx = _i32(indexes).reshape(-1)
        out = np.zeros(idx.size, dtype=np.int32)
        lib().rgbd_oracle_rans_decode_chunk(self.state.ctypes.data, self.stream.ctypes.data, idx.ctypes.data,
                                            idx.size, t.cdf.ctypes.data, t.cdf.shape[1], t.lengths.ctypes.data,
                                            t.offsets.ctypes.data, out.ctypes.data)
        return out


def decode_with_indexes(stream, indexes, t):
    return Decoder(stream).decode_stream(indexes, t)


def pmf_to_quantized_cdf(pmf, precision=16):
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib().rgbd_oracle_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, out.ctypes.data)
    if rc:
        raise RuntimeError("oracle pmf_to_quantized_cdf failed")
    return out.astype(np.int64)
