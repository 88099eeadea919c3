"""Pins the plain-C rANS / pmf->cdf restatement (oracle/rans_oracle.c) against golden vectors the
UNMODIFIED reference coder produced (tests/golden/rans_kat.npz, pmf_kat.npz; oracle/make_golden.py)
and, where oracle/_ref exists, against the reference's compiled module live."""
import numpy as np
import pytest

from oracle import coder
from oracle.ref_loader import ref_ext_available


@pytest.fixture(scope="module")
def tables(golden_dir):
    g = np.load(f"{golden_dir}/gauss_tables.npz")
    return coder.Tables(g["cdf"], g["lengths"], g["offsets"])


@pytest.fixture(scope="module")
def kats(golden_dir):
    z = np.load(f"{golden_dir}/rans_kat.npz")
    names = sorted({k.rsplit(".", 1)[0] for k in z.files})
    return {n: (z[n + ".sym"], z[n + ".idx"].astype(np.int32), z[n + ".bytes"].tobytes()) for n in names}


def test_encode_matches_reference_bytes(kats, tables):
    assert len(kats) >= 10
    for name, (sym, idx, want) in kats.items():
        assert coder.encode_with_indexes(sym, idx, tables) == want, name


def test_decode_one_shot_and_chunked(kats, tables):
    for name, (sym, idx, stream) in kats.items():
        assert np.array_equal(coder.decode_with_indexes(stream, idx, tables), sym), name
        d = coder.Decoder(stream)
        got, p = [], 0
        for c in (1, 2, 29, 32, 33, 1000, 10 ** 9):
            if p >= len(idx):
                break
            got.append(d.decode_stream(idx[p:p + c], tables))
            p += c
        assert np.array_equal(np.concatenate(got), sym), name


def test_single_symbol_stream_is_defined_here(tables):
    # UB in the reference (2-word flush into a 1-word buffer); the restatement must still round-trip
    s = coder.encode_with_indexes([0], [0], tables)
    assert len(s) == 8
    assert coder.decode_with_indexes(s, [0], tables).tolist() == [0]


def test_pmf_to_quantized_cdf(golden_dir):
    z = np.load(f"{golden_dir}/pmf_kat.npz")
    names = sorted({k.rsplit(".", 1)[0] for k in z.files})
    assert "steal_left" in names and "steal_right" in names
    for n in names:
        assert np.array_equal(coder.pmf_to_quantized_cdf(z[n + ".pmf"]), z[n + ".cdf"]), n


@pytest.mark.skipif(not ref_ext_available(), reason="oracle/_ref not built (reference sources absent)")
def test_live_against_reference_module(tables):
    from oracle import ref_coder
    rng = np.random.default_rng(99)
    for n in (2, 100, 4097):
        idx = rng.integers(0, 64, n).astype(np.int32)
        sym = np.rint(rng.standard_normal(n) * np.exp(rng.uniform(-2, 5, n))).astype(np.int32)
        want = ref_coder.encode_with_indexes(sym, idx, tables)
        assert coder.encode_with_indexes(sym, idx, tables) == want
        assert np.array_equal(ref_coder.decode_with_indexes(want, idx, tables), sym)
    p = (rng.random(50) ** 6).astype(np.float32)
    assert np.array_equal(coder.pmf_to_quantized_cdf(p), ref_coder.pmf_to_quantized_cdf(p))
