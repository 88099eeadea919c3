"""GPU rANS coder == compressai.ans, byte for byte (BASELINE north_star level 1).
Golden vectors come from the unmodified reference coder; large cases are checked against the
C oracle (itself pinned to those vectors)."""
import numpy as np
import pytest

from oracle import coder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(golden_dir):
    from gpu_utils import device_tables
    g = np.load(f"{golden_dir}/gauss_tables.npz")
    z = np.load(f"{golden_dir}/rans_kat.npz")
    names = sorted({k.rsplit(".", 1)[0] for k in z.files})
    kats = {n: (z[n + ".sym"], z[n + ".idx"], z[n + ".bytes"].tobytes()) for n in names}
    return dict(dev=device_tables(g["cdf"], g["lengths"], g["offsets"]),
                host=coder.Tables(g["cdf"], g["lengths"], g["offsets"]), kats=kats)


def test_encode_golden_bytes(G):
    from gpu_utils import gpu_rans_encode
    for name, (sym, idx, want) in G["kats"].items():
        got = gpu_rans_encode(sym[None], idx[None], G["dev"])[0]
        assert got == want, name


def test_decode_golden_one_shot_and_resumable(G):
    from gpu_utils import gpu_rans_decode
    for name, (sym, idx, stream) in G["kats"].items():
        got, _ = gpu_rans_decode([stream], idx[None], G["dev"])
        assert np.array_equal(got[0], sym), name
        got, _ = gpu_rans_decode([stream], idx[None], G["dev"], chunks=[1, 2, 29, 32, 33, 1000, 10 ** 9])
        assert np.array_equal(got[0], sym), name + " (chunked)"


def test_batched_streams_full_size_vs_oracle(G):
    """8 independent y-sized streams (409 600 symbols each, NYUv2 480x640) in one launch."""
    from gpu_utils import gpu_rans_decode, gpu_rans_encode
    rng = np.random.default_rng(2024)
    S, n = 8, 409600
    idx = rng.integers(0, 64, (S, n)).astype(np.uint8)
    scale = np.exp(np.linspace(np.log(0.11), np.log(256), 64))[idx]
    sym = np.rint(rng.standard_normal((S, n)) * scale * rng.uniform(0.2, 1.5, (S, 1))).astype(np.int32)
    sym[3, ::50021] = 5000
    sym[3, 7::70001] = -4000
    got = gpu_rans_encode(sym, idx, G["dev"])
    for s in range(S):
        assert got[s] == coder.encode_with_indexes(sym[s], idx[s].astype(np.int32), G["host"]), s
    # the 10-chunk schedule of one image (5 groups x anchor/non-anchor)
    chunks = []
    for g in (16, 16, 32, 64, 192):
        chunks += [g * 32 * 20] * 2
    dec, state = gpu_rans_decode(got, idx, G["dev"], chunks=chunks)
    assert np.array_equal(dec, sym)
    # every stream consumed exactly: next word == stream length
    assert state[:, 1].tolist() == [len(b) // 4 for b in got]


def test_edge_cases(G):
    from gpu_utils import gpu_rans_decode, gpu_rans_encode
    # single symbol (UB in the reference, defined here and equal to the oracle)
    for sym, idx in (([0], [0]), ([3], [10]), ([-900], [2])):
        s = np.array([sym], dtype=np.int32)
        i = np.array([idx], dtype=np.uint8)
        got = gpu_rans_encode(s, i, G["dev"])[0]
        assert got == coder.encode_with_indexes(s[0], i[0].astype(np.int32), G["host"])
        assert gpu_rans_decode([got], i, G["dev"])[0][0].tolist() == sym
    # overflow is reported, not written past the buffer
    rng = np.random.default_rng(1)
    sym = rng.integers(-30000, 30000, (1, 4096)).astype(np.int32)
    idx = np.zeros((1, 4096), dtype=np.uint8)
    assert gpu_rans_encode(sym, idx, G["dev"], cap_words=64)[0] is None
    # ragged: streams of different entropy in one batch, lengths differ
    idx = rng.integers(0, 64, (3, 1000)).astype(np.uint8)
    sym = np.stack([np.zeros(1000), rng.integers(-3, 4, 1000), rng.integers(-2000, 2000, 1000)]).astype(np.int32)
    got = gpu_rans_encode(sym, idx, G["dev"])
    assert len({len(b) for b in got}) == 3
    for s in range(3):
        assert got[s] == coder.encode_with_indexes(sym[s], idx[s].astype(np.int32), G["host"])
    dec, _ = gpu_rans_decode(got, idx, G["dev"], chunks=[500, 500])
    assert np.array_equal(dec, sym)


def test_z_tables_192_channels():
    """EntropyBottleneck tables: 192 tables, index = channel id."""
    import rgbd_b200
    from gpu_utils import device_tables, gpu_rans_decode, gpu_rans_encode
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, "mid"))
    net.update(force=True)
    eb = net.rgb_entropy_bottleneck
    host = coder.Tables(eb.quantized_cdf.numpy(), eb.cdf_length.numpy(), eb.offset.numpy())
    dev = device_tables(eb.quantized_cdf.numpy(), eb.cdf_length.numpy(), eb.offset.numpy())
    rng = np.random.default_rng(5)
    idx = np.tile(np.repeat(np.arange(192, dtype=np.uint8), 80), (4, 1))
    sym = np.rint(rng.standard_normal(idx.shape) * 6).astype(np.int32)
    got = gpu_rans_encode(sym, idx, dev)
    for s in range(4):
        assert got[s] == coder.encode_with_indexes(sym[s], idx[s].astype(np.int32), host)
    dec, _ = gpu_rans_decode(got, idx, dev)
    assert np.array_equal(dec, sym)
