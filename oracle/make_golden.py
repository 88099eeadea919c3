"""ORACLE (test infrastructure): regenerate tests/golden/* by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python -m oracle.make_golden
The fixtures pin the oracle restatement and the CUDA path on machines where the reference
does not exist (the GPU box).  Everything is seeded; inputs and weights come from
rgbd_b200.synthetic so they can be rebuilt anywhere.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def rans_kats(ans, net):
    gc = net.rgb_gaussian_conditional
    cdf = gc.quantized_cdf.numpy()
    lens = gc.cdf_length.numpy()
    offs = gc.offset.numpy()
    cdf_l, lens_l, offs_l = cdf.tolist(), lens.tolist(), offs.tolist()
    rng = np.random.default_rng(1234)
    cases = {}

    def add(name, sym, idx, chunks=None):
        sym = np.asarray(sym, dtype=np.int32)
        idx = np.asarray(idx, dtype=np.int32)
        enc = ans.BufferedRansEncoder()
        enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdf_l, lens_l, offs_l)
        s1 = enc.flush()
        s2 = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdf_l, lens_l, offs_l)
        assert s1 == s2
        dec = ans.RansDecoder().decode_with_indexes(s1, idx.tolist(), cdf_l, lens_l, offs_l)
        assert dec == sym.tolist(), name
        if chunks:
            d = ans.RansDecoder()
            d.set_stream(s1)
            got, p = [], 0
            for c in chunks:
                got += d.decode_stream(idx[p:p + c].tolist(), cdf_l, lens_l, offs_l)
                p += c
            assert got == sym.tolist()
        cases[name] = (sym, idx, np.frombuffer(s1, dtype=np.uint8))

    n = 40000
    idx = rng.integers(0, 64, n)
    scale = np.exp(np.linspace(np.log(0.11), np.log(256), 64))[idx]
    sym = np.rint(rng.standard_normal(n) * scale).astype(np.int32)
    add("gauss_mixed", sym, idx, chunks=[1000, 1, 31, 32, 33, 7935, n - 9000 - 32])
    sym2 = sym.copy()
    sym2[::997] = 5000
    sym2[7::1499] = -4000
    sym2[11::2999] = 70000      # 5-nibble escapes
    sym2[13::3001] = -(1 << 20)
    add("escapes", sym2, idx, chunks=[n // 2, n - n // 2])
    # NOTE: a 1-record stream is UB in the reference (flush() writes 2 words into a 1-word
    # buffer, rans_interface.cpp:171-187 -> heap corruption), so the smallest KAT has 2 records.
    add("one_escape", [-77], [3])
    add("two_symbols", [1, -1], [5, 63])
    add("max_table_only", np.rint(rng.standard_normal(3000) * 256).astype(np.int32), np.full(3000, 63))
    add("min_table_only", rng.integers(-1, 2, 5000), np.zeros(5000, dtype=np.int32))
    add("all_zero_sym", np.zeros(4096, dtype=np.int32), rng.integers(0, 64, 4096))
    add("len31", sym[:31], idx[:31])
    add("len32", sym[:32], idx[:32])
    add("len33", sym[:33], idx[:33])
    out = {}
    for k, (s, i, b) in cases.items():
        out[k + ".sym"], out[k + ".idx"], out[k + ".bytes"] = s, i.astype(np.uint8), b
    np.savez_compressed(os.path.join(GOLD, "rans_kat.npz"), **out)
    np.savez_compressed(os.path.join(GOLD, "gauss_tables.npz"), cdf=cdf, lengths=lens, offsets=offs,
                        scale_table=gc.scale_table.numpy())
    print("rans KATs:", {k: len(v[2]) for k, v in cases.items()})


def pmf_kats(cxx):
    rng = np.random.default_rng(7)
    out = {}
    cases = {
        "uniform8": np.full(8, 1 / 8, dtype=np.float32),
        "steal_left": np.array([0.9, 1e-9, 1e-9, 0.05, 0.05 - 2e-9], dtype=np.float32),
        "steal_right": np.array([1e-9, 1e-9, 0.5, 0.3, 0.2 - 2e-9], dtype=np.float32),
        "many_tiny": np.concatenate([np.full(200, 1e-8), [1 - 2e-6]]).astype(np.float32),
        "random64": (lambda p: (p / p.sum()).astype(np.float32))(rng.random(64) ** 4),
        "unnormalised": (rng.random(33) * 3).astype(np.float32),
    }
    for k, p in cases.items():
        out[k + ".pmf"] = p
        out[k + ".cdf"] = np.asarray(cxx.pmf_to_quantized_cdf(p.tolist(), 16), dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "pmf_kat.npz"), **out)
    print("pmf KATs:", list(cases))


def index_kats(net):
    gc = net.rgb_gaussian_conditional
    table = gc.scale_table.numpy()
    vals = [np.float32(0.0), np.float32(-1.0), np.float32(0.11), np.float32(1e9)]
    for t in table:
        vals += [np.nextafter(t, np.float32(0)), t, np.nextafter(t, np.float32(1e9))]
    rng = np.random.default_rng(3)
    vals += list(np.exp(rng.uniform(np.log(0.05), np.log(400), 4000)).astype(np.float32))
    s = torch.tensor(np.array(vals, dtype=np.float32)).reshape(1, 1, 1, -1)
    idx = gc.build_indexes(s).reshape(-1).numpy()
    np.savez_compressed(os.path.join(GOLD, "index_kat.npz"), scales=s.reshape(-1).numpy(), indexes=idx)
    print("index KAT:", idx.shape)


def model_golden(cls, name, cross, H, W, preset, seed):
    from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict
    from config.config import model_config
    net = cls(config=model_config(), channel=4).eval()
    net.load_state_dict(synthetic_state_dict(net, seed, preset))
    net.update(force=True)
    rgb, depth = synthetic_pairs(1, H, W, seed=4321)
    with torch.no_grad():
        yr, yd = net.g_a(rgb, depth)
        zr, zd = net.h_a(yr, yd)
        c = net.compress(rgb, depth)
        d = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        f = net(rgb, depth)
    out = {
        "y_r": yr.numpy(), "y_d": yd.numpy(), "z_r": zr.numpy(), "z_d": zd.numpy(),
        "xhat_r": d["x_hat"]["r"].numpy(), "xhat_d": d["x_hat"]["d"].numpy(),
        "fwd_xhat_r": f["x_hat"]["r"].numpy(), "fwd_xhat_d": f["x_hat"]["d"].numpy(),
        "lik_y_r": f["r_likelihoods"]["y"].numpy(), "lik_y_d": f["d_likelihoods"]["y"].numpy(),
        "lik_z_r": f["r_likelihoods"]["z"].numpy(), "lik_z_d": f["d_likelihoods"]["z"].numpy(),
        "shape": np.array(list(c["shape"])),
    }
    for key, tag in (("r_strings", "r"), ("d_strings", "d")):
        out[f"{tag}_y"] = np.frombuffer(c[key][0][0], dtype=np.uint8)
        out[f"{tag}_z"] = np.frombuffer(c[key][1][0], dtype=np.uint8)
    out["meta"] = np.array(json.dumps(dict(H=H, W=W, preset=preset, seed=seed, input_seed=4321, cross=cross)))
    np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **out)
    keys = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(GOLD, f"state_dict_keys_{name}.json"), "w") as fh:
        json.dump(keys, fh, indent=0)
    print(name, "bytes", {k: len(out[k]) for k in ("r_y", "r_z", "d_y", "d_z")})


def main():
    from oracle.ref_loader import import_reference, load_ref_ext
    os.makedirs(GOLD, exist_ok=True)
    U, R2D, cfg = import_reference()
    ans, cxx = load_ref_ext("ans"), load_ref_ext("_CXX")
    net = U(config=cfg(), channel=4).eval()
    net.update(force=True)
    rans_kats(ans, net)
    pmf_kats(cxx)
    index_kats(net)
    model_golden(U, "united", True, 128, 128, "mid", 0)
    model_golden(R2D, "r2d", False, 128, 192, "mid", 0)


if __name__ == "__main__":
    main()
