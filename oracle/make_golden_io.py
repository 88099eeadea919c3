"""ORACLE (test infrastructure): container known-answer files made by the UNMODIFIED reference
(utils/IOutils.py write_uints / write_body, exactly as testing/tester_united.py:152-162 calls them).

    python -m oracle.make_golden_io          # build container only (needs /root/reference)
"""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, "/root/reference")
    from utils.IOutils import read_body, read_uints, write_body, write_uints   # the reference's own functions
    rng = np.random.default_rng(7)
    cases = {}

    def add(name, hw, shape, strings):
        fd = io.BytesIO()
        write_uints(fd, hw)
        write_body(fd, shape, strings)
        blob = fd.getvalue()
        fd.seek(0)
        assert tuple(read_uints(fd, 2)) == tuple(hw)
        back, shp = read_body(fd)
        assert back == [list(g) for g in strings] and tuple(shp) == tuple(shape)
        cases[name] = (hw, shape, strings, blob)

    def stream(n):
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()

    add("batch1", (480, 640), (8, 10), [[stream(5004)], [stream(2080)]])
    add("batch3_per_image_y", (530, 730), (9, 12), [[stream(12), stream(8), stream(1000)], [stream(8)] * 3])
    add("tiny", (128, 128), (2, 2), [[stream(8)], [stream(8)]])
    add("three_entries", (64, 64), (1, 1), [[stream(16)], [stream(4), stream(4)], [stream(40)]])
    out = {}
    for name, (hw, shape, strings, blob) in cases.items():
        out[name + ".hw"] = np.array(hw, dtype=np.int64)
        out[name + ".shape"] = np.array(shape, dtype=np.int64)
        out[name + ".counts"] = np.array([len(g) for g in strings], dtype=np.int64)
        out[name + ".lens"] = np.array([len(s) for g in strings for s in g], dtype=np.int64)
        out[name + ".payload"] = np.frombuffer(b"".join(s for g in strings for s in g), dtype=np.uint8)
        out[name + ".file"] = np.frombuffer(blob, dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "container_kat.npz"), **out)
    print("wrote", os.path.join(GOLD, "container_kat.npz"), {k: len(v[3]) for k, v in cases.items()})


if __name__ == "__main__":
    main()
