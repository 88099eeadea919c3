"""On-disk container (utils/IOutils.py:58-88 as called by testing/tester_united.py:152-176): our writer produces the
reference's bytes, our reader parses the reference's files (golden made by oracle/make_golden_io.py from the
unmodified reference)."""
import io
import os

import numpy as np
import pytest

from rgbd_b200 import bitstream_io as bio


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "container_kat.npz"))
    for name in sorted({k.rsplit(".", 1)[0] for k in z.files}):
        counts, lens = z[name + ".counts"].tolist(), z[name + ".lens"].tolist()
        payload = z[name + ".payload"].tobytes()
        strings, p, k = [], 0, 0
        for c in counts:
            group = []
            for _ in range(c):
                group.append(payload[p:p + lens[k]])
                p += lens[k]
                k += 1
            strings.append(group)
        yield name, tuple(z[name + ".hw"].tolist()), tuple(z[name + ".shape"].tolist()), strings, z[name + ".file"].tobytes()


def test_writer_matches_reference_bytes_and_reader_parses_them(golden_dir):
    n = 0
    for name, hw, shape, strings, blob in _cases(golden_dir):
        fd = io.BytesIO()
        assert bio.write_modality(fd, hw, shape, strings) == len(blob), name
        assert fd.getvalue() == blob, name
        got_hw, got_strings, got_shape = bio.read_modality(io.BytesIO(blob))
        assert (got_hw, got_shape, got_strings) == (hw, shape, strings), name
        n += 1
    assert n >= 4


def test_file_helpers_roundtrip_and_bpp(tmp_path):
    out = {"r_strings": [[b"\x01" * 40], [b"\x02" * 8]], "d_strings": [[b"\x03" * 24], [b"\x04" * 12]], "shape": (8, 10)}
    rp, dp = str(tmp_path / "rgb" / "a.bin"), str(tmp_path / "depth" / "a.bin")
    bpp_r, bpp_d = bio.save_compressed(out, (480, 640), rp, dp)
    # header 5 words + 2 group counts + 2 string lengths = 9 words = 36 bytes, + payload
    assert os.path.getsize(rp) == 36 + 48 and bpp_r == (36 + 48) * 8 / (480 * 640)
    assert os.path.getsize(dp) == 36 + 36 and bpp_d == (36 + 36) * 8 / (480 * 640)
    rs, ds, shape, hw = bio.load_compressed(rp, dp)
    assert (rs, ds, shape, hw) == (out["r_strings"], out["d_strings"], (8, 10), (480, 640))


def test_truncated_file_raises():
    fd = io.BytesIO()
    bio.write_modality(fd, (64, 64), (1, 1), [[b"abcdefgh"], [b"12345678"]])
    blob = fd.getvalue()
    for cut in (3, 19, len(blob) - 1):
        with pytest.raises(ValueError, match="truncated"):
            bio.read_modality(io.BytesIO(blob[:cut]))
