"""On-disk container of one modality's bitstream, as the reference harness writes and reads it
(testing/tester_united.py:152-176 with utils/IOutils.py:58-88).

Layout, all integers big-endian uint32:

    orig_H  orig_W                      size of the un-padded image (for the crop after decoding)
    z_H  z_W  n_entries                 `shape` returned by compress() and the number of string groups (2: y, z)
    per entry:   n_strings
      per string:  n_bytes  <bytes>     a rANS stream: little-endian uint32 words (CompressAI layout)

One file per modality (`.../rgb/<name>`, `.../depth/<name>`).  `compress()` output goes in unchanged; what
`read_modality_file` returns is what `decompress()` takes.  Batch > 1 (several y strings in entry 0) uses the same
layout — the reference's own reader parses it, its model just reads `strings[0][0]` only.
"""
import os
import struct

_U32 = struct.Struct(">I")


def _put(fd, *vals):
    fd.write(struct.pack(">%dI" % len(vals), *vals))
    return 4 * len(vals)


def _get(fd, n):
    raw = fd.read(4 * n)
    if len(raw) != 4 * n:
        raise ValueError("truncated bitstream container")
    return struct.unpack(">%dI" % n, raw)


def write_modality(fd, orig_hw, shape, strings):
    """Write one modality to an open binary file; returns the number of bytes written."""
    n = _put(fd, int(orig_hw[0]), int(orig_hw[1]))
    n += _put(fd, int(shape[0]), int(shape[1]), len(strings))
    for group in strings:
        n += _put(fd, len(group))
        for s in group:
            s = bytes(s)
            n += _put(fd, len(s))
            fd.write(s)
            n += len(s)
    return n


MAX_SIDE = 1 << 16      # largest image side the reader accepts (a corrupt header must not size a launch plan)


def _remaining(fd):
    try:
        return os.fstat(fd.fileno()).st_size - fd.tell()
    except (AttributeError, OSError, ValueError):
        return None     # not a real file (BytesIO ...): the per-read length checks still apply


def read_modality(fd):
    """-> (orig_hw, strings, shape): strings = list (entries) of lists of bytes, shape = (z_H, z_W).
    The container itself is generic (any number of entries, like the reference's read_body); every count and
    length is checked against the bytes that are left before it drives a loop or a read.  What the codec
    additionally requires of a file is checked by `check_codec_header` (called from `load_compressed`)."""
    orig_hw = _get(fd, 2)
    zh, zw, n_entries = _get(fd, 3)
    left = _remaining(fd)
    if left is not None and 4 * n_entries > left:
        raise ValueError("corrupt bitstream container: entry count exceeds the file")
    strings = []
    for _ in range(n_entries):
        (count,) = _get(fd, 1)
        left = _remaining(fd)
        if left is not None and 4 * count > left:
            raise ValueError("corrupt bitstream container: string count exceeds the file")
        group = []
        for _ in range(count):
            (nbytes,) = _get(fd, 1)
            left = _remaining(fd)
            if left is not None and nbytes > left:
                raise ValueError("truncated bitstream container")
            s = fd.read(nbytes)
            if len(s) != nbytes:
                raise ValueError("truncated bitstream container")
            group.append(s)
        strings.append(group)
    return tuple(orig_hw), strings, (zh, zw)


def check_codec_header(orig_hw, shape, strings):
    """What ELIC_united.decompress needs of a parsed file before a launch plan is sized from it: two entries (y, z),
    a sane image size, latent shape = ceil(orig / 64) (dataset/utils.py:58-67 pads to multiples of 64), one z string
    per image and as many y strings."""
    if len(strings) != 2:
        raise ValueError(f"corrupt bitstream container: {len(strings)} entries (expected 2: y, z)")
    if not (0 < orig_hw[0] <= MAX_SIDE and 0 < orig_hw[1] <= MAX_SIDE):
        raise ValueError(f"corrupt bitstream container: image size {tuple(orig_hw)}")
    if tuple(shape) != (-(-orig_hw[0] // 64), -(-orig_hw[1] // 64)):
        raise ValueError(f"corrupt bitstream container: latent shape {tuple(shape)} does not belong to a "
                         f"{tuple(orig_hw)} image")
    if not strings[1] or len(strings[0]) % len(strings[1]):
        raise ValueError("corrupt bitstream container: y / z string counts do not match")


def write_modality_file(path, orig_hw, shape, strings):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fd:
        return write_modality(fd, orig_hw, shape, strings)


def read_modality_file(path):
    with open(path, "rb") as fd:
        return read_modality(fd)


def save_compressed(out, orig_hw, rgb_path, depth_path):
    """Write the dict returned by `ELIC_united.compress` as the two files the reference tester produces.
    Returns (rgb_bpp, depth_bpp) computed like the tester: file size * 8 / (orig_H * orig_W)."""
    npx = float(orig_hw[0] * orig_hw[1])
    nr = write_modality_file(rgb_path, orig_hw, out["shape"], out["r_strings"])
    nd = write_modality_file(depth_path, orig_hw, out["shape"], out["d_strings"])
    return nr * 8.0 / npx, nd * 8.0 / npx


def load_compressed(rgb_path, depth_path):
    """-> (rgb_strings, depth_strings, shape, orig_hw), the arguments of `ELIC_united.decompress` (+ the crop size)."""
    hw_r, rs, shape_r = read_modality_file(rgb_path)
    hw_d, ds, shape_d = read_modality_file(depth_path)
    if hw_r != hw_d or shape_r != shape_d:
        raise ValueError("rgb and depth containers disagree on the image / latent size")
    check_codec_header(hw_r, shape_r, rs)
    check_codec_header(hw_d, shape_d, ds)
    return rs, ds, shape_r, hw_r
