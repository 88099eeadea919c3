"""Helpers for the `-m gpu` parity tests: thin torch-tensor wrappers over the C-ABI."""
import ctypes as C

import numpy as np
import torch

import rgbd_b200
from rgbd_b200 import lib as L
from rgbd_b200.entropy_models import DeviceTables

DEV = "cuda:0"


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_tables(cdf, lengths, offsets):
    return DeviceTables(torch.as_tensor(np.asarray(cdf)), torch.as_tensor(np.asarray(lengths)),
                        torch.as_tensor(np.asarray(offsets)), DEV)


def gpu_rans_encode(sym2d, idx2d, tables, cap_words=None):
    """sym2d/idx2d: [n_streams, n_sym] arrays -> list of bytes (one per stream)"""
    sym = torch.as_tensor(np.ascontiguousarray(sym2d, dtype=np.int32)).to(DEV)
    idx = torch.as_tensor(np.ascontiguousarray(idx2d, dtype=np.uint8)).to(DEV)
    S, n = sym.shape
    cap = cap_words or (2 * n + 64)
    out = torch.zeros(S, cap, dtype=torch.int32, device=DEV)
    nw = torch.zeros(S, dtype=torch.int32, device=DEV)
    L.call("rgbd_rans_encode", sym.data_ptr(), idx.data_ptr(), n, n, S, C.byref(tables.struct), out.data_ptr(), cap,
           nw.data_ptr(), stream_ptr())
    torch.cuda.synchronize()
    nw = nw.cpu().numpy()
    host = out.cpu().numpy()
    return [host[s, cap - nw[s]:].tobytes() if nw[s] >= 0 else None for s in range(S)]


def gpu_rans_decode(streams, idx2d, tables, chunks=None):
    """streams: list of bytes; idx2d [n_streams, n_sym]; chunks: list of chunk sizes (resumable)"""
    idx = torch.as_tensor(np.ascontiguousarray(idx2d, dtype=np.uint8)).to(DEV)
    S, n = idx.shape
    lens = np.array([len(s) // 4 for s in streams], dtype=np.int64)
    offs = np.zeros_like(lens)
    offs[1:] = np.cumsum(lens[:-1])
    words = torch.as_tensor(np.frombuffer(b"".join(streams), dtype=np.int32).copy()).to(DEV)
    word_off, word_len = torch.as_tensor(offs).to(DEV), torch.as_tensor(lens).to(DEV)
    state = torch.zeros(S, 2, dtype=torch.int64, device=DEV)
    sym = torch.full((S, n), -12345, dtype=torch.int32, device=DEV)
    L.call("rgbd_rans_decode_init", words.data_ptr(), word_off.data_ptr(), S, state.data_ptr(), stream_ptr())
    p = 0
    for c in (chunks or [n]):
        c = min(c, n - p)
        if c <= 0:
            break
        L.call("rgbd_rans_decode_chunk", words.data_ptr(), word_off.data_ptr(), word_len.data_ptr(), S,
               state.data_ptr(), idx.data_ptr(), sym.data_ptr(), n, p, c, C.byref(tables.struct), stream_ptr())
        p += c
    torch.cuda.synchronize()
    return sym.cpu().numpy(), state.cpu().numpy()


def make_model(cls=None, preset="mid", seed=0, precision="fp32", **kw):
    cls = cls or rgbd_b200.ELIC_united
    net = cls(config=rgbd_b200.model_config(), channel=4, precision=precision, **kw).eval()
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, seed, preset))
    net.update(force=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return net.to(DEV), sd


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def nchw(view):
    """NHWC channel view of a launch plan -> fp32 NCHW tensor on the CPU."""
    return view.torch().float().cpu().permute(0, 3, 1, 2).contiguous()


def recon_fidelity_db(x_gpu, x_ref):
    """Distance between a GPU reconstruction and the oracle's synthesis transform run ON THE SAME y_hat, as a PSNR whose
    peak is 4 x the standard deviation of the oracle's output (the peak-to-sigma ratio of a natural image in [0, 1]).
    Why not PSNR(x_hat_gpu, x_hat_oracle) of two complete codecs: their y_hat differ wherever a bf16 rounding flips a
    quantised symbol (a few % of sites, by one step), and the random-init synthesis transform amplifies every flip to
    O(1) pixel differences (its outputs are not images: |x_hat| reaches 50), so that figure measures the chaos of the
    stand-in weights, not the accuracy of the kernels.  With y_hat held fixed the figure isolates g_s."""
    x_gpu, x_ref = x_gpu.double().cpu(), x_ref.double().cpu()
    mse = float(((x_gpu - x_ref) ** 2).mean())
    peak2 = 16.0 * float(x_ref.var())
    return 99.0 if mse == 0 else 10 * np.log10(peak2 / mse)


# bf16 mode: the reconstruction may sit at most this far below what bf16 arithmetic itself costs (oracle/bf16_emulation.py:
# 40 - 45 dB on the synthetic weights), and never below the absolute floor
FIDELITY_MARGIN_DB = 1.5
FIDELITY_FLOOR_DB = 38.0


def check_recon_fidelity(tag, x_gpu, gs_fp32, gs_bf16_emulated, floor_db=FIDELITY_FLOOR_DB):
    """x_gpu against the fp32 oracle's g_s on the same y_hat, held to the bf16 emulation's distance from that oracle
    (floor_db = None for the `stress` weights, whose scales in the hundreds put bf16 itself at ~29 dB)."""
    got = recon_fidelity_db(x_gpu, gs_fp32)
    emu = recon_fidelity_db(gs_bf16_emulated, gs_fp32)
    assert got >= emu - FIDELITY_MARGIN_DB and (floor_db is None or got >= floor_db), \
        f"{tag}: reconstruction {got:.2f} dB from the fp32 oracle; bf16 arithmetic itself costs {emu:.2f} dB"
    return got, emu
