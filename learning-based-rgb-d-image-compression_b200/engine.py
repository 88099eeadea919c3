"""Launch-plan compiler: turns the ELIC_united layer graph into a static list of C-ABI
kernel launches over pre-allocated NHWC buffers in HBM.

A `Program` is built once per (kind, batch, H, W, precision): buffer addresses never change,
so the launch list can be replayed as-is or captured into a CUDA graph.  torch is used for
device memory and streams only; every arithmetic op on the path is one of our kernels.
"""
import ctypes as C

import torch

from . import lib as L

_DT = {torch.float32: L.DT_F32, torch.bfloat16: L.DT_BF16}


class View:
    """Channel-slice view [coff, coff+C) of an NHWC buffer."""

    __slots__ = ("buf", "coff", "C")

    def __init__(self, buf, coff=0, C=None):
        self.buf = buf
        self.coff = int(coff)
        self.C = int(buf.shape[3] - coff if C is None else C)
        assert 0 <= self.coff and self.coff + self.C <= buf.shape[3], (buf.shape, coff, C)

    N = property(lambda s: s.buf.shape[0])
    H = property(lambda s: s.buf.shape[1])
    W = property(lambda s: s.buf.shape[2])
    cstride = property(lambda s: s.buf.shape[3])
    dtype = property(lambda s: s.buf.dtype)

    def ptr(self):
        return self.buf.data_ptr()

    def sub(self, coff, C):
        return View(self.buf, self.coff + coff, C)

    def torch(self):
        return self.buf[..., self.coff:self.coff + self.C]


class PackedConv:
    """Device-resident packed weights of one nn.Conv2d / nn.ConvTranspose2d.

    fp32 CUDA-core layout: [k*k taps][Cin][cout_pad] (cout fastest).  `in_perm` reorders the
    input channels (used where our concat buffers are laid out differently from the
    reference's torch.cat order).
    """

    def __init__(self, mod, device, in_perm=None, split3=False, s2d=False):
        w = mod.weight.detach().to(device=device, dtype=torch.float32)
        self.transposed = isinstance(mod, torch.nn.ConvTranspose2d)
        if self.transposed:
            cin, cout, kh, kw = w.shape
            taps = w.permute(2, 3, 0, 1)  # [kh, kw, cin, cout]
        else:
            cout, cin, kh, kw = w.shape
            taps = w.permute(2, 3, 1, 0)
        assert kh == kw
        self.k = kh
        self.stride = mod.stride[0]
        self.pad = mod.padding[0]
        self.Cin, self.Cout = cin, cout
        taps = taps.reshape(kh * kw, cin, cout)
        if in_perm is not None:
            taps = taps[:, in_perm.to(device), :]
        self.flop_cin = cin
        if split3:
            # two-term bf16 expansion of the weights against an input stored as [hi | lo | hi]:
            # x*w ~= hi*w_hi + lo*w_hi + hi*w_lo  (see rgbd_nchw_to_nhwc split3)
            hi = taps.to(torch.bfloat16).to(torch.float32)
            taps = torch.cat([hi, hi, taps - hi], dim=1)
            cin = self.Cin = 3 * cin
        self.flop_taps = None
        if s2d:
            # Conv2d(k5, s2, p2) over an image == Conv2d(k3, s1, p1) over its space-to-depth map (rgbd_nchw_to_nhwc
            # split3 = 2): tap (qy, qx) of parity block (ry, rx) is filter tap (dy, dx) = (2 qy + ry, 2 qx + rx)
            assert not self.transposed and kh == 5 and self.stride == 2 and self.pad == 2
            t3 = torch.zeros(9, 4 * cin, cout, device=device, dtype=torch.float32)
            for qy in (-1, 0, 1):
                for qx in (-1, 0, 1):
                    for ry in range(2):
                        for rx in range(2):
                            dy, dx = 2 * qy + ry, 2 * qx + rx
                            if -2 <= dy <= 2 and -2 <= dx <= 2:
                                blk = 2 * ry + rx
                                t3[(qy + 1) * 3 + qx + 1, blk * cin:(blk + 1) * cin] = taps[(dy + 2) * 5 + dx + 2]
            taps = t3
            self.flop_taps = 25          # algorithmic count stays the dense 5x5
            self.k, self.stride, self.pad = 3, 1, 1
            kh = kw = 3
            cin = self.Cin = 4 * cin
        self.cout_pad = (cout + 15) // 16 * 16
        packed = torch.zeros(kh * kw, cin, self.cout_pad, device=device, dtype=torch.float32)
        packed[:, :, :cout] = taps
        self.w32 = packed.contiguous()
        self.bias = None if mod.bias is None else mod.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        self._wtc = None
        self._wtc32 = None

    @property
    def wtc32(self):
        """fp32 master copy in the tensor-core layout [taps][cout_pad][cin_pad]: the source of the per-image
        SE-folded filters (rgbd_scale_weights rounds w * scale to bf16 once, from fp32)."""
        if self._wtc32 is None:
            T, cin, cp = self.w32.shape
            self.cin_pad = (cin + 63) // 64 * 64
            w = torch.zeros(T, cp, self.cin_pad, device=self.w32.device, dtype=torch.float32)
            w[:, :, :cin] = self.w32.permute(0, 2, 1)
            self._wtc32 = w.contiguous()
        return self._wtc32

    @property
    def wtc(self):
        """bf16 [taps][cout_pad][cin_pad] (K-major) for the tcgen05 path; cin_pad = multiple of 64."""
        if self._wtc is None:
            T, cin, cp = self.w32.shape
            self.cin_pad = (cin + 63) // 64 * 64
            w = torch.zeros(T, cp, self.cin_pad, device=self.w32.device, dtype=torch.bfloat16)
            w[:, :, :cin] = self.w32.permute(0, 2, 1).to(torch.bfloat16)
            self._wtc = w.contiguous()
        return self._wtc

    @property
    def wtc_shuffle(self):
        """ConvTranspose2d(5, 2, 2, output_padding=1) with Cout <= 4 as ONE 3x3 conv over the input lattice whose 16
        output columns are (output parity, channel): bf16 [9 union taps][16][cin_pad].  Union tap (dy, dx) in
        {-1,0,1}^2 feeds parity (py, px) through filter tap (ky, kx) = (py + 2 - 2 dy, px + 2 - 2 dx) when that
        lies inside the 5x5 filter (same arithmetic as launches(): dy = (py + p - ky) / 2)."""
        if getattr(self, "_wtc_shuffle", None) is None:
            assert self.transposed and self.k == 5 and self.stride == 2 and self.pad == 2 and self.Cout <= 4
            T, cin, cp = self.w32.shape
            self.cin_pad = (cin + 63) // 64 * 64
            w = torch.zeros(9, 16, self.cin_pad, device=self.w32.device, dtype=torch.float32)
            for u, (dy, dx) in enumerate((a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)):
                for py in range(2):
                    for px in range(2):
                        ky, kx = py + 2 - 2 * dy, px + 2 - 2 * dx
                        if 0 <= ky < 5 and 0 <= kx < 5:
                            q = 2 * py + px
                            w[u, 4 * q:4 * q + self.Cout, :cin] = self.w32[ky * 5 + kx, :, :self.Cout].t()
            self._wtc_shuffle = w.to(torch.bfloat16).contiguous()
        return self._wtc_shuffle

    @property
    def wtc_rb3x3(self):
        """The 3x3 (96 -> 96) of a bottleneck block for rgbd_rb_*: bf16 [14 planes][96][64].  Plane t < 9 = tap t, input
        channels 0-63; plane 9 + i = input channels 64-95 of tap 2 i in elements [0, 32) and of tap 2 i + 1 in [32, 64)."""
        if getattr(self, "_wtc_rb", None) is None:
            assert not self.transposed and self.k == 3 and self.stride == 1 and self.pad == 1 and self.Cin == 96 and self.Cout == 96
            w = self.w32[:, :, :96].permute(0, 2, 1)         # [tap][co][ci]
            planes = torch.zeros(14, 96, 64, device=self.w32.device, dtype=torch.float32)
            planes[:9] = w[:, :, :64]
            for t in range(9):
                planes[9 + t // 2, :, 32 * (t % 2):32 * (t % 2) + 32] = w[t, :, 64:96]
            self._wtc_rb = planes.to(torch.bfloat16).contiguous()
        return self._wtc_rb

    def launches(self, H, W):
        """-> (Ho, Wo, [dict(Hs, Ws, o_step, o_off_y, o_off_x, i_step, taps=[(dy, dx, wtap)])])"""
        k, s, p = self.k, self.stride, self.pad
        if not self.transposed:
            Ho = (H + 2 * p - k) // s + 1
            Wo = (W + 2 * p - k) // s + 1
            taps = [(ky - p, kx - p, ky * k + kx) for ky in range(k) for kx in range(k)]
            return Ho, Wo, [dict(Hs=Ho, Ws=Wo, o_step=1, o_off_y=0, o_off_x=0, i_step=s, taps=taps)]
        if s == 1:  # ConvTranspose2d(k, 1, p): oy = iy - p + ky
            taps = [(p - ky, p - kx, ky * k + kx) for ky in range(k) for kx in range(k)]
            return H, W, [dict(Hs=H, Ws=W, o_step=1, o_off_y=0, o_off_x=0, i_step=1, taps=taps)]
        assert s == 2 and k == 5 and p == 2, "only ConvTranspose2d(5, 2, 2, output_padding=1) strided"
        out = []
        for py in range(2):
            for px in range(2):
                taps = [((py + p - ky) // 2, (px + p - kx) // 2, ky * k + kx)
                        for ky in range(k) if (py + p - ky) % 2 == 0
                        for kx in range(k) if (px + p - kx) % 2 == 0]
                out.append(dict(Hs=H, Ws=W, o_step=2, o_off_y=py, o_off_x=px, i_step=1, taps=taps))
        return 2 * H, 2 * W, out


class Program:
    """Static launch list + the buffers it owns."""

    def __init__(self, device):
        self.device = device
        self.ops = []       # callables taking the raw stream pointer
        self.keep = []      # ctypes structs / tensors that must outlive the ops
        self.pool = {}      # (shape, dtype) -> [free tensors]
        self.bytes = 0
        self.io = {}
        self.graph = None
        self.graph_launches = 0
        self.flops = 0      # dense conv flops of one run (MAC * 2)
        self.tc_plans = []  # rgbd_conv_tc_plan handles owned by this program
        self.rb_plans = []  # rgbd_rb_plan handles (fused bottleneck blocks)
        self.n_tc = 0
        self.n_simt = 0

    def __del__(self):
        try:
            lib = L.load()
            for h in self.tc_plans:
                lib.rgbd_conv_tc_plan_destroy(h)
            for h in self.rb_plans:
                lib.rgbd_rb_plan_destroy(h)
        except Exception:
            pass

    def run(self, use_graph=False):
        if use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
            L.load().rgbd_count_launch(self.graph_launches)   # the replayed kernel nodes
            return
        sp = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        for op in self.ops:
            op(sp)

    def _capture(self):
        # warm-up run on a side stream (lazy cudaFuncSetAttribute calls must not be captured)
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.run(False)
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        lib = L.load()
        before = lib.rgbd_launch_count(0)
        with torch.cuda.graph(g):
            sp = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            for op in self.ops:
                op(sp)
        self.graph_launches = int(lib.rgbd_launch_count(0) - before)   # kernel nodes in the graph
        lib.rgbd_count_launch(-self.graph_launches)                    # capture itself launched nothing
        self.graph = g


def _subtract(ranges, lo, hi):
    out = []
    for a, b in ranges:
        if b <= lo or a >= hi:
            out.append((a, b))
        else:
            if a < lo:
                out.append((a, lo))
            if b > hi:
                out.append((hi, b))
    return out


def _union(ranges, lo, hi):
    keep = []
    for a, b in ranges:
        if b < lo or a > hi:
            keep.append((a, b))
        else:
            lo, hi = min(lo, a), max(hi, b)
    return sorted(keep + [(lo, hi)])


def _missing(ranges, lo, hi):
    """Sub-ranges of [lo, hi) not covered by the sorted disjoint `ranges`."""
    out, p = [], lo
    for a, b in ranges:
        if b <= p:
            continue
        if a >= hi:
            break
        if a > p:
            out.append((p, a))
        p = max(p, b)
    if p < hi:
        out.append((p, hi))
    return out


class Builder:
    def __init__(self, device, act_dtype, tensor_cores=None):
        self.device = device
        self.act_dtype = act_dtype
        self.prog = Program(device)
        self._wscratch = None   # per-image SE-folded filters of the layer being run (shared by all layers)
        self._sched = None      # tile counters of the persistent conv kernels (one int32 per plan, zero between launches)
        self._sched_used = 0
        self._se_caches = {}    # id(buffer) -> table of SE partial sums kept across se_scale() calls
        self.stage = ""         # label of the part of the codec being compiled (g_a, h_a, h_s, chain, coder, g_s, io)
        # tcgen05 path: bf16 activations only
        self.tensor_cores = (act_dtype == torch.bfloat16) if tensor_cores is None else tensor_cores

    # ---- memory ----
    def alloc(self, N, H, W, Cc, dtype=None, zero=False):
        """NHWC buffer carved from the plan's byte arena (a small heap allocator: best-fit free ranges, split on
        allocation, neighbours coalesced on release).  The launch list is serial on one stream, so storage
        released at build time is dead by the time a later op runs; the high-resolution buffers of g_a / g_s
        thereby back the context-model and low-resolution tensors as well."""
        dtype = dtype or self.act_dtype
        # 16-byte aligned pixel rows (TMA global strides, vector epilogue stores)
        Cp = (Cc + 7) // 8 * 8 if dtype == torch.bfloat16 else Cc
        shape = (N, H, W, Cp)
        esz = 2 if dtype == torch.bfloat16 else 4
        nbytes = N * H * W * Cp * esz
        need = (nbytes + 1023) // 1024 * 1024          # 1 KB granules keep every carve aligned
        chunks = self.prog.pool.setdefault("chunks", [])   # [tensor, free ranges [(off, size)] sorted by off]
        pick = None
        if not zero:
            for ci, (_, free) in enumerate(chunks):
                for ri, (off, size) in enumerate(free):
                    if size >= need and (pick is None or size < pick[2]):
                        pick = (ci, ri, size)
        if pick is None:
            # new chunk: at least 64 MB per image, so that later tensors pack into its tail instead of each
            # opening a chunk of their own (zero-initialised buffers get an exact, private chunk)
            size = need if zero else max(need, (64 << 20) * N)
            store = (torch.zeros if zero else torch.empty)((size,), device=self.device, dtype=torch.uint8)
            self.prog.bytes += size
            chunks.append([store, [(need, size - need)] if size > need else []])
            ci, off = len(chunks) - 1, 0
        else:
            ci, ri, size = pick
            off = chunks[ci][1][ri][0]
            if size > need:
                chunks[ci][1][ri] = (off + need, size - need)
            else:
                del chunks[ci][1][ri]
        t = chunks[ci][0][off:off + nbytes].view(dtype).view(shape)
        t._rgbd_range = [ci, off, need]      # emptied on release (two views of one buffer release it once)
        return View(t, 0, Cc)

    def release(self, *views):
        chunks = self.prog.pool.get("chunks", [])
        for v in views:
            rng = getattr(v.buf, "_rgbd_range", None)
            if not rng:
                continue
            ci, off, size = rng
            del rng[:]
            free = chunks[ci][1]
            free.append((off, size))
            free.sort()
            merged = []
            for o, sz in free:
                if merged and merged[-1][0] + merged[-1][1] == o:
                    merged[-1] = (merged[-1][0], merged[-1][1] + sz)
                else:
                    merged.append((o, sz))
            chunks[ci][1] = merged

    def raw(self, shape, dtype, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, device=self.device, dtype=dtype)
        self.prog.bytes += t.numel() * t.element_size()
        self.prog.keep.append(t)
        return t

    def reserve_wscratch(self, elems):
        if self._wscratch is None or self._wscratch.numel() < elems:
            self._wscratch = self.raw((int(elems),), torch.bfloat16)
        return self._wscratch

    # ---- ops ----
    def op(self, name, *args):
        fn = getattr(L.load(), name)
        label = name

        def run(sp, fn=fn, args=args, name=name):
            rc = fn(*args, sp)
            if rc:
                L.check(rc, name)
        run.label = label
        run.stage = self.stage
        self.prog.ops.append(run)
        return run

    def torch_op(self, f):
        self.prog.ops.append(lambda sp, f=f: f())

    def conv(self, pc, x, out=None, act=L.ACT_NONE, epi=L.EPI_LINEAR, res=None, mul=None, in_scale=None,
             y2=None, out_dtype=None):
        assert x.C == pc.Cin, (x.C, pc.Cin)
        Ho, Wo, launches = pc.launches(x.H, x.W)
        if out is None:
            out = self.alloc(x.N, Ho, Wo, pc.Cout, out_dtype)
        assert (out.N, out.H, out.W, out.C) == (x.N, Ho, Wo, pc.Cout), ((out.N, out.H, out.W, out.C), (x.N, Ho, Wo, pc.Cout))
        # small Cin (the 3 / 1-channel image layers) still goes to the tensor cores: TMA zero-fills the
        # channel block beyond Cin, so a tap costs one K=16 MMA
        use_tc = self.tensor_cores and x.dtype == torch.bfloat16 and x.cstride % 8 == 0 and x.coff % 8 == 0
        scaled = None
        w_folded = None
        if use_tc and in_scale is not None:
            T, _, cp = pc.w32.shape
            w_elems = T * cp * ((pc.Cin + 63) // 64 * 64)
            if w_elems < x.H * x.W * x.C:
                # the gate is per (image, input channel): fold it into per-image copies of the filter, which is
                # smaller than the activation tensor (the 1x1 of EntropyParametersEX: <= 2816 x 480 vs 1280 px)
                src = pc.wtc32
                w_folded = self.reserve_wscratch(x.N * w_elems)
                self.op("rgbd_scale_weights", src.data_ptr(), in_scale.data_ptr(), w_folded.data_ptr(), x.N, T, cp,
                        pc.cin_pad, pc.Cin)
                self.prog.keep.append(src)
                in_scale = None
        if use_tc and in_scale is not None:
            # the TMA-fed A operand never passes through registers: apply the SE gate in a separate pass
            scaled = self.alloc(x.N, x.H, x.W, x.C, x.dtype)
            self.op("rgbd_scale_channels", x.ptr(), scaled.ptr(), _DT[x.dtype], in_scale.data_ptr(), x.N, x.H * x.W,
                    x.C, x.cstride, x.coff, scaled.cstride, scaled.coff)
            x, in_scale = scaled, None
        shuffle = (use_tc and pc.transposed and pc.stride == 2 and pc.Cout <= 4 and res is None and mul is None
                   and y2 is None and in_scale is None and epi == L.EPI_LINEAR)
        if shuffle:
            # the 4 output-parity launches of a transposed conv with <= 4 output channels fused into one 3x3 launch
            # (N = 16 columns = parity x channel, pixel-shuffle epilogue): the input is read once instead of 4 times
            launches = [dict(Hs=x.H, Ws=x.W, o_step=2, o_off_y=0, o_off_x=0, i_step=1, shuffle=True,
                             taps=[(dy, dx, u) for u, (dy, dx) in enumerate((a, b_) for a in (-1, 0, 1) for b_ in (-1, 0, 1))])]
        for ln in launches:
            d = L.ConvDesc()
            d.x, d.y, d.w = x.ptr(), out.ptr(), pc.w32.data_ptr()
            d.y2 = y2.ptr() if y2 is not None else None
            d.bias = pc.bias.data_ptr() if pc.bias is not None else None
            d.in_scale = in_scale.data_ptr() if in_scale is not None else None
            d.N, d.H, d.W = x.N, x.H, x.W
            d.Cin, d.x_cstride, d.x_coff = x.C, x.cstride, x.coff
            d.Ho, d.Wo = Ho, Wo
            d.Cout, d.y_cstride, d.y_coff = pc.Cout, out.cstride, out.coff
            if y2 is not None:
                assert y2.dtype == out.dtype and (y2.N, y2.H, y2.W, y2.C) == (out.N, out.H, out.W, out.C)
                d.y2_cstride, d.y2_coff = y2.cstride, y2.coff
            for f in ("Hs", "Ws", "o_step", "o_off_y", "o_off_x", "i_step"):
                setattr(d, f, ln[f])
            d.ntaps = len(ln["taps"])
            for i, (dy, dx, wt) in enumerate(ln["taps"]):
                d.dy[i], d.dx[i], d.wtap[i] = dy, dx, wt
            if res is not None:
                assert res.dtype == x.dtype
                d.res, d.res_cstride, d.res_coff = res.ptr(), res.cstride, res.coff
                if epi == L.EPI_BILERP:
                    d.res_H, d.res_W = res.H, res.W
                else:
                    assert (res.H, res.W, res.C) == (Ho, Wo, pc.Cout)
            if mul is not None:
                assert mul.dtype == x.dtype and (mul.H, mul.W, mul.C) == (Ho, Wo, pc.Cout)
                d.mul, d.mul_cstride, d.mul_coff = mul.ptr(), mul.cstride, mul.coff
            d.act, d.epi = act, (L.EPI_SHUFFLE2 if ln.get("shuffle") else epi)
            d.x_dtype, d.y_dtype = _DT[x.dtype], _DT[out.dtype]
            d.cout_pad = 16 if ln.get("shuffle") else pc.cout_pad
            self.prog.keep.append(d)
            if use_tc:
                d.sched_ws = self._sched_slot()
                d.w = pc.wtc.data_ptr()
                if w_folded is not None:
                    d.w = w_folded.data_ptr()
                    d.w_image_stride = pc.w32.shape[0]
                if ln.get("shuffle"):
                    d.w = pc.wtc_shuffle.data_ptr()
                handle = C.c_void_p()
                L.check(L.load().rgbd_conv_tc_plan_create(C.byref(d), pc.cin_pad, C.byref(handle)),
                        "rgbd_conv_tc_plan_create")
                self.prog.tc_plans.append(handle)
                run = self.op("rgbd_conv_tc_run", handle)
                self.prog.n_tc += 1
            else:
                L.check(L.load().rgbd_conv_validate(C.byref(d)), "rgbd_conv_validate")
                run = self.op("rgbd_conv_simt", C.byref(d))
                self.prog.n_simt += 1
            run.is_conv = True
            run.is_tc = use_tc
            kind = ("deconv" if pc.transposed else "conv") + f"{pc.k}x{pc.k}" + (f"s{pc.stride}" if pc.stride > 1 else "")
            run.label = f"{'tc' if use_tc else 'simt'} {kind} {pc.Cin}->{pc.Cout} @{ln['Hs']}x{ln['Ws']}"
            # dense algorithmic count; the fused transposed conv still counts its 25 filter taps
            ntap_alg = 25 if ln.get("shuffle") else (pc.flop_taps or len(ln["taps"]))
            run.flops = 2 * x.N * ln["Hs"] * ln["Ws"] * ntap_alg * pc.flop_cin * pc.Cout
            self.prog.flops += run.flops
            # algorithmic HBM bytes of this launch: the input once (shared by the parity launches of a transposed
            # conv), the output lattice it writes (+ second copy), residual / gate operands, the weights
            esz, osz = x.buf.element_size(), out.buf.element_size()
            opix = x.N * ln["Hs"] * ln["Ws"] * (4 if ln.get("shuffle") else 1)
            run.bytes = (x.N * x.H * x.W * x.C * esz // len(launches) + opix * pc.Cout * osz * (2 if y2 is not None else 1)
                         + (opix * pc.Cout * esz if res is not None and epi != L.EPI_BILERP else 0)
                         + (opix * pc.Cout * esz if mul is not None else 0)
                         + len(ln["taps"]) * pc.Cin * pc.Cout * 2)
        self.prog.keep.extend([pc, x.buf, out.buf])
        self._wrote(out)
        if y2 is not None:
            self._wrote(y2)
        if scaled is not None:
            self.release(scaled)
        return out

    def can_fuse_block(self, pc1, pc2, pc3, x):
        """1x1 (Cin -> 96) / 3x3 (96 -> 96) / 1x1 (96 -> 192) on a bf16 tensor-core plan: the shapes rgbd_rb_* takes."""
        return (self.tensor_cores and x.dtype == torch.bfloat16 and not pc1.transposed and pc1.k == 1 and pc3.k == 1
                and pc2.k == 3 and pc2.stride == 1 and pc2.pad == 1 and pc1.Cout == 96 and pc2.Cin == 96 and pc2.Cout == 96
                and pc3.Cin == 96 and pc3.Cout == 192 and pc1.Cin % 64 == 0 and x.cstride % 8 == 0 and x.coff % 8 == 0
                and x.H >= 2 and x.W >= 2)

    def fused_block(self, pc1, pc2, pc3, x, res, out=None, final_relu=False):
        """y = act(res + conv1x1(relu(conv3x3(relu(conv1x1(x))))) in one launch (csrc/conv_rb.cu): the bottleneck's two
        96-channel intermediates never reach HBM."""
        assert x.C == pc1.Cin and (res.N, res.H, res.W, res.C) == (x.N, x.H, x.W, pc3.Cout) and res.dtype == x.dtype
        if out is None:
            out = self.alloc(x.N, x.H, x.W, pc3.Cout, x.dtype)
        assert (out.N, out.H, out.W, out.C) == (x.N, x.H, x.W, pc3.Cout) and out.dtype == x.dtype
        d = L.RbDesc()
        w1, w2, w3 = pc1.wtc, pc2.wtc_rb3x3, pc3.wtc
        assert pc1.cin_pad == pc1.Cin and pc3.cin_pad == 128
        d.x, d.res, d.y = x.ptr(), res.ptr(), out.ptr()
        d.w1, d.w2, d.w3 = w1.data_ptr(), w2.data_ptr(), w3.data_ptr()
        for name, pc in (("b1", pc1), ("b2", pc2), ("b3", pc3)):
            setattr(d, name, pc.bias.data_ptr() if pc.bias is not None else None)
        d.N, d.H, d.W = x.N, x.H, x.W
        d.Cin, d.x_cstride, d.x_coff = x.C, x.cstride, x.coff
        d.Cmid, d.Cout = 96, pc3.Cout
        d.res_cstride, d.res_coff, d.y_cstride, d.y_coff = res.cstride, res.coff, out.cstride, out.coff
        d.final_relu = int(bool(final_relu))
        d.sched_ws = self._sched_slot()
        handle = C.c_void_p()
        L.check(L.load().rgbd_rb_plan_create(C.byref(d), C.byref(handle)), "rgbd_rb_plan_create")
        self.prog.rb_plans.append(handle)
        run = self.op("rgbd_rb_run", handle)
        self.prog.n_tc += 1
        run.is_conv = True
        run.is_tc = True
        run.label = f"tc fused 1x1-3x3-1x1 {x.C}->96->96->{pc3.Cout} @{x.H}x{x.W}"
        px = x.N * x.H * x.W
        run.flops = 2 * px * (x.C * 96 + 9 * 96 * 96 + 96 * pc3.Cout)
        self.prog.flops += run.flops
        run.bytes = px * (x.C + 2 * pc3.Cout) * 2 + (x.C * 96 + 9 * 96 * 96 + 96 * pc3.Cout) * 2
        self.prog.keep.extend([d, pc1, pc2, pc3, w1, w2, w3, x.buf, res.buf, out.buf])
        self._wrote(out)
        return out

    def se_scale(self, x, w1, w2, plus_one):
        """SE_Block channel gate of `x` -> fp32 [N, C] scale tensor.  For a buffer registered with se_cache() the table of
        per-channel partial sums persists between calls and only channels written since (tracked by conv / fused_block)
        are summed again — the same bits as a full recomputation, the sums being per channel and in a fixed order."""
        Cr = w1.shape[0]
        HW = x.H * x.W
        nchunk = max(1, min(64, HW // 64))
        scale = self.raw((x.N, x.C), torch.float32)
        cache = self._se_caches.get(id(x.buf))
        self.prog.keep.extend([w1, w2])
        if cache is None:
            partial = self.raw((x.N * (nchunk * x.C + x.C + Cr),), torch.float32)
            self.op("rgbd_se_scale", x.ptr(), _DT[x.dtype], x.N, HW, x.C, x.cstride, x.coff, w1.data_ptr(),
                    w2.data_ptr(), Cr, int(plus_one), partial.data_ptr(), nchunk, scale.data_ptr())
            return scale
        assert x.coff == 0, "cached SE sums are indexed by the buffer's own channels"
        pstride = x.cstride
        if "partial" not in cache:
            cache["partial"] = self.raw((x.N * nchunk * pstride,), torch.float32)
            cache["work"] = self.raw((x.N * (pstride + 4096),), torch.float32)
            cache["valid"] = []            # disjoint sorted [lo, hi) channel ranges whose sums are current
        assert Cr <= 4096
        for lo, hi in _missing(cache["valid"], 0, x.C):
            self.op("rgbd_se_partial", x.ptr(), _DT[x.dtype], x.N, HW, x.cstride, 0, lo, hi, nchunk,
                    cache["partial"].data_ptr(), pstride)
        cache["valid"] = _union(cache["valid"], 0, x.C)
        self.op("rgbd_se_gate", cache["partial"].data_ptr(), pstride, nchunk, x.N, HW, x.C, w1.data_ptr(), w2.data_ptr(),
                Cr, int(plus_one), cache["work"].data_ptr(), scale.data_ptr())
        return scale

    def _sched_slot(self):
        """Address of a fresh zero-initialised int32: the tile counter of one persistent-kernel plan."""
        if self._sched is None or self._sched_used == self._sched.numel():
            self._sched = self.raw((4096,), torch.int32, zero=True)
            self._sched_used = 0
        ptr = self._sched.data_ptr() + 4 * self._sched_used
        self._sched_used += 1
        return ptr

    def se_cache(self, view):
        """Keep the SE partial sums of this buffer between se_scale() calls (see there)."""
        self._se_caches.setdefault(id(view.buf), {})

    def _wrote(self, view):
        """A launch writes channels [coff, coff + C) of view.buf: their cached SE sums are stale."""
        cache = self._se_caches.get(id(view.buf))
        if cache and "valid" in cache:
            cache["valid"] = _subtract(cache["valid"], view.coff, view.coff + view.C)

    def maxpool7s3(self, x):
        # dense buffers; a bf16 buffer whose channel count was padded to a multiple of 8 (f = 12 in STF_united's narrowest
        # ESA) is pooled over its padded width — the padding lanes are never read as channels by anyone
        out = self.alloc(x.N, (x.H - 7) // 3 + 1, (x.W - 7) // 3 + 1, x.C, x.dtype)
        assert x.coff == 0 and out.cstride == x.cstride and x.cstride - x.C < 8
        self.op("rgbd_maxpool7s3", x.ptr(), out.ptr(), _DT[x.dtype], x.N, x.H, x.W, x.cstride)
        return out
