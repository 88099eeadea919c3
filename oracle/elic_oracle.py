"""ORACLE (test infrastructure): torch-CPU fp32 restatement of the reference's single-modality ELIC
(models/elic.py:15-351) as functions over a state_dict; primitives (conv, bottleneck, attention, checkerboard
squeeze, indexes, likelihoods, coder) are the ones of oracle/model_oracle.py.

What it follows:  g_a / g_s  modules/transform/analysis.py:29-52, synthesis.py:32-51;  h_a / h_s  analysis.py:207-217,
synthesis.py:276-285;  EntropyParameters  modules/transform/entropy.py:7-28;  per-slice order and concat orders
models/elic.py:73-156 (forward), :192-241 (compress), :264-315 (decompress).
Parity pin: tests/test_oracle_elic.py against tests/golden/model_elic_*.npz, produced by the unmodified reference
(oracle/make_golden_elic.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from .model_oracle import OracleCodec


class ElicOracle(OracleCodec):
    _GA1 = ["c", "rb", "rb", "rb", "c", "rb", "rb", "rb", "at", "c", "rb", "rb", "rb", "c", "at"]
    _GS1 = ["at", "dc", "rb", "rb", "rb", "dc", "at", "rb", "rb", "rb", "dc", "rb", "rb", "rb", "dc"]

    def __init__(self, state_dict, **kw):
        sd = dict(state_dict)
        for k, v in state_dict.items():          # the shared helpers address the entropy models as "<which>_<model>."
            if k.startswith("gaussian_conditional.") or k.startswith("entropy_bottleneck."):
                sd["x_" + k] = v
        super().__init__(sd, **kw)

    def _seq(self, prefix, kinds, x):
        f = {"c": lambda p, t: self._c(p, t, stride=2, pad=2), "dc": self._d, "rb": self._rb, "at": self._attn}
        for i, kind in enumerate(kinds):
            x = f[kind](f"{prefix}.{i}", x)
        return x

    def g_a1(self, x):
        return self._seq("g_a.analysis_transform", self._GA1, x)

    def g_s1(self, y):
        return self._seq("g_s.synthesis_transform", self._GS1, y)

    def h_a1(self, y):
        t = F.relu(self._c("h_a.reduction.0", y, pad=1))
        t = F.relu(self._c("h_a.reduction.2", t, stride=2, pad=2))
        return self._c("h_a.reduction.4", t, stride=2, pad=2)

    def h_s1(self, z):
        t = F.relu(self._d("h_s.increase.0", z))
        t = F.relu(self._d("h_s.increase.2", t))
        return self._d("h_s.increase.4", t, k=3, s=1)

    def _ep1(self, p, x):
        t = F.relu(self._c(p + ".fusion.0", x))
        t = F.relu(self._c(p + ".fusion.2", t))
        return self._c(p + ".fusion.4", t).chunk(2, 1)

    def _chain1(self, hyper, step):
        yh = []
        for idx in range(len(self.slice_ch)):
            base = [hyper]
            if idx:
                base = [self._chctx(f"channel_context.{idx}", torch.cat(yh, 1))] + base      # [channel_ctx, hyper]
            s, m = self._ep1(f"entropy_parameters_anchor.{idx}", torch.cat(base, 1))
            a = step(idx, 0, s, m)
            loc = self._c(f"local_context.{idx}", a, pad=2)
            s, m = self._ep1(f"entropy_parameters_nonanchor.{idx}", torch.cat([loc] + base, 1))
            n = step(idx, 1, s, m)
            yh.append(a + n)
        return torch.cat(yh, 1)

    @torch.no_grad()
    def forward(self, x):
        y = self.g_a1(x)
        z = self.h_a1(y)
        z_hat, lz = self.eb_forward("x", z)
        hyper = self.h_s1(z_hat)
        lik = torch.zeros_like(y)
        H, W = y.shape[2:]

        def step(idx, parity, scales, means):
            m = self._mask(H, W, parity)
            ys = self._slice(y, idx)
            d = (ys - means) * m
            a = sum(self.slice_ch[:idx])
            lik[:, a:a + self.slice_ch[idx]] += self._gauss_likelihood("x", ys, scales, means) * m
            return (torch.round(d) - d + d) + means * m

        y_hat = self._chain1(hyper, step)
        return {"x_hat": self.g_s1(y_hat), "likelihoods": {"y_likelihoods": lik, "z_likelihoods": lz},
                "_trace": {"y": y, "z": z, "hyper": hyper, "yhat": y_hat}}

    @torch.no_grad()
    def compress(self, x, trace=False):
        y = self.g_a1(x)
        z = self.h_a1(y)
        z_strings, z_hat, zsym = self._z_code("x", z)
        hyper = self.h_s1(z_hat)
        B = x.shape[0]
        syms, idxs = [[] for _ in range(B)], [[] for _ in range(B)]

        def step(idx, parity, scales, means):
            ys = self._squeeze(self._slice(y, idx), parity)
            ss, mm = self._squeeze(scales, parity), self._squeeze(means, parity)
            ind = self._indexes("x", ss)
            sym = torch.round(ys - mm).int()
            for i in range(B):
                syms[i].append(sym[i].reshape(-1).numpy())
                idxs[i].append(ind[i].reshape(-1).numpy())
            return self._unsqueeze(sym.float() + mm, parity)

        y_hat = self._chain1(hyper, step)
        t = self.gc_tables("x")
        flat = [(np.concatenate(syms[i]), np.concatenate(idxs[i])) for i in range(B)]
        out = {"strings": [[self.coder.encode_with_indexes(s, i_, t) for s, i_ in flat], z_strings],
               "shape": tuple(z.shape[-2:])}
        if trace:
            out["_trace"] = {"y": y, "z": z, "zsym": zsym, "hyper": hyper, "yhat": y_hat, "symbols": flat}
        return out

    @torch.no_grad()
    def decompress(self, strings, shape):
        B = len(strings[1])
        hz, wz = int(shape[0]), int(shape[1])
        t = self.eb_tables("x")
        C = self.sd["entropy_bottleneck._quantized_cdf"].shape[0]
        med = self.sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        idx = np.repeat(np.arange(C, dtype=np.int32), hz * wz)
        vals = [self.coder.decode_with_indexes(strings[1][i], idx, t).reshape(C, hz, wz) for i in range(B)]
        hyper = self.h_s1(torch.from_numpy(np.stack(vals)).float() + med)
        dec = [self.coder.Decoder(strings[0][i]) for i in range(B)]
        tab = self.gc_tables("x")

        def step(idx, parity, scales, means):
            ss, mm = self._squeeze(scales, parity), self._squeeze(means, parity)
            ind = self._indexes("x", ss)
            v = [dec[i].decode_stream(ind[i].reshape(-1).numpy(), tab).reshape(ind[i].shape) for i in range(B)]
            return self._unsqueeze(torch.from_numpy(np.stack(v)).float() + mm, parity)

        y_hat = self._chain1(hyper, step)
        return {"x_hat": self.g_s1(y_hat), "_trace": {"yhat": y_hat}}     # (the reference does not clamp here, :317-330)
