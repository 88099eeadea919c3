#!/usr/bin/env python
"""Throughput bench of the ELIC_united compress+decompress hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the CPU implementation of the path

A step = compress + decompress of one batch of synthetic NYUv2-shaped pairs (480x640 padded to
512x640, calibrated random-init weights: rgbd_b200.synthetic).  `value` has the inputs resident
in HBM; `e2e` goes through the public API with pinned HOST buffers (H2D of the images and D2H of
the reconstruction inside the timed region).  The rANS strings are host `bytes` in both, because
that is the codec's API contract (they are the compressed file).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RGB-D pairs/sec compress+decompress 480x640"
UNIT = "pairs/s"
GFLOP_PER_PAIR = 721.2 + 785.1   # dense conv count, SURVEY §8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU per step (default: by precision)")
    ap.add_argument("--precision", default=os.environ.get("RGBD_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--slots", type=int, default=8, help="batches in flight per GPU (own program + CUDA stream each)")
    ap.add_argument("--graphs", type=int, default=1, help="replay each slot's launch list as a CUDA graph")
    ap.add_argument("--threads", type=int, default=0, help="drive every pipeline slot from its own host thread")
    ap.add_argument("--dec-slots", type=int, default=0, help="decompress jobs in flight (default: = --slots)")
    ap.add_argument("--hiprio", type=int, default=1, help="decoder slots on high-priority CUDA streams")
    ap.add_argument("--preset", default="realistic")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--cpu-pairs", type=int, default=2, help="pairs in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
def make_inputs(n, H, W, seed):
    from rgbd_b200.synthetic import pad_to_multiple, synthetic_pairs
    rgb, depth = synthetic_pairs(n, H, W, seed=seed)
    return pad_to_multiple(rgb), pad_to_multiple(depth)


def cpu_arm(args, pairs, steps, warmup):
    """The path's CPU implementation timed on this box's host cores: oracle/model_oracle.py (torch
    CPU fp32 restatement, pinned bit-exactly to the reference) + the reference's own compiled rANS
    coder from oracle/_ref when it was built (else the C restatement).  Batch 1 loop like
    testing/tester_united.py, no file I/O."""
    import torch
    import rgbd_b200
    from oracle.model_oracle import OracleCodec
    from oracle.ref_loader import ref_ext_available
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4).eval()
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, args.preset))
    net.update(force=True)
    use_ref = ref_ext_available()
    orc = OracleCodec(net.state_dict(), use_ref_coder=use_ref)
    rgb, depth = make_inputs(pairs, args.height, args.width, seed=1234)
    times = []
    sample = []      # per pair: stream bytes and PSNR of the CPU path (the parity reference of the GPU line)
    H, W = args.height, args.width
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        outs = []
        for i in range(pairs):
            c = orc.compress(rgb[i:i + 1], depth[i:i + 1])
            outs.append((c, orc.decompress(c["r_strings"], c["d_strings"], c["shape"])))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
        if not sample:
            for i, (c, r) in enumerate(outs):
                row = {}
                for key, m, x in (("r_strings", "r", rgb), ("d_strings", "d", depth)):
                    row["bytes_" + m] = sum(len(s_) for grp in c[key] for s_ in grp)
                    mse = float(((r["x_hat"][m][:, :, :H, :W].double() - x[i:i + 1, :, :H, :W].double()) ** 2).mean())
                    row["psnr_" + m] = 99.0 if mse <= 0 else 10 * math.log10(1.0 / mse)
                sample.append(row)
    total = sum(times)
    return {"pairs_coded": sample, "value": pairs * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "coder": "reference ans (oracle/_ref)" if use_ref else "C restatement",
            "sample": f"{pairs} pair(s) x {len(times)} timed pass(es) of {args.height}x{args.width} "
                      f"compress+decompress, batch 1, torch CPU fp32 ({warmup} warm-up)",
            "ms_per_pair": 1e3 * total / (pairs * len(times))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    cb = cpu_arm(args, 1, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_pair"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ELIC_united compress+decompress, {args.height}x{args.width} -> 512x640 "
                                   f"pairs, batch 1 per step, preset {args.preset}", "host": "cpu"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "coder")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    import rgbd_b200
    from rgbd_b200 import lib as L
    from rgbd_b200.parallel import add_pair_stats, allreduce_stats, new_stats, summarize

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4, precision=args.precision).eval()
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, args.preset))
    net.update(force=True)
    net = net.to(dev)
    net.use_cuda_graph = bool(args.graphs)

    def prepare(B):
        """Inputs, pinned host buffers, the pipeline and its 2S launch plans for B pairs per job (warm-up included:
        building the plans is what allocates the HBM)."""
        # each rank owns its contiguous shard of the global batch (weak scaling: S x B pairs per GPU and
        # step); S batches are in flight on S CUDA streams so that the serial rANS kernels of one batch
        # overlap the convolutions of the others
        S = max(1, args.slots)
        rgb_h, depth_h = make_inputs(S * B, args.height, args.width, seed=1234 + rank * S * B)
        rgb_h, depth_h = rgb_h.pin_memory(), depth_h.pin_memory()
        rgb_d, depth_d = rgb_h.to(dev), depth_h.to(dev)
        Hp, Wp = rgb_h.shape[-2:]
        sl = [slice(i * B, (i + 1) * B) for i in range(S)]
        host_out = [(torch.empty((B, 3, Hp, Wp)).pin_memory(), torch.empty((B, 1, Hp, Wp)).pin_memory()) for _ in range(max(S, args.dec_slots or S))]

        # Round-trip pipeline (rgbd_b200.pipeline): S compress jobs and S decompress jobs in flight, each on
        # its own stream + launch plan, so the decoder's serial rANS chain of batch k hides behind the
        # convolutions of batches k+1..; one step = S batches of B pairs, every pair compressed AND
        # decompressed inside the timed region.
        from rgbd_b200.pipeline import RoundTripPipeline
        pipe = RoundTripPipeline(net, S, threads=bool(args.threads), high_priority_decode=bool(args.hiprio), dec_slots=args.dec_slots or None)

        def steps_device(k):
            jobs = [(rgb_d[sl[i % S]], depth_d[sl[i % S]]) for i in range(k * S)]
            return pipe.run(jobs)

        def steps_e2e(k):
            def stage_input(j, slot, stream):    # H2D of this batch's images from pinned host memory
                return rgb_h[sl[slot]].to(dev, non_blocking=True), depth_h[sl[slot]].to(dev, non_blocking=True)

            def sink(j, slot, stream, x_r, x_d):  # D2H of the reconstruction into pinned host buffers
                host_out[slot][0].copy_(x_r, non_blocking=True)
                host_out[slot][1].copy_(x_d, non_blocking=True)

            res = pipe.run([None] * (k * S), stage_input=stage_input, sink=sink)
            torch.cuda.synchronize(dev)
            return res
        W_ = max(3, args.warmup)
        if os.environ.get("RGBD_BENCH_FAKE_OOM") and B == int(os.environ["RGBD_BENCH_FAKE_OOM"]):
            raise torch.OutOfMemoryError("fake OOM (test hook)")
        steps_device(W_)
        steps_e2e(1)
        import types
        return types.SimpleNamespace(**{k: v for k, v in locals().items() if k != "types"})

    # pairs per job: as many as fit.  Plans for 8 + 8 jobs of 24 pairs take ~107 GB of the 180 GB; if building them
    # runs out of memory on this GPU (every rank must agree), fall back to a smaller job instead of failing the run.
    candidates = [args.batch] if args.batch else ([2] if args.precision == "fp32" else [24, 16, 8])
    ns = None
    for B in candidates:
        ok = 1
        try:
            ns = prepare(B)
        except torch.OutOfMemoryError as e:
            ok, ns = 0, None
            print(f"[bench] rank {rank}: {B} pairs per job do not fit ({str(e)[:80]}...)", file=sys.stderr, flush=True)
        if world > 1:
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            break
        ns = None
        net._invalidate()
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if ns is None:
        raise SystemExit("bench: no job size fits this GPU")
    S, rgb_h, depth_h, rgb_d, depth_d, Hp, Wp, sl, host_out = (ns.S, ns.rgb_h, ns.depth_h, ns.rgb_d, ns.depth_d, ns.Hp,
                                                               ns.Wp, ns.sl, ns.host_out)
    pipe, steps_device, steps_e2e = ns.pipe, ns.steps_device, ns.steps_e2e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(steps)     # returns once every job of the K steps has completed (all slots drained)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    W_, K = max(3, args.warmup), max(1, args.steps)   # (the W_ warm-up steps ran inside prepare())
    # L2 note: one step streams > 1 GB of activations per image through HBM, far beyond the 126 MB L2,
    # so consecutive steps cannot serve each other from cache (no explicit flush needed).
    L.load().rgbd_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res = timed(steps_device, K)
    cs = [c for _, c, _ in res]
    outs = [xs for _, _, xs in res]
    last_slots = [j % S for j, _, _ in res]
    launches = int(L.load().rgbd_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    pairs_per_step = S * B
    value = world * pairs_per_step * K / (ms / 1e3)
    stats = new_stats()
    for i, slot in enumerate(last_slots):
        add_pair_stats(stats, cs[i]["r_strings"], cs[i]["d_strings"],
                       rgb_d[sl[slot]][:, :, :args.height, :args.width], depth_d[sl[slot]][:, :, :args.height, :args.width],
                       outs[i][0][:, :, :args.height, :args.width], outs[i][1][:, :, :args.height, :args.width])

    ms_e2e, res2 = timed(steps_e2e, K)
    cs2 = [c for _, c, _ in res2]
    outs2 = [xs for _, _, xs in res2]
    e2e_value = world * pairs_per_step * K / (ms_e2e / 1e3)
    stream_bytes = sum(len(x) for c in cs2 for key in ("r_strings", "d_strings") for grp in c[key] for x in grp)
    h2d = rgb_h.numel() * 4 + depth_h.numel() * 4 + stream_bytes
    d2h = stream_bytes + sum(a.numel() * 4 + b_.numel() * 4 for a, b_ in outs2)

    # roofline of the dominant kernel family (the implicit-GEMM conv): per-launch CUDA events on the
    # launching stream, over the same workload
    roof = conv_roofline(net, B, Hp, Wp, dev, dec_slot=S)
    stats = summarize(allreduce_stats(stats, dev))
    if rank == 0:
        pk, pk_src = peaks()
        tens_peak = pk["bf16_tflops_sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"ELIC_united compress+decompress, {world}x{S}x{B} pairs/step of "
                                   f"{args.height}x{args.width} (padded {Hp}x{Wp}), preset {args.preset}, "
                                   f"weights calibrated random-init", "pairs_per_gpu": S * B, "batch": B, "slots_in_flight": S, "schedule": "pipelined: S compress + S decompress jobs in flight (rgbd_b200.pipeline)", "host_threads": bool(args.threads), "precision": args.precision, "cuda_graphs": bool(args.graphs),
                       "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)",
                       "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1),
                       "parallelism": f"dp{world} (images sharded, no data-path collective)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": roof["tflops"], "peak": tens_peak, "unit": "TFLOP/s",
                         "frac": roof["tflops"] / tens_peak, "traffic": None,
                         "kernel": roof["kernel"], "launches": roof["launches"], "peak_source": pk_src + " bf16 sustained",
                         "share_of_step": S * roof["ms"] / (ms / K), "algorithmic_gflop_per_pair": GFLOP_PER_PAIR, "by_class": roof["by_class"],
                         "timing": "CUDA events around every run of consecutive conv launches of one compress + decompress plan",
                         "achieved_with_per_launch_events": roof["per_launch_events_tflops"],
                         "whole_step_tflops": GFLOP_PER_PAIR * pairs_per_step * K / ms},
            "quality": stats,
        }
        if not args.no_cpu_baseline:
            cb = cpu_arm(args, args.cpu_pairs, 1, 0)
            coded = cb.pop("pairs_coded")
            line["cpu_baseline"] = cb
            # correctness gate next to the throughput number: the same pairs through the CPU path (fp32
            # restatement of the reference) and through this GPU run — bpp deviation and PSNR deviation per modality
            first = [e for e in res2 if e[0] % S == 0]     # slot 0 = pairs 0..B-1 of this rank (most recent run)
            if first and rank == 0:
                _, c0, (xr, xd) = first[-1]
                npx = args.height * args.width
                dev_bpp, dev_psnr = [], []
                for i, row in enumerate(coded[:B]):
                    for key, m, x, xh in (("r_strings", "r", rgb_d, xr), ("d_strings", "d", depth_d, xd)):
                        nbytes = sum(len(grp[i]) for grp in c0[key])
                        dev_bpp.append(100.0 * abs(nbytes - row["bytes_" + m]) / row["bytes_" + m])
                        mse = float(((xh[i:i + 1, :, :args.height, :args.width].double() -
                                      x[i:i + 1, :, :args.height, :args.width].double()) ** 2).mean())
                        psnr = 99.0 if mse <= 0 else 10 * math.log10(1.0 / mse)
                        dev_psnr.append(abs(psnr - row["psnr_" + m]))
                line["parity_vs_cpu_path"] = {"pairs": len(coded[:B]), "max_bpp_dev_pct": round(max(dev_bpp), 4),
                                              "max_psnr_dev_db": round(max(dev_psnr), 5),
                                              "tolerance": "bpp 0.5 %, PSNR 0.05 dB (BASELINE north_star, bf16 mode)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(net, B, Hp, Wp, dev, dec_slot=0):
    """Algorithmic conv flops / sum of conv launch durations for one encoder + decoder pass, plus the same split
    by which roof bounds each launch (arithmetic intensity above / below the ridge of the measured peaks)."""
    import ctypes as C
    import torch
    pk, _ = peaks()
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)     # flop per byte
    total_ms, total_flops, n = 0.0, 0.0, 0
    cls = {"tensor": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}       # ms, flops, bytes, launches
    with torch.cuda.device(dev):
        for prog in (net._program("encoder", B, Hp, Wp),
                     net._program("decoder", B, Hp // 64, Wp // 64, slot=dec_slot)):   # plans the timed run used: their buffers hold real streams
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            evs = []
            for op in prog.ops:
                if getattr(op, "is_conv", False):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    op(sp)
                    b.record()
                    evs.append((a, b, op))
                else:
                    op(sp)
            torch.cuda.synchronize(dev)
            for a, b, op in evs:
                ms = a.elapsed_time(b)
                total_ms += ms
                c = cls["tensor" if op.flops / max(1, op.bytes) >= ridge else "hbm"]
                c[0] += ms
                c[1] += op.flops
                c[2] += op.bytes
                c[3] += 1
            total_flops += prog.flops
            n += len(evs)
    # Second pass: one event pair around every maximal RUN of consecutive conv launches instead of around every launch.
    # Per-launch events put ~10 us of bubble around each of the ~640 launches (the kernels measure 10+ us shorter under
    # ncu) and forbid the programmatic-dependent-launch overlap the real run has; per-run events keep launch gaps and
    # PDL exactly as in production.  `achieved` uses this pass; the per-launch pass feeds the by-class split.
    run_ms = 0.0
    with torch.cuda.device(dev):
        for prog in (net._program("encoder", B, Hp, Wp),
                     net._program("decoder", B, Hp // 64, Wp // 64, slot=dec_slot)):
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            evs, open_ev = [], None
            for op in prog.ops:
                is_conv = getattr(op, "is_conv", False)
                if is_conv and open_ev is None:
                    open_ev = torch.cuda.Event(enable_timing=True)
                    open_ev.record()
                if not is_conv and open_ev is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    evs.append((open_ev, e))
                    open_ev = None
                op(sp)
            if open_ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append((open_ev, e))
            torch.cuda.synchronize(dev)
            run_ms += sum(a.elapsed_time(b) for a, b in evs)
    by_class = {
        "ridge_flop_per_byte": round(ridge, 1),
        "timing": "one CUDA-event pair per launch (adds ~10 us of bubble per launch)",
        "tensor_bound_launches": {"launches": cls["tensor"][3], "ms": round(cls["tensor"][0], 3),
                                  "tflops": round(cls["tensor"][1] / max(1e-9, cls["tensor"][0]) / 1e9, 1),
                                  "frac_of_bf16_peak": round(cls["tensor"][1] / max(1e-9, cls["tensor"][0]) / 1e9 / pk["bf16_tflops_sustained"], 3)},
        "hbm_bound_launches": {"launches": cls["hbm"][3], "ms": round(cls["hbm"][0], 3),
                               "gbs": round(cls["hbm"][2] / max(1e-9, cls["hbm"][0]) / 1e6, 1),
                               "frac_of_hbm_peak": round(cls["hbm"][2] / max(1e-9, cls["hbm"][0]) / 1e6 / pk["hbm_gbs"], 3)},
    }
    return {"tflops": total_flops / (run_ms / 1e3) / 1e12, "ms": run_ms, "launches": n, "by_class": by_class,
            "per_launch_events_tflops": total_flops / (total_ms / 1e3) / 1e12,
            "kernel": "conv_simt_kernel" if net.precision == "fp32" else "conv_halo_kernel (tcgen05 implicit GEMM, halo-resident A tiles)"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
