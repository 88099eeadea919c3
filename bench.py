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

UNIT = "pairs/s"
MODELS = {"ELIC_united": True, "ELIC_united_R2D": False, "STF_united": True}      # name -> bidirectional (OracleCodec cross)


def make_oracle(model, sd, use_ref_coder=False, bf16=False):
    """The CPU restatement of `model` over the state_dict `sd` (bf16: the bf16-arithmetic emulation of the parity gate)."""
    if model == "STF_united":
        from oracle.bf16_emulation import Bf16StfOracle
        from oracle.stf_oracle import StfOracle
        return (Bf16StfOracle if bf16 else StfOracle)(sd, use_ref_coder=use_ref_coder)
    from oracle.bf16_emulation import Bf16OracleCodec
    from oracle.model_oracle import OracleCodec
    return (Bf16OracleCodec if bf16 else OracleCodec)(sd, cross=MODELS[model], use_ref_coder=use_ref_coder)


def metric_name(args):
    return f"RGB-D pairs/sec compress+decompress {args.height}x{args.width}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="ELIC_united", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=0, help="pairs per job (default: as many as fit the HBM budget)")
    ap.add_argument("--precision", default=os.environ.get("RGBD_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--slots", type=int, default=4, help="jobs in flight per GPU (own launch plan + CUDA stream each)")
    ap.add_argument("--jobs-per-step", type=int, default=0, help="jobs of --batch pairs per GPU and step (default: --slots)")
    ap.add_argument("--total-pairs", type=int, default=0,
                    help="strong scaling (BASELINE configs[2]): this many pairs per step in total, sharded "
                         "contiguously over the ranks; 0 = weak scaling (jobs-per-step x batch pairs per GPU)")
    ap.add_argument("--graphs", type=int, default=1, help="replay each slot's launch list as a CUDA graph")
    ap.add_argument("--threads", type=int, default=0, help="drive every pipeline slot from its own host thread")
    ap.add_argument("--dec-slots", type=int, default=0, help="decompress jobs in flight (default: = --slots)")
    ap.add_argument("--hiprio", type=int, default=1, help="decoder slots on high-priority CUDA streams")
    ap.add_argument("--preset", default="realistic")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the bounded cpu_baseline / parity sample (default: by size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 latency measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(args, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv launches of one encoder + decoder plan, from the committed
    ncu capture (profiles/r02_traffic.json, written by profiles/tools/traffic_from_ncu.py); None when there is no capture
    for this workload."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        for e in json.load(open(p)):
            if (e["model"], e["height"], e["width"], e["precision"]) == (args.model, args.height, args.width, args.precision):
                return {"dram_bytes": e["dram_bytes_per_pair"] * B, "bytes_per_pair": e["dram_bytes_per_pair"],
                        "algorithmic_bytes_per_pair": e.get("algorithmic_bytes_per_pair"), "source": e.get("source")}
    except Exception:
        return None
    return None


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
def make_inputs(n, H, W, seed, depth_div=10000.0):
    from rgbd_b200.synthetic import pad_to_multiple, synthetic_pairs
    rgb, depth = synthetic_pairs(n, H, W, seed=seed, depth_div=depth_div)
    return pad_to_multiple(rgb), pad_to_multiple(depth)


def depth_div(args):
    # SUN RGB-D stores depth * 10 (testing/tester_united.py:101-104): the R2D config is SUN-shaped
    return 100000.0 if args.model == "ELIC_united_R2D" else 10000.0


def default_cpu_pairs(args):
    if args.cpu_pairs:
        return args.cpu_pairs
    return max(1, min(8, int(8 * (512 * 640) / (args.height * args.width))))


def cpu_arm(args, pairs, steps, warmup, trace=False):
    """The path's CPU implementation timed on this box's host cores: oracle/model_oracle.py (torch CPU fp32 restatement,
    pinned bit-exactly to the reference) + the reference's own compiled rANS coder and pmf->cdf from oracle/_ref when
    they were built (else the C restatement).  Batch-1 loop like testing/tester_united.py, no file I/O.  This arm does
    not load the product library: the weights come from the synthetic generator (pure Python) and the CDF tables from
    oracle/tables.py."""
    import torch
    import rgbd_b200
    from oracle.ref_loader import ref_ext_available
    from oracle.tables import updated_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = getattr(rgbd_b200, args.model)(config=rgbd_b200.model_config(), channel=4).eval()   # key / shape holder only
    sd = updated_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, args.preset))
    use_ref = ref_ext_available()
    orc = make_oracle(args.model, sd, use_ref_coder=use_ref)
    rgb, depth = make_inputs(pairs, args.height, args.width, seed=1234, depth_div=depth_div(args))
    times = []
    sample = []      # per pair: stream bytes, symbols and reconstruction of the CPU path (the parity reference of the GPU line)
    H, W = args.height, args.width
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        outs = []
        for i in range(pairs):
            c = orc.compress(rgb[i:i + 1], depth[i:i + 1], trace=trace and not sample)
            tr = c.pop("_trace", None)
            outs.append((c, orc.decompress(c["r_strings"], c["d_strings"], c["shape"]), tr))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
        if not sample:
            for i, (c, r, tr) in enumerate(outs):
                row = {}
                for key, m, name, x in (("r_strings", "r", "rgb", rgb), ("d_strings", "d", "depth", depth)):
                    row["bytes_" + m] = sum(len(s_) for grp in c[key] for s_ in grp)
                    mse = float(((r["x_hat"][m][:, :, :H, :W].double() - x[i:i + 1, :, :H, :W].double()) ** 2).mean())
                    row["psnr_" + m] = 99.0 if mse <= 0 else 10 * math.log10(1.0 / mse)
                    row["xhat_" + m] = r["x_hat"][m]
                    if tr is not None:
                        row["sym_" + m] = tr["symbols"][(name, 0)][0]
                sample.append(row)
    total = sum(times)
    return {"pairs_coded": sample, "oracle": orc, "value": pairs * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "coder": "reference ans + _CXX (oracle/_ref)" if use_ref else "C restatement",
            "sample": f"{pairs} pair(s) x {len(times)} timed pass(es) of {args.height}x{args.width} "
                      f"{args.model} compress+decompress, batch 1, torch CPU fp32 ({warmup} warm-up)",
            "ms_per_pair": 1e3 * total / (pairs * len(times))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step = one pair; big images keep the whole run within a few minutes by capping the timed passes
    per_pair_s = 1.5 * (args.height * args.width) / (480 * 640)
    steps = max(1, min(args.steps, int(150 / per_pair_s)))
    warmup = max(0, min(args.warmup, max(1, int(20 / per_pair_s))))      # the asked-for warm-up, within ~20 s of CPU time
    cb = cpu_arm(args, 1, steps, warmup)
    line = {"impl": "reference", "metric": metric_name(args), "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_pair"], "higher_is_better": True,
            "scaling": "strong" if args.total_pairs else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "model": args.model, "sample": "batch 1 per step",
                       "preset": args.preset, "host": "cpu"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "coder")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    Hp, Wp = (args.height + 63) // 64 * 64, (args.width + 63) // 64 * 64
    return f"{args.model} compress+decompress of {args.height}x{args.width} pairs (padded {Hp}x{Wp})"


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    import rgbd_b200
    from rgbd_b200 import lib as L
    from rgbd_b200.parallel import add_pair_stats, allreduce_stats, new_stats, shard_range, summarize

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = getattr(rgbd_b200, args.model)(config=rgbd_b200.model_config(), channel=4, precision=args.precision).eval()
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, args.preset))
    net.update(force=True)
    net = net.to(dev)
    net.use_cuda_graph = bool(args.graphs)
    H, W = args.height, args.width
    Hp, Wp = (H + 63) // 64 * 64, (W + 63) // 64 * 64
    S = max(1, args.slots)
    D = args.dec_slots or S

    # ---- how many pairs this rank owns per step, and how they are cut into jobs
    if args.total_pairs:
        lo, hi = shard_range(args.total_pairs, rank, world)      # contiguous shard of the global batch
        own = hi - lo
        first_pair = lo
    else:
        own, first_pair = None, None

    def job_sizes(B):
        if own is None:
            return [B] * (args.jobs_per_step or S)
        return [B] * (own // B) + ([own % B] if own % B else [])

    def prepare(B):
        """Inputs, pinned host buffers, the pipeline and its launch plans for jobs of B pairs (warm-up included: building
        the plans is what allocates the HBM)."""
        sizes = job_sizes(B)
        n_pairs = sum(sizes)
        seed0 = 1234 + (first_pair if first_pair is not None else rank * n_pairs)
        rgb_h, depth_h = make_inputs(n_pairs, H, W, seed=seed0, depth_div=depth_div(args))
        rgb_h, depth_h = rgb_h.pin_memory(), depth_h.pin_memory()
        rgb_d, depth_d = rgb_h.to(dev), depth_h.to(dev)
        bounds = [0]
        for n in sizes:
            bounds.append(bounds[-1] + n)
        sl = [slice(bounds[i], bounds[i + 1]) for i in range(len(sizes))]
        J = len(sl)
        host_out = [(torch.empty((B, 3, Hp, Wp)).pin_memory(), torch.empty((B, 1, Hp, Wp)).pin_memory()) for _ in range(D)]

        # Round-trip pipeline (rgbd_b200.pipeline): S compress jobs and D decompress jobs in flight, each on its own
        # stream + launch plan, so the decoder's serial rANS chain of job k hides behind the convolutions of jobs
        # k+1..; one step = J jobs, every pair compressed AND decompressed inside the timed region.
        from rgbd_b200.pipeline import RoundTripPipeline
        pipe = RoundTripPipeline(net, S, threads=bool(args.threads), high_priority_decode=bool(args.hiprio), dec_slots=D)

        def steps_device(k):
            jobs = [(rgb_d[sl[i % J]], depth_d[sl[i % J]]) for i in range(k * J)]
            return pipe.run(jobs)

        def steps_e2e(k):
            def stage_input(j, slot, stream):
                # pinned HOST tensors straight into the public API: compress_async() copies them into its launch plan's input
                # buffers on the slot's stream (the H2D copy of this job's images, inside the timed region)
                return rgb_h[sl[j % J]], depth_h[sl[j % J]]

            def sink(j, slot, stream, x_r, x_d):
                # D2H of the reconstruction into pinned host buffers, image by image: a job's 126 MB as ONE copy would hold the
                # copy engine for milliseconds while other slots' small transfers (decoder states, stream words) queue behind it
                for i in range(x_r.shape[0]):
                    host_out[slot][0][i].copy_(x_r[i], non_blocking=True)
                    host_out[slot][1][i].copy_(x_d[i], non_blocking=True)

            res = pipe.run([None] * (k * J), stage_input=stage_input, sink=sink)
            torch.cuda.synchronize(dev)
            return res
        W_ = max(3, args.warmup)
        if os.environ.get("RGBD_BENCH_FAKE_OOM") and B == int(os.environ["RGBD_BENCH_FAKE_OOM"]):
            raise torch.OutOfMemoryError("fake OOM (test hook)")
        # every slot's launch plans (and CUDA graphs) must exist before the timed region: with few jobs per step (strong
        # scaling at many GPUs: one job per rank) W_ steps would not reach the higher slots
        steps_device(max(W_, -(-max(S, D) // J)))
        steps_e2e(max(1, -(-max(S, D) // J)))
        import types
        return types.SimpleNamespace(**{k: v for k, v in locals().items() if k != "types"})

    # pairs per job: as many as fit.  Plans cost ~0.28 GB per 512x640 image; 8 + 8 plans of 24 such pairs take ~107 GB of the
    # 180 GB.  If building them runs out of memory on this GPU (every rank must agree), fall back to a smaller job.
    if args.batch:
        candidates = [args.batch]
    elif args.precision == "fp32":
        candidates = [2, 1]
    else:
        # measured on B200 (profiles/README.md, round 2, one box): at the same HBM footprint 4 + 4 plans of 48 pairs (337 / 347 e2e
        # pairs/s, conv roofline 0.50) >= 5 + 5 of 40 (326 / 346, 0.49) >= 6 + 6 of 32 (327 / 329, 0.49) > 10 + 10 of 20 (318 / 304,
        # 0.45): the 32x40 context-model launches fill the GPU better the bigger the job
        b0 = 24 * (512 * 640) / (Hp * Wp) * 16 / (S + D)
        b0 = max(1, min(48, int(round(b0 / 4) * 4) if b0 >= 16 else int(b0)))
        if own is not None:
            # strong scaling: equal jobs only (a ragged last job would need a second set of launch plans per slot)
            divs = [d_ for d_ in range(1, own + 1) if own % d_ == 0 and d_ <= b0]
            candidates = sorted(set(divs[-3:]), reverse=True)
        else:
            candidates = sorted({b0, max(1, b0 * 2 // 3), max(1, b0 // 3), 1}, reverse=True)
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
    rgb_h, depth_h, rgb_d, depth_d, sl, J = ns.rgb_h, ns.depth_h, ns.rgb_d, ns.depth_d, ns.sl, ns.J
    pipe, steps_device, steps_e2e = ns.pipe, ns.steps_device, ns.steps_e2e
    pairs_per_step = ns.n_pairs                      # this rank's pairs per step

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
    total_pairs_step = torch.tensor([pairs_per_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_pairs_step, op=dist.ReduceOp.SUM)
    total_pairs_step = int(total_pairs_step.item())
    L.load().rgbd_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res = timed(steps_device, K)
    launches = int(L.load().rgbd_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    value = total_pairs_step * K / (ms / 1e3)
    stats = new_stats()
    for j, c, (xr, xd) in res:
        s_ = sl[j % J]
        add_pair_stats(stats, c["r_strings"], c["d_strings"], rgb_d[s_][:, :, :H, :W], depth_d[s_][:, :, :H, :W],
                       xr[:, :, :H, :W], xd[:, :, :H, :W])

    s0 = ClockSampler(local)
    if rank == 0:
        s0.start()
    ms_e2e, res2 = timed(steps_e2e, K)
    clocks_e2e = s0.stop() if rank == 0 else None
    e2e_value = total_pairs_step * K / (ms_e2e / 1e3)
    # bytes per step: every pair's images up, its four strings down and up again (they are host `bytes`), its reconstruction down
    bytes_per_pair_strings = sum(len(x) for _, c, _ in res2 for key in ("r_strings", "d_strings") for grp in c[key]
                                 for x in grp) / max(1, sum(len(c["r_strings"][1]) for _, c, _ in res2))
    img_bytes = 4 * 4 * Hp * Wp
    h2d = int(pairs_per_step * (img_bytes + bytes_per_pair_strings))
    d2h = int(pairs_per_step * (img_bytes + bytes_per_pair_strings))

    # roofline of the dominant kernel family (the implicit-GEMM conv): CUDA events on the launching stream over the
    # plans the timed run used
    roof = conv_roofline(net, B, Hp, Wp, dev, dec_slot=S)
    gflop_per_pair = roof["plan_gflop"] / B
    stats = summarize(allreduce_stats(stats, dev))
    hbm_peak = round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)

    line = None
    if rank == 0:
        pk, pk_src = peaks()
        tens_peak = pk["bf16_tflops_sustained"]
        traffic = ncu_traffic(args, B)
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.total_pairs else "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args) + f", {total_pairs_step} pairs/step over {world} GPU(s), preset "
                                   f"{args.preset}, weights calibrated random-init",
                       "model": args.model, "pairs_per_step": total_pairs_step, "pairs_per_gpu": pairs_per_step,
                       "jobs_per_gpu_step": J, "batch": B, "slots_in_flight": S, "dec_slots": D,
                       "schedule": "pipelined: S compress + D decompress jobs in flight (rgbd_b200.pipeline)",
                       "host_threads": bool(args.threads), "precision": args.precision, "cuda_graphs": bool(args.graphs),
                       "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)", "hbm_peak_gb": hbm_peak,
                       "parallelism": f"dp{world} (images sharded, no data-path collective)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / K, "sm_mhz": clocks_e2e.get("sm_mhz") if clocks_e2e else None},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": roof["tflops"], "peak": tens_peak, "unit": "TFLOP/s",
                         "frac": roof["tflops"] / tens_peak,
                         "traffic": traffic["dram_bytes"] if traffic else None, "traffic_detail": traffic,
                         "kernel": roof["kernel"], "launches": roof["launches"], "peak_source": pk_src + " bf16 sustained",
                         "share_of_step": (pairs_per_step / B) * roof["ms"] / (ms / K),
                         "algorithmic_gflop_per_pair": gflop_per_pair, "by_class": roof["by_class"],
                         "by_stage": roof["by_stage"],
                         "timing": "CUDA events around every run of consecutive conv launches of one compress + decompress plan",
                         "achieved_with_per_launch_events": roof["per_launch_events_tflops"],
                         "whole_step_tflops": gflop_per_pair * total_pairs_step * K / ms / world},
            "quality": stats,
        }
    # ---- untimed extras on rank 0: free the pipeline's plans, then batch-1 latency and the parity gate
    del ns, pipe, steps_device, steps_e2e, res, res2
    if world > 1:
        # the CPU baseline, the parity gate and the batch-1 latency belong to the N = 1 line only (the other ranks would sit in
        # a barrier meanwhile; measured at N = 8: the CPU path ran 35x slower beside seven spinning ranks)
        args.no_cpu_baseline = args.no_latency = True
    if rank == 0 and not (args.no_latency and args.no_cpu_baseline):
        net._invalidate()
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        if not args.no_latency:
            line["latency_b1_ms"] = latency_b1(net, rgb_d[:1], depth_d[:1], dev, args)
        if not args.no_cpu_baseline:
            n = default_cpu_pairs(args)
            cb = cpu_arm(args, n, 1, 0, trace=True)
            coded, orc = cb.pop("pairs_coded"), cb.pop("oracle")
            line["cpu_baseline"] = cb
            line["parity_vs_cpu_path"] = parity_gate(net, args, coded, dev, orc)
    if rank == 0:
        line.setdefault("cpu_baseline", None)      # (N = 1 only, see above)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_b1(net, rgb1, depth1, dev, args=None):
    """Batch-1 latency through the public API (BASELINE configs[1] is a single pair): median of 5 calls, wall clock with
    a device synchronize on both sides; in the reference-decodable single-stream layout and in the opt-in multi-stream
    layout (the y stream of an image cut into equal sub-streams, decoded concurrently per coding step)."""
    import torch
    out = _latency_once(net, rgb1, depth1, dev)
    out["mode"] = "one pair, one job in flight, reference-decodable single-stream layout"
    if args is not None:
        import rgbd_b200
        try:
            multi = getattr(rgbd_b200, args.model)(config=rgbd_b200.model_config(), channel=4, precision=args.precision,
                                                   stream_layout="multi", sub_channels=4).eval()
            multi.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
            multi.update(force=True)
            multi = multi.to(dev)
            multi.use_cuda_graph = net.use_cuda_graph
            m = _latency_once(multi, rgb1, depth1, dev)
            m["mode"] = "multi-stream layout, sub_channels=4 (160 sub-streams per image and modality; opt-in, not reference-decodable)"
            out["multi_stream"] = m
            del multi
        except Exception as e:       # the headline numbers above do not depend on this extra
            out["multi_stream"] = {"error": str(e)[:200]}
    return out


def _latency_once(net, rgb1, depth1, dev):
    import torch
    c = net.compress(rgb1, depth1)
    net.decompress(c["r_strings"], c["d_strings"], c["shape"])      # plans (and graphs) exist now
    tc, td = [], []
    for _ in range(5):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        c = net.compress(rgb1, depth1)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        tc.append(1e3 * (t1 - t0))
        td.append(1e3 * (t2 - t1))
    tc.sort()
    td.sort()
    return {"compress": round(tc[2], 2), "decompress": round(td[2], 2), "total": round(tc[2] + td[2], 2)}


def parity_gate(net, args, coded, dev, orc):
    """Correctness gate next to the throughput number: the pairs of the CPU sample through this GPU build —
    per pair and modality the bpp deviation, the symbol mismatch rate and the distance between the two reconstructions."""
    import numpy as np
    import torch
    n = len(coded)
    H, W = args.height, args.width
    rgb, depth = make_inputs(n, H, W, seed=1234, depth_div=depth_div(args))
    use_graph, net.use_cuda_graph = net.use_cuda_graph, False
    bpp_dev, sym_mis, xpsnr, psnr_dev, emu_db = [], [], [], [], []
    for i, row in enumerate(coded):
        r1, d1 = rgb[i:i + 1].to(dev), depth[i:i + 1].to(dev)
        c = net.compress(r1, d1)
        prog = net._program("encoder", 1, r1.shape[2], r1.shape[3])
        sym = {k: prog.io["st"][k]["ysym"][0].cpu().numpy() for k in ("r", "d")}
        rec = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        # reconstruction fidelity: the CPU path's synthesis transform on the GPU's own y_hat (the two complete codecs'
        # y_hat differ wherever a bf16 rounding flips a symbol, which random-init weights amplify chaotically)
        dec = net._program("decoder", 1, int(c["shape"][0]), int(c["shape"][1]))
        nchw = lambda v: v.torch().float().cpu().permute(0, 3, 1, 2).contiguous()
        yh = [nchw(dec.io["yhat"][k]) for k in ("r", "d")]
        gs = dict(zip(("r", "d"), orc.g_s(*yh)))
        pre = {k: nchw(dec.io["x_nhwc"][k]) for k in ("r", "d")}       # before decompress()'s clamp to [0, 1]
        if i == 0 and net.precision == "bf16":                         # what bf16 arithmetic itself costs on these weights
            for m, e in zip(("r", "d"), make_oracle(args.model, orc.sd, bf16=True).g_s(*yh)):
                emu_db.append(10 * math.log10(16.0 * float(gs[m].double().var()) / max(1e-30, float(((e.double() - gs[m].double()) ** 2).mean()))))
        for key, m, x in (("r_strings", "r", rgb), ("d_strings", "d", depth)):
            nbytes = sum(len(s_) for grp in c[key] for s_ in grp)
            bpp_dev.append(100.0 * abs(nbytes - row["bytes_" + m]) / row["bytes_" + m])
            if "sym_" + m in row:
                sym_mis.append(100.0 * float((sym[m] != np.asarray(row["sym_" + m])).mean()))
            xh = rec["x_hat"][m].cpu()
            mse = float(((pre[m].double() - gs[m].double()) ** 2).mean())
            peak2 = 16.0 * float(gs[m].double().var())      # peak = 4 sigma of the CPU path's output (a natural image's ratio)
            xpsnr.append(99.0 if mse <= 0 else 10 * math.log10(peak2 / mse))
            mse_in = float(((xh[:, :, :H, :W].double() - x[i:i + 1, :, :H, :W].double()) ** 2).mean())
            psnr_dev.append(abs((99.0 if mse_in <= 0 else 10 * math.log10(1.0 / mse_in)) - row["psnr_" + m]))
    net.use_cuda_graph = use_graph
    return {"pairs": n, "max_bpp_dev_pct": round(max(bpp_dev), 4),
            "max_symbol_mismatch_pct": round(max(sym_mis), 4) if sym_mis else None,
            "min_psnr_xhat_gpu_vs_cpu_gs_on_same_yhat_db": round(min(xpsnr), 2),
            "bf16_emulation_of_cpu_gs_db": round(min(emu_db), 2) if emu_db else None,
            "max_psnr_dev_db": round(max(psnr_dev), 5),
            "tolerance": "bpp 0.5 %, PSNR vs input 0.05 dB (BASELINE north_star); reconstruction: the GPU's pre-clamp x_hat "
                         "against the CPU path's g_s on the same y_hat (PSNR, peak = 4 sigma) within 1.5 dB of a torch-CPU "
                         "emulation of bf16 convs (oracle/bf16_emulation.py)"}


def conv_roofline(net, B, Hp, Wp, dev, dec_slot=0):
    """Algorithmic conv flops / sum of conv launch durations for one encoder + decoder pass, the same split by which roof
    bounds each launch (arithmetic intensity above / below the ridge of the measured peaks) and by stage of the codec."""
    import ctypes as C
    import torch
    pk, _ = peaks()
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)     # flop per byte
    total_ms, total_flops, n = 0.0, 0.0, 0
    # ms, flops, bytes, launches per roof: arithmetic intensity > 1.1 x the ridge -> tensor, < 0.9 x -> hbm, else "ridge" (the fused
    # bottleneck blocks: 208 flop/byte against a ridge of 210 — both roofs bound them equally)
    cls = {"tensor": [0.0, 0.0, 0.0, 0], "ridge": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}
    stages = {}                                                             # stage -> [conv ms, conv flops, conv launches, other ms, other launches]
    progs = (net._program("encoder", B, Hp, Wp),
             net._program("decoder", B, Hp // 64, Wp // 64, slot=dec_slot))   # plans the timed run used: their buffers hold real streams
    with torch.cuda.device(dev):
        for prog in progs:
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            evs = []
            for op in prog.ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                op(sp)
                b.record()
                evs.append((a, b, op))
            torch.cuda.synchronize(dev)
            for a, b, op in evs:
                ms = a.elapsed_time(b)
                st = stages.setdefault(getattr(op, "stage", "") or "other", [0.0, 0.0, 0, 0.0, 0])
                if not getattr(op, "is_conv", False):
                    st[3] += ms
                    st[4] += 1
                    continue
                total_ms += ms
                st[0] += ms
                st[1] += op.flops
                st[2] += 1
                ai = op.flops / max(1, op.bytes)
                c = cls["tensor" if ai > 1.1 * ridge else ("hbm" if ai < 0.9 * ridge else "ridge")]
                c[0] += ms
                c[1] += op.flops
                c[2] += op.bytes
                c[3] += 1
                n += 1
            total_flops += prog.flops
    # Second pass: one event pair around every maximal RUN of consecutive conv launches instead of around every launch.
    # Per-launch events put ~10 us of bubble around each of the launches (the kernels measure 10+ us shorter under
    # ncu) and forbid the programmatic-dependent-launch overlap the real run has; per-run events keep launch gaps and
    # PDL exactly as in production.  `achieved` uses this pass; the per-launch pass feeds the by-class / by-stage splits.
    run_ms = 0.0
    with torch.cuda.device(dev):
        for prog in progs:
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
    peak = pk["bf16_tflops_sustained"]
    by_class = {
        "ridge_flop_per_byte": round(ridge, 1),
        "timing": "one CUDA-event pair per launch (adds ~10 us of bubble per launch)",
        "tensor_bound_launches": {"launches": cls["tensor"][3], "ms": round(cls["tensor"][0], 3),
                                  "tflops": round(cls["tensor"][1] / max(1e-9, cls["tensor"][0]) / 1e9, 1),
                                  "frac_of_bf16_peak": round(cls["tensor"][1] / max(1e-9, cls["tensor"][0]) / 1e9 / peak, 3)},
        "ridge_launches": {"launches": cls["ridge"][3], "ms": round(cls["ridge"][0], 3),
                           "tflops": round(cls["ridge"][1] / max(1e-9, cls["ridge"][0]) / 1e9, 1),
                           "frac_of_bf16_peak": round(cls["ridge"][1] / max(1e-9, cls["ridge"][0]) / 1e9 / peak, 3),
                           "gbs": round(cls["ridge"][2] / max(1e-9, cls["ridge"][0]) / 1e6, 1),
                           "frac_of_hbm_peak": round(cls["ridge"][2] / max(1e-9, cls["ridge"][0]) / 1e6 / pk["hbm_gbs"], 3),
                           "share_of_conv_ms": round(cls["ridge"][0] / max(1e-9, total_ms), 3),
                           "what": "arithmetic intensity within 10 % of the ridge: the fused bottleneck blocks"},
        "hbm_bound_launches": {"launches": cls["hbm"][3], "ms": round(cls["hbm"][0], 3),
                               "gbs": round(cls["hbm"][2] / max(1e-9, cls["hbm"][0]) / 1e6, 1),
                               "frac_of_hbm_peak": round(cls["hbm"][2] / max(1e-9, cls["hbm"][0]) / 1e6 / pk["hbm_gbs"], 3),
                               "share_of_conv_ms": round(cls["hbm"][0] / max(1e-9, total_ms), 3)},
    }
    by_stage = {}
    for name, (cms, cfl, cn, oms, on) in sorted(stages.items()):
        by_stage[name] = {"conv_launches": cn, "conv_ms": round(cms, 3), "gflop": round(cfl / 1e9, 1),
                          "tflops": round(cfl / max(1e-9, cms) / 1e9, 1) if cn else None,
                          "frac_of_bf16_peak": round(cfl / max(1e-9, cms) / 1e9 / peak, 3) if cn else None,
                          "other_launches": on, "other_ms": round(oms, 3)}
    return {"tflops": total_flops / (run_ms / 1e3) / 1e12, "ms": run_ms, "launches": n, "by_class": by_class,
            "by_stage": by_stage, "plan_gflop": total_flops / 1e9,
            "per_launch_events_tflops": total_flops / (total_ms / 1e3) / 1e12,
            "kernel": "conv_simt_kernel" if net.precision == "fp32" else
                      "conv_halo_kernel (tcgen05 implicit GEMM, halo-resident A tiles) + rb_fused_kernel (fused bottleneck blocks)"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
