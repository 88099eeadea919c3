"""Round-trip pipeline: keeps several batches in flight on one GPU.

The decoder's context chain is 20 serial rANS chunks per modality (elic_united.py:454-541) — about
125 ms of dependent-instruction latency per batch that occupies a handful of warps.  The codec's
convolutions are throughput work.  `RoundTripPipeline` therefore runs S compress jobs and S
decompress jobs concurrently, each on its own CUDA stream and launch-plan instance
(`compress_async` / `decompress_async` slots 0..S-1 and S..2S-1): while batch k is being decoded,
batches k+1.. are already being analysed, so the serial chains hide behind the tensor-core work.

With `threads=True` every slot is driven by its own host thread (compress -> collect the
strings -> decompress, job after job): a slot that is waiting for its coder kernels or copying its
bitstreams to the host does not hold back the submission of the other slots' work.  The CUDA calls
and the ctypes launches release the GIL.

Images are independent; nothing here changes what any single compress() /
decompress() call computes or the bytes it produces.
"""
import contextlib
import threading
from collections import deque

import torch


class RoundTripPipeline:
    def __init__(self, net, slots, threads=False, high_priority_decode=True, dec_slots=None):
        self.net = net
        self.S = max(1, int(slots))
        # a decompress job lives ~2.5x longer than a compress job (its serial rANS chain): with the same number of
        # launch plans, fewer compress slots and more decompress slots keep more of the chains in flight
        self.D = self.S if dec_slots is None else max(1, int(dec_slots))
        self.threads = bool(threads)
        self._ready = set()     # (B, H, W) whose 2S launch plans (and CUDA graphs) exist
        # The decoder's chain is ~260 small dependent kernels between its rANS chunks: on high-priority streams they
        # are scheduled ahead of the other slots' pending big convolutions as soon as SMs drain, which keeps the
        # chain's latency (and with it the number of slots needed to hide it) down.
        if high_priority_decode:
            device = next(net.parameters()).device
            streams = net.__dict__.setdefault("_slot_streams", {})
            for slot in range(self.S, self.S + self.D):
                if slot not in streams:
                    streams[slot] = torch.cuda.Stream(device, priority=-1)

    # ------------------------------------------------------------------ helpers
    def _input(self, n, slot, job, stage_input):
        if stage_input is None:
            return job
        with torch.cuda.stream(self.net._slot_stream(slot)):
            return stage_input(n, slot, self.net._slot_stream(slot))

    def _finish_decode(self, entry, sink, done):
        j, slot, c, h = entry
        r = h.result(clone=False)
        if sink is not None:
            with torch.cuda.stream(h.stream):
                sink(j, slot, h.stream, r["x_hat"]["r"], r["x_hat"]["d"])
        done.append((j, c, (r["x_hat"]["r"], r["x_hat"]["d"]), h.stream))

    def _prepare(self, jobs, stage_input):
        """Plans are built and CUDA graphs captured from ONE thread (stream capture is process-global): the first
        time a shape is seen, every slot runs one job to completion before the worker threads start."""
        rgb, depth = self._input(0, 0, jobs[0], stage_input)
        device = next(self.net.parameters()).device      # (the inputs may be pinned host tensors)
        key = (tuple(rgb.shape), str(device))
        if key in self._ready:
            return
        net, S = self.net, self.S
        c = None
        for slot in range(S):
            c = net.compress_async(rgb, depth, slot=slot).result()
        for slot in range(self.D):
            net.decompress_async(c["r_strings"], c["d_strings"], c["shape"], slot=S + slot).result(clone=False)
        torch.cuda.synchronize(device)
        self._ready.add(key)

    # ------------------------------------------------------------------ public
    def run(self, jobs, stage_input=None, sink=None, keep_last=None):
        """jobs: sequence of (rgb, depth) batches (device tensors, or whatever `stage_input` accepts).
        stage_input(job_index, slot, stream) -> (rgb_dev, depth_dev): optional H2D staging, called with
        the slot's stream current.  sink(job_index, slot, stream, x_r, x_d): optional consumer of the
        reconstruction, enqueued on the decoder's stream right after the decode (the buffers are
        reused by the slot's next job).  Returns [(job_index, compress_dict, (x_r, x_d))] for the last
        `keep_last` jobs (default and at most the number of decompress slots; their device buffers are still
        intact when run() returns)."""
        jobs = list(jobs)
        keep_last = self.D if keep_last is None else keep_last
        if not jobs:
            return []
        if self.threads and self.S > 1 and self.D == self.S and len(jobs) > 1:
            self._prepare(jobs, stage_input)
            done = self._run_threads(jobs, stage_input, sink)
        else:
            done = self._run_serial(jobs, stage_input, sink)
        done.sort(key=lambda e: e[0])
        out = []
        for j, c, xs, stream in done[-max(1, keep_last):]:
            stream.synchronize()
            out.append((j, c, xs))
        return out

    def _run_threads(self, jobs, stage_input, sink):
        net, S = self.net, self.S
        done, errors = [], []
        lock = threading.Lock()
        device = next(net.parameters()).device

        def worker(slot):
            try:
                on_dev = torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()
                with on_dev, torch.no_grad():
                    mine, prev = [], None
                    for n in range(slot, len(jobs), S):
                        rgb, depth = self._input(n, slot, jobs[n], stage_input)
                        c = net.compress_async(rgb, depth, slot=slot).result()
                        if prev is not None:          # the decoder plan of this slot is about to be reused
                            self._finish_decode(prev, sink, mine)
                        prev = (n, slot, c, net.decompress_async(c["r_strings"], c["d_strings"], c["shape"], slot=S + slot))
                    if prev is not None:
                        self._finish_decode(prev, sink, mine)
                    with lock:
                        done.extend(mine)
            except BaseException as e:   # surfaced in the caller's thread
                with lock:
                    errors.append(e)

        ts = [threading.Thread(target=worker, args=(s,), daemon=True) for s in range(min(S, len(jobs)))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return done

    def _run_serial(self, jobs, stage_input, sink):
        net, S, D = self.net, self.S, self.D
        enc, dec = deque(), deque()
        done = []

        def finish_encode():
            j, slot, h = enc.popleft()
            c = h.result()
            if len(dec) == D:           # the decoder slot of this job is still busy with job j - D
                self._finish_decode(dec.popleft(), sink, done)
            dslot = j % D
            dec.append((j, dslot, c, net.decompress_async(c["r_strings"], c["d_strings"], c["shape"], slot=S + dslot)))

        for n, job in enumerate(jobs):
            slot = n % S
            if len(enc) == S:
                finish_encode()
            rgb, depth = self._input(n, slot, job, stage_input)
            enc.append((n, slot, net.compress_async(rgb, depth, slot=slot)))
        while enc:
            finish_encode()
        while dec:
            self._finish_decode(dec.popleft(), sink, done)
        return done
