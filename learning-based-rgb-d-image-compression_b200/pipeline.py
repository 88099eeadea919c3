"""Round-trip pipeline: keeps several batches in flight on one GPU.

The decoder's context chain is 20 serial rANS chunks per modality (elic_united.py:454-541) — about
125 ms of dependent-instruction latency per batch that occupies a handful of warps.  The codec's
convolutions are throughput work.  `RoundTripPipeline` therefore runs S compress jobs and S
decompress jobs concurrently, each on its own CUDA stream and launch-plan instance
(`compress_async` / `decompress_async` slots 0..S-1 and S..2S-1): while batch k is being decoded,
batches k+1.. are already being analysed, so the serial chains hide behind the tensor-core work.

Images are independent; nothing here changes what any single compress() /
decompress() call computes or the bytes it produces.
"""
from collections import deque

import torch


class RoundTripPipeline:
    def __init__(self, net, slots):
        self.net = net
        self.S = max(1, int(slots))

    def run(self, jobs, stage_input=None, sink=None, keep_last=None):
        """jobs: iterable of (rgb, depth) batches (device tensors, or whatever `stage_input` accepts).
        stage_input(job_index, slot, stream) -> (rgb_dev, depth_dev): optional H2D staging, called with
        the slot's stream current.  sink(job_index, slot, stream, x_r, x_d): optional consumer of the
        reconstruction, enqueued on the decoder's stream right after the decode (the buffers are
        reused by the slot's next job).  Returns [(job_index, compress_dict, (x_r, x_d))] for the last
        `keep_last` jobs (default S; their device buffers are still intact when run() returns)."""
        net, S = self.net, self.S
        keep_last = S if keep_last is None else keep_last
        enc, dec = deque(), deque()
        done = deque(maxlen=max(1, keep_last))
        n = 0

        def finish_decode():
            j, slot, c, h = dec.popleft()
            r = h.result(clone=False)
            if sink is not None:
                with torch.cuda.stream(h.stream):
                    sink(j, slot, h.stream, r["x_hat"]["r"], r["x_hat"]["d"])
            done.append((j, c, (r["x_hat"]["r"], r["x_hat"]["d"]), h.stream))

        def finish_encode():
            j, slot, h = enc.popleft()
            c = h.result()
            if len(dec) == S:           # the decoder slot of this job is still busy with job j - S
                finish_decode()
            dec.append((j, slot, c, net.decompress_async(c["r_strings"], c["d_strings"], c["shape"], slot=S + slot)))

        for job in jobs:
            slot = n % S
            if len(enc) == S:
                finish_encode()
            if stage_input is not None:
                with torch.cuda.stream(net._slot_stream(slot)):
                    rgb, depth = stage_input(n, slot, net._slot_stream(slot))
            else:
                rgb, depth = job
            enc.append((n, slot, net.compress_async(rgb, depth, slot=slot)))
            n += 1
        while enc:
            finish_encode()
        while dec:
            finish_decode()
        out = []
        for j, c, xs, stream in done:
            stream.synchronize()
            out.append((j, c, xs))
        return out
