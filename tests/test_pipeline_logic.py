"""RoundTripPipeline scheduling invariants on a fake codec (no GPU): every job is compressed and decompressed exactly
once, a launch-plan slot never has two jobs in flight, results come back in job order — for the serial scheduler, the
per-slot host threads and asymmetric compress / decompress slot counts."""
import threading
import time

import pytest

from rgbd_b200.pipeline import RoundTripPipeline


class _Stream:
    def synchronize(self):
        pass


class _Handle:
    def __init__(self, net, kind, slot, payload, delay):
        self.net, self.kind, self.slot, self.payload = net, kind, slot, payload
        self.stream = _Stream()
        self.t_done = time.monotonic() + delay
        self.collected = False

    def result(self, clone=True):
        time.sleep(max(0.0, self.t_done - time.monotonic()))
        with self.net.lock:
            assert not self.collected
            self.collected = True
            self.net.busy[self.slot] -= 1
        if self.kind == "c":
            return {"r_strings": [[b"r%d" % self.payload], [b"z"]], "d_strings": [[b"d%d" % self.payload], [b"z"]],
                    "shape": (1, 1)}
        return {"x_hat": {"r": ("xr", self.payload), "d": ("xd", self.payload)}}


class FakeNet:
    """compress_async / decompress_async with the real signatures; a slot that is entered twice trips an assert."""

    def __init__(self):
        self.lock = threading.Lock()
        self.busy = {}
        self.log = []

    def parameters(self):
        import torch
        return iter([torch.zeros(1)])

    def _enter(self, slot):
        with self.lock:
            self.busy[slot] = self.busy.get(slot, 0) + 1
            assert self.busy[slot] == 1, f"slot {slot} has two jobs in flight"

    def _slot_stream(self, slot):
        return _Stream()

    def compress_async(self, rgb, depth, slot=0):
        self._enter(slot)
        with self.lock:
            self.log.append(("c", rgb, slot))
        return _Handle(self, "c", slot, rgb, 0.004)

    def decompress_async(self, r_strings, d_strings, shape, slot=0):
        self._enter(slot)
        job = int(r_strings[0][0][1:])
        assert d_strings[0][0] == b"d%d" % job
        with self.lock:
            self.log.append(("d", job, slot))
        return _Handle(self, "d", slot, job, 0.010)


@pytest.mark.parametrize("threads,S,D,njobs", [(False, 3, None, 11), (True, 3, None, 11), (False, 2, 5, 13),
                                                (False, 4, None, 2), (True, 4, None, 1), (False, 1, None, 4)])
def test_pipeline_schedule_invariants(threads, S, D, njobs):
    net = FakeNet()
    pipe = RoundTripPipeline(net, S, threads=threads, high_priority_decode=False, dec_slots=D)
    if threads:
        pipe._ready.add(((), "fake"))             # plans "exist": skip the single-threaded warm-up pass
        pipe._prepare = lambda jobs, stage_input: None
    jobs = [(j, j) for j in range(njobs)]          # the fake codec carries the job index as its "tensors"
    res = pipe.run(jobs, keep_last=njobs)
    D = D or S
    assert [j for j, _, _ in res] == list(range(njobs))
    for j, c, (xr, xd) in res:
        assert c["r_strings"][0][0] == b"r%d" % j and xr == ("xr", j) and xd == ("xd", j)
    comp = [e for e in net.log if e[0] == "c"]
    dec = [e for e in net.log if e[0] == "d"]
    assert sorted(e[1] for e in comp) == list(range(njobs)) and sorted(e[1] for e in dec) == list(range(njobs))
    assert all(0 <= e[2] < S for e in comp) and all(S <= e[2] < S + D for e in dec)
    assert all(v == 0 for v in net.busy.values())


def test_pipeline_sink_and_staging_are_called_per_job():
    net = FakeNet()
    pipe = RoundTripPipeline(net, 2, high_priority_decode=False)
    staged, sunk = [], []
    import torch
    orig = torch.cuda.stream
    torch.cuda.stream = lambda s: __import__("contextlib").nullcontext()     # no CUDA in this test
    try:
        res = pipe.run([None] * 5, stage_input=lambda j, slot, st: (staged.append((j, slot)) or (j, j)),
                       sink=lambda j, slot, st, xr, xd: sunk.append((j, slot, xr)))
    finally:
        torch.cuda.stream = orig
    assert staged == [(j, j % 2) for j in range(5)]
    assert sorted(sunk) == [(j, j % 2, ("xr", j)) for j in range(5)]
    assert [j for j, _, _ in res] == [3, 4]
