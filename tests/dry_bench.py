"""bench.py's B200 arm on the dry-run engine (TEST INFRASTRUCTURE, build container): the host
code of `run_b200` -- workload set-up, step loops, JSON assembly, parity check against the golden
grid, CPU baseline -- runs with tests/dryrun_engine.py standing in for the CUDA entry points and
stubs for CUDA events / graphs.  The numbers it prints are meaningless; the LINE's shape is what
tests/test_bench_contract.py checks.  Arguments are passed on to bench.py."""
import contextlib
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch                      # noqa: E402
import dryrun_engine              # noqa: E402

dryrun_engine.install()
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None


class _Event:
    def __init__(self, **k):
        self.t = None

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t)*1e3


class _Graph:
    def replay(self):
        time.sleep(0.001)


@contextlib.contextmanager
def _capture(graph):
    yield


torch.cuda.Event, torch.cuda.CUDAGraph, torch.cuda.graph = _Event, _Graph, _capture

import torch.distributed as dist  # noqa: E402

_init = dist.init_process_group


def _init_gloo(backend=None, **kw):              # multi-rank runs: gloo on CPU tensors
    kw.pop('device_id', None)
    return _init('gloo', **kw)


dist.init_process_group = _init_gloo

from rayoptics_b200 import engine as E      # noqa: E402

E.measure_fp64_peak = lambda device=0: 34.0

import bench                      # noqa: E402

if __name__ == '__main__':
    sys.argv = ['bench.py'] + sys.argv[1:]
    bench.main()
