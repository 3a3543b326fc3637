"""Host<->device copy bandwidth of the box through the library's runtime: H2D alone, D2H alone, both at once on two
streams (pinned host memory). Tells whether the end-to-end leg of bench.py is bound by one direction or by the sum.

    python tools/pcie_duplex.py [MiB per copy]
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oflibnumpy_b200 as of  # noqa: E402
from oflibnumpy_b200 import _lib  # noqa: E402
from oflibnumpy_b200.device import Event, Stream  # noqa: E402

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = MB << 20
of.device.require_gpu()
c = _lib.call
h_in, h_out, d_a, d_b = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
c('ofk_rt_host_alloc', C.byref(h_in), n)
c('ofk_rt_host_alloc', C.byref(h_out), n)
s1, s2 = Stream(), Stream()
c('ofk_rt_malloc', C.byref(d_a), n, s1.handle)
c('ofk_rt_malloc', C.byref(d_b), n, s2.handle)
C.memset(h_in, 1, n)
C.memset(h_out, 0, n)
s1.synchronize(); s2.synchronize()


def run(h2d, d2h, reps=8, piece=None):
    piece = piece or n
    t0 = time.perf_counter()
    for _ in range(reps):
        for off in range(0, n, piece):
            if h2d:
                c('ofk_rt_memcpy_h2d', d_a.value + off, h_in.value + off, min(piece, n - off), s1.handle)
            if d2h:
                c('ofk_rt_memcpy_d2h', h_out.value + off, d_b.value + off, min(piece, n - off), s2.handle)
    s1.synchronize(); s2.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


run(True, True, 2)
print("H2D alone      %6.1f GB/s" % run(True, False))
print("D2H alone      %6.1f GB/s" % run(False, True))
r = run(True, True)
print("both at once   %6.1f GB/s per direction, %6.1f GB/s total" % (r, 2 * r))
for piece in (32 << 20, 8 << 20, 2 << 20):
    r = run(True, True, piece=piece)
    print("both, %3d MiB pieces  %6.1f GB/s per direction" % (piece >> 20, r))
