"""Concurrent callers: the C-ABI may be entered from several host threads at once, each on its own stream
(the reference serialises its callers with the BitGenerator's lock, polyagamma/_polyagamma.pyx:172-176; a
GPU library is called from one thread per stream instead).  The library's shared state is the per-device
occupancy cache and helper streams (mutex-guarded), the per-thread event sets and the stream-ordered
scratch.  The test runs in a fresh interpreter so that the threads race on the library's *first* call too
(device cache, helper-stream creation, saddle start tables), then keeps them calling concurrently and
compares every result bit for bit with the same call made alone."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent(
    """
    import ctypes, sys, threading
    import numpy as np, torch
    sys.path.insert(0, %r)
    from polyagamma_b200 import _lib
    L = _lib.lib()
    torch.cuda.init(); torch.zeros(1, device="cuda")
    T, REPS = 4, 12
    sizes = [300_001, 70_000, 5_000, 1_200_003]          # helper streams, pipeline, one-launch kernel, two passes
    rng = np.random.default_rng(5)
    ins = []
    for n in sizes:
        h = rng.uniform(0.1, 200.0, n); h[::5] = rng.integers(1, 5, h[::5].shape)
        z = rng.uniform(-50.0, 50.0, n)
        ins.append((torch.from_numpy(h).cuda(), torch.from_numpy(z).cuda()))
    torch.cuda.synchronize()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    results = [[None] * REPS for _ in range(T)]
    errors = []
    barrier = threading.Barrier(T)
    def worker(t):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                barrier.wait()                            # all threads make the library's first call together
                for r in range(REPS):
                    h, z = ins[(t + r) %% len(sizes)]
                    out = torch.empty_like(h)
                    rc = L.pgm_cuda_random_polyagamma_fill2(100 + r, t, ptr(h), ptr(z), _lib.HYBRID, h.numel(), ptr(out),
                                                            0, ctypes.c_void_p(st.cuda_stream))
                    assert rc == 0, rc
                    results[t][r] = out
                st.synchronize()
        except BaseException as e:                        # noqa
            errors.append(repr(e))
    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]; [x.join() for x in th]
    assert not errors, errors
    torch.cuda.synchronize()
    bad = 0
    for t in range(T):
        for r in range(REPS):
            h, z = ins[(t + r) %% len(sizes)]
            ref = torch.empty_like(h)
            rc = L.pgm_cuda_random_polyagamma_fill2(100 + r, t, ptr(h), ptr(z), _lib.HYBRID, h.numel(), ptr(ref), 0,
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0
            torch.cuda.synchronize()
            if not torch.equal(results[t][r].view(torch.int64), ref.view(torch.int64)):
                bad += 1
    print("threads ok" if bad == 0 else f"{bad} of {T * REPS} concurrent results differ from the serial ones")
    sys.exit(0 if bad == 0 else 1)
    """
) % ROOT


def test_concurrent_host_threads_draw_what_serial_calls_draw():
    p = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "threads ok" in p.stdout
