"""Host-layer behaviour of libsigops on real devices: pageable vs pinned buffers (the library's own staging), launch
geometry variants, error discipline (fault injection: all-or-nothing, nothing left in flight, pool stays usable),
per-device locking (two callers on two devices overlap), `_device` calls on several streams sharing a device's tables,
multi-device queues and the device-resident producer.  The reference has none of this to mirror: it creates and destroys
a wgpu device per call (src/gpu.rs:5-35,167) and its only error channel is all-or-nothing (src/secp256k1_ecdsa.rs:203-205)."""
import ctypes
import os
import threading
import time

import numpy as np
import pytest

import batches

pytestmark = pytest.mark.gpu


def _pinned(lib, a):
    ptr = lib.sigops_host_alloc(max(1, a.nbytes))
    buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(1, a.nbytes),))
    buf[: a.nbytes] = a.reshape(-1).view(np.uint8)
    return ptr, buf


@pytest.mark.parametrize("staging", ["auto", "0", "1"])
@pytest.mark.parametrize("n", [1000, 16384, 200003])
def test_pageable_and_pinned_buffers_agree(sigops, monkeypatch, staging, n):
    """Pageable (numpy) buffers go through the library's pinned staging from 16,384 signatures per shard on; the result
    must not depend on the path (SIGOPS_STAGING: auto / never / always) nor on where the caller's memory lives."""
    if staging != "auto":
        monkeypatch.setenv("SIGOPS_STAGING", staging)
    lib = sigops.load()
    s, m, pk, st, _ = batches.ecdsa_batch(0, n, edge_every=101, seed=61)
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s2, m2, p2, v, _ = batches.ed25519_batch(n, edge_every=7, seed=62)
    assert (sigops.ed25519_eddsa.ecverify_array(s2, m2, p2) == v).all()
    # pinned inputs, pageable outputs and the other way round
    ps, _ = _pinned(lib, s)
    pm, _ = _pinned(lib, m)
    po, vo = _pinned(lib, np.zeros(n * 64, np.uint8))
    pt, vt = _pinned(lib, np.zeros(n, np.uint8))
    try:
        o2, t2 = np.zeros((n, 64), np.uint8), np.zeros(n, np.uint8)
        assert lib.sigops_secp256k1_ecrecover(ps, pm, n, o2.ctypes.data, t2.ctypes.data) == 0
        assert (o2 == pk).all() and (t2 == st).all()
        assert lib.sigops_secp256k1_ecrecover(s.ctypes.data, m.ctypes.data, n, po, pt) == 0
        assert (vo.reshape(n, 64) == pk).all() and (vt == st).all()
        vo[:] = 0
        assert lib.sigops_secp256k1_ecrecover(ps, pm, n, po, None) == 0  # status not requested
        assert (vo.reshape(n, 64) == pk).all()
    finally:
        for p in (ps, pm, po, pt):
            lib.sigops_host_free(p)


@pytest.mark.parametrize("n", [75776 - 1, 75776, 75776 + 1, 131072, 2 * 75776 + 5, 303104 + 77])
def test_launch_geometry_variants(sigops, monkeypatch, n):
    """Full blocks plus a separate tail launch (the default), a single launch (SIGOPS_TAIL_SPLIT=0), the balanced
    geometry (SIGOPS_BALANCED=1: P passes of ceil(n / (SMs x P)) threads per SM) and the unpipelined path must give
    identical, exact results on both sides of every seam."""
    s, m, pk, st, _ = batches.ecdsa_batch(1, n, edge_every=97, seed=71, mix_high_s=True)
    s2, m2, p2, v, _ = batches.ed25519_batch(n, edge_every=11, seed=72)
    for env in ({}, {"SIGOPS_TAIL_SPLIT": "0"}, {"SIGOPS_BALANCED": "1"}, {"SIGOPS_MAX_CHUNKS": "1"}):
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        out, got = sigops.secp256r1_ecdsa.ecrecover_with_status(s, m)
        assert (out == pk).all() and (got == st).all(), env
        assert (sigops.ed25519_eddsa.ecverify_array(s2, m2, p2) == v).all(), env
        for k in env:
            monkeypatch.delenv(k)


def test_fault_injection_is_all_or_nothing(sigops, monkeypatch):
    """SIGOPS_FAIL_DEVICE=<pool index> fails that device's shard with uploads, a kernel and downloads in flight: the call
    returns nonzero (-> ShaderFailureError), no partial results are left in the caller's buffers, nothing is still
    running when it returns, and the very next call succeeds."""
    lib = sigops.load()
    n = 200003
    s, m, pk, st, _ = batches.ecdsa_batch(0, n, edge_every=101, seed=81)
    G = lib.sigops_num_devices()
    for k in sorted({0, G - 1}):
        monkeypatch.setenv("SIGOPS_FAIL_DEVICE", str(k))
        out = np.full((n, 64), 0xAA, np.uint8)
        got = np.full(n, 0xAA, np.uint8)
        rc = lib.sigops_secp256k1_ecrecover(s.ctypes.data, m.ctypes.data, n, out.ctypes.data, got.ctypes.data)
        assert rc != 0
        assert b"injected failure" in lib.sigops_last_error()
        assert not out.any() and (got == 1).all()  # reset, not partially filled
        with pytest.raises(sigops.ShaderFailureError):
            sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
        v = np.full(n, 0xAA, np.uint8)
        s2, m2, p2, _v, _ = batches.ed25519_batch(1000, edge_every=7, seed=82)
        if k == 0:  # small batches run on pool device 0
            assert lib.sigops_ed25519_ecverify(s2.ctypes.data, m2.ctypes.data, p2.ctypes.data, 1000, v.ctypes.data) != 0
            assert not v[:1000].any()
        monkeypatch.delenv("SIGOPS_FAIL_DEVICE")
        out2, got2 = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
        assert (out2 == pk).all() and (got2 == st).all()


def test_conflicting_init_is_refused(sigops):
    lib = sigops.load()
    G = lib.sigops_num_devices()  # lazy init with the defaults
    assert G >= 1
    same = (ctypes.c_int * G)(*range(G))
    assert lib.sigops_init(same, G) == 0  # the same set: fine
    assert lib.sigops_init(None, 0) == 0  # defaults: satisfied by the existing pool
    if G > 1:
        one = (ctypes.c_int * 1)(G - 1)
        assert lib.sigops_init(one, 1) != 0
        assert b"different device set" in lib.sigops_last_error()
    bad = (ctypes.c_int * 1)(99)
    assert lib.sigops_init(bad, 1) != 0


def test_device_calls_on_two_streams_do_not_share_tables(sigops):
    """Two `_device` launches on different streams (and a host-buffer call on top) use the same per-device work tables:
    the library orders them with events.  Without that they overlap and silently corrupt each other's tables."""
    import torch

    lib = sigops.load()
    lib.sigops_num_devices()
    torch.cuda.set_device(0)
    n = 60000
    a = batches.ecdsa_batch(0, n, edge_every=101, seed=91)
    b = batches.ecdsa_batch(1, n, edge_every=103, seed=92, mix_high_s=True)
    dev = torch.device("cuda", 0)
    da = [torch.from_numpy(x).to(dev) for x in a[:2]]
    db = [torch.from_numpy(x).to(dev) for x in b[:2]]
    oa, ta = torch.zeros((n, 64), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    ob, tb = torch.zeros((n, 64), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    errs = []

    def host_call():
        out, st = sigops.secp256k1_ecdsa.ecrecover_with_status(a[0][:5000], a[1][:5000])
        if not ((out == a[2][:5000]).all() and (st == a[3][:5000]).all()):
            errs.append("host")

    for _ in range(3):
        assert lib.sigops_secp256k1_ecrecover_device(da[0].data_ptr(), da[1].data_ptr(), n, oa.data_ptr(), ta.data_ptr(),
                                                     ctypes.c_void_p(s1.cuda_stream)) == 0
        assert lib.sigops_secp256r1_ecrecover_device(db[0].data_ptr(), db[1].data_ptr(), n, ob.data_ptr(), tb.data_ptr(),
                                                     ctypes.c_void_p(s2.cuda_stream)) == 0
        t = threading.Thread(target=host_call)
        t.start()
        t.join()
    torch.cuda.synchronize()
    assert not errs
    assert (oa.cpu().numpy() == a[2]).all() and (ta.cpu().numpy() == a[3]).all()
    assert (ob.cpu().numpy() == b[2]).all() and (tb.cpu().numpy() == b[3]).all()


def test_two_callers_on_two_devices_overlap(sigops):
    """Per-device locks: two threads that use disjoint devices (sigops_batch_on_devices) run concurrently -- together they
    take clearly less than the sum of their solo times -- and a whole-pool caller still gets exact results meanwhile."""
    lib = sigops.load()
    if lib.sigops_num_devices() < 2:
        pytest.skip("one visible GPU")
    n = 400000
    k1 = batches.ecdsa_batch(0, n, edge_every=1000, seed=101)
    ed = batches.ed25519_batch(n, edge_every=100, seed=102)
    res = {}

    def run(tag, dev):
        idx = (ctypes.c_int * 1)(dev)
        t0 = time.perf_counter()
        if tag == "k1":
            out, st = np.zeros((n, 64), np.uint8), np.zeros(n, np.uint8)
            rc = lib.sigops_batch_on_devices(0, idx, 1, k1[0].ctypes.data, k1[1].ctypes.data, None, n, out.ctypes.data, st.ctypes.data)
            ok = rc == 0 and (out == k1[2]).all() and (st == k1[3]).all()
        else:
            v = np.zeros(n, np.uint8)
            rc = lib.sigops_batch_on_devices(2, idx, 1, ed[0].ctypes.data, ed[1].ctypes.data, ed[2].ctypes.data, n, v.ctypes.data, None)
            ok = rc == 0 and (v == ed[3]).all()
        res[tag] = (ok, time.perf_counter() - t0)

    for _ in range(2):  # warm both devices' buffers
        run("k1", 0)
        run("ed", 1)
    solo = res["k1"][1] + res["ed"][1]
    t0 = time.perf_counter()
    ts = [threading.Thread(target=run, args=("k1", 0)), threading.Thread(target=run, args=("ed", 1))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    both = time.perf_counter() - t0
    assert res["k1"][0] and res["ed"][0]
    assert both < 0.8 * solo, (both, solo)
    bad = (ctypes.c_int * 2)(0, 0)
    assert lib.sigops_batch_on_devices(0, bad, 2, k1[0].ctypes.data, k1[1].ctypes.data, None, 10, k1[2].ctypes.data, None) != 0


def test_queue_over_all_devices_and_device_resident_producer(sigops):
    """A queue with device_index = -1 spreads its slots over the pool (slot i on device i mod G).  submit_device takes the
    request from device memory of ANY GPU of the box -- cudaMemcpyPeerAsync into the slot's buffers (NVLink between
    peers) -- so the host never touches the inputs.  Every result bit-exact, edge rows included."""
    import torch

    lib = sigops.load()
    G = lib.sigops_num_devices()
    n = 3000
    rows = batches.ecdsa_batch(0, 8 * n, edge_every=53, seed=111)
    with sigops.service.SigQueue("secp256k1", n, depth=8, device_index=-1) as q:
        assert q.info()["device_index"] == -1
        assert {q.slot_device(i) for i in range(8)} == set(range(min(G, 8)))
        for rep in range(2):
            for slot in range(8):
                a = slot * n
                q.sigs(slot)[:n] = rows[0][a:a + n]
                q.msgs(slot)[:n] = rows[1][a:a + n]
                q.submit(slot, n)
            for slot in range(8):
                a = slot * n
                out, st = q.wait(slot)
                assert (out == rows[2][a:a + n]).all() and (st == rows[3][a:a + n]).all()
        # device-resident producer: inputs live on the LAST GPU, slots on all of them
        src = G - 1
        dev = torch.device("cuda", src)
        d_s = torch.from_numpy(rows[0]).to(dev)
        d_m = torch.from_numpy(rows[1]).to(dev)
        torch.cuda.synchronize(dev)
        for slot in range(8):
            a = slot * n
            q.submit_device(slot, d_s.data_ptr() + a * 64, d_m.data_ptr() + a * 32, None, n, src)
        for slot in range(8):
            a = slot * n
            out, st = q.wait(slot)
            assert (out == rows[2][a:a + n]).all() and (st == rows[3][a:a + n]).all(), slot
    ed = batches.ed25519_batch(2 * n, edge_every=5, seed=112)
    with sigops.service.SigQueue("ed25519", n, depth=2, device_index=0) as q:
        dev = torch.device("cuda", G - 1)
        d = [torch.from_numpy(x).to(dev) for x in ed[:3]]
        torch.cuda.synchronize(dev)
        for slot in range(2):
            a = slot * n
            q.submit_device(slot, d[0].data_ptr() + a * 64, d[1].data_ptr() + a * 32, d[2].data_ptr() + a * 32, n, G - 1)
        for slot in range(2):
            a = slot * n
            v, _ = q.wait(slot)
            assert (v == ed[3][a:a + n]).all()
        # a second waiter on a slot that is not in flight returns at once; closing with views handed out needs force
        view, _ = q.wait(0, copy=False)
        with pytest.raises(RuntimeError):
            q.close()
        del view


_GWIN_CHILD = r"""
import os, sys
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import batches, unit_checks as uc
from simlib import UnitRunner
import wgpu_sigops_b200 as w
lib = w.load()
w_bits = int(os.environ["SIGOPS_GWIN"])
uc.check_fixed_base(UnitRunner(lib.sigops_test_unit, lib.sigops_test_unit_shape), w_bits)
for n, env in ((700, "1"), (40000, "0")):           # lane-group kernels, then one signature per thread
    os.environ["SIGOPS_LANEGROUP"] = env
    for cid, mod in ((0, w.secp256k1_ecdsa), (1, w.secp256r1_ecdsa)):
        s, m, pk, st, _ = batches.ecdsa_batch(cid, n, edge_every=53, seed=70 + cid)
        out, got = mod.ecrecover_with_status(s, m)
        assert (out == pk).all() and (got == st).all(), (w_bits, cid, n)
    s, m, p, v, _ = batches.ed25519_batch(n, edge_every=11, seed=73)
    assert (w.ed25519_eddsa.ecverify_array(s, m, p) == v).all(), (w_bits, n)
print("ok", w_bits)
"""


@pytest.mark.parametrize("gwin", ["4", "9", "17"])
def test_fixed_base_window_widths(gwin):
    """SIGOPS_GWIN: the positional fixed-base tables at other window widths than the default (the width is a run-time
    parameter of the kernels and of the table generator): unit ops on width-specific edge scalars and both kernel families
    end to end, in a child process because the tables are built once per context."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SIGOPS_GWIN=gwin)
    r = subprocess.run([sys.executable, "-c", _GWIN_CHILD, root], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ("ok " + gwin) in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_fixed_base_window_out_of_range_is_refused():
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, sys.argv[1]); import wgpu_sigops_b200 as w; lib = w.load(); "
            "rc = lib.sigops_init(None, 0); print('rc', rc, lib.sigops_last_error().decode())")
    for bad in ("3", "25", "abc"):
        r = subprocess.run([sys.executable, "-c", code, root], env=dict(os.environ, SIGOPS_GWIN=bad), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0 and "rc 0" not in r.stdout and "SIGOPS_GWIN" in r.stdout, r.stdout + r.stderr
