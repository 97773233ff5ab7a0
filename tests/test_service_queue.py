"""Streaming service mode: `sigops_queue_*` / `service.SigQueue`.

CPU: argument validation and the no-device behaviour (no CPU fallback).  GPU: every request that goes through a queue --
several in flight, ragged sizes, slots reused, graph replay and re-capture, several driver threads -- is bit-exact against
the C oracle, exactly like the blocking entry points."""
import ctypes
import threading

import numpy as np
import pytest

import batches
import coracle


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_queue_argument_checks_without_device(sigops):
    lib = sigops.load()
    q = ctypes.c_void_p()
    for args in ((3, 0, 64, 2), (-1, 0, 64, 2), (0, 0, 0, 2), (0, 0, 64, 0), (0, 0, 64, 65), (0, 0, (1 << 24) + 1, 1)):
        assert lib.sigops_queue_create(*args, ctypes.byref(q)) != 0, args
        assert not q.value
        assert b"bad arguments" in lib.sigops_last_error()
    assert lib.sigops_queue_create(0, 0, 64, 2, None) != 0
    # NULL queue: every entry point refuses instead of crashing; destroy(NULL) is a no-op
    d = ctypes.c_int()
    assert lib.sigops_queue_submit(None, 0, 1) != 0
    assert lib.sigops_queue_poll(None, 0, ctypes.byref(d)) != 0
    assert lib.sigops_queue_wait(None, 0, None, None) != 0
    assert lib.sigops_queue_buffers(None, 0, None, None, None, None, None) != 0
    assert lib.sigops_queue_info(None, None, None, None, None, None, None) != 0
    assert lib.sigops_queue_destroy(None) == 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_queue_has_no_cpu_fallback(sigops):
    with pytest.raises(sigops.ShaderFailureError) as e:
        sigops.service.SigQueue("secp256k1", 64, 2)
    assert "CUDA" in str(e.value)


def _requests(curve, sizes, seed):
    """One big batch with edge rows, cut into requests of the given sizes."""
    total = int(sum(sizes))
    if curve == "ed25519":
        sigs, msgs, pks, want, _ = batches.ed25519_batch(total, edge_every=5, seed=seed)
        cols = (sigs, msgs, pks)
        expect = (want, None)
    else:
        cid = 0 if curve == "secp256k1" else 1
        sigs, msgs, keys, st, _ = batches.ecdsa_batch(cid, total, edge_every=7, seed=seed, mix_high_s=True)
        cols = (sigs, msgs)
        expect = (keys, st)
    reqs, exp, lo = [], [], 0
    for n in sizes:
        reqs.append(tuple(c[lo:lo + n] for c in cols))
        exp.append(tuple(None if e is None else e[lo:lo + n] for e in expect))
        lo += n
    return reqs, exp


def _check(curve, got, want):
    out, st = got
    w_out, w_st = want
    assert out.shape[0] == w_out.shape[0]
    assert np.array_equal(out.reshape(w_out.shape), w_out)
    if curve != "ed25519":
        assert np.array_equal(st, w_st)


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ["secp256k1", "secp256r1", "ed25519"])
def test_queue_matches_oracle_many_in_flight(sigops, curve):
    sizes = [1, 300, 0, 64, 300, 300, 33, 257, 300, 300, 300, 5, 300, 128, 300, 300]
    reqs, exp = _requests(curve, sizes, seed=0x51600100)
    got = list(sigops.service.run_stream(curve, reqs, max_batch=300, depth=4))
    assert len(got) == len(reqs)
    for g, w in zip(got, exp):
        _check(curve, g, w)


@pytest.mark.gpu
def test_queue_slot_protocol_and_graph_replay(sigops):
    sizes = [200] * 12 + [77] * 4 + [200] * 4
    reqs, exp = _requests("secp256k1", sizes, seed=0x51600101)
    launches0 = sigops.load().sigops_kernel_launches()
    with sigops.service.SigQueue("secp256k1", 256, depth=2) as q:
        assert q.info()["depth"] == 2 and q.info()["max_batch"] == 256
        assert q.pks(0) is None
        for i, (sg, m) in enumerate(reqs):
            slot = i % 2
            if i >= 2:
                _check("secp256k1", q.wait(slot), exp[i - 2])
                assert q.last_device_ms > 0
            assert q.done(slot)
            q.sigs(slot)[: len(sg)] = sg
            q.msgs(slot)[: len(sg)] = m
            q.submit(slot, len(sg))
            with pytest.raises(sigops.ShaderFailureError):  # a slot in flight cannot be submitted again
                q.submit(slot, len(sg))
        for i in (len(reqs) - 2, len(reqs) - 1):
            _check("secp256k1", q.wait(i % 2), exp[i])
        with pytest.raises(sigops.ShaderFailureError):
            q.submit(0, 257)  # larger than max_batch
        info = q.info()
        # one graph launch per request; captures only when a slot's request size changes (200 -> 77 -> 200 per slot)
        assert info["graph_launches"] == len(reqs)
        assert info["graph_captures"] == 6
    assert sigops.load().sigops_kernel_launches() - launches0 == len(reqs)
    # shutdown is refused while a queue is alive, allowed afterwards
    lib = sigops.load()
    q = sigops.service.SigQueue("ed25519", 32, depth=1)
    assert lib.sigops_shutdown() != 0
    q.close()


@pytest.mark.gpu
def test_queue_driven_by_several_threads(sigops):
    """Two threads drive disjoint slots of one queue while a third makes blocking calls: all results exact."""
    sizes = [500] * 16
    reqs, exp = _requests("ed25519", sizes, seed=0x51600102)
    errors = []
    with sigops.service.SigQueue("ed25519", 512, depth=4) as q:
        def drive(slots, idxs):
            try:
                for k, i in enumerate(idxs):
                    slot = slots[k % len(slots)]
                    sg, m, pk = reqs[i]
                    q.wait(slot)
                    q.sigs(slot)[:500], q.msgs(slot)[:500], q.pks(slot)[:500] = sg, m, pk
                    q.submit(slot, 500)
                    out, _ = q.wait(slot)
                    if not np.array_equal(out.reshape(-1), exp[i][0]):
                        errors.append(i)
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        def blocking():
            try:
                sigs, msgs, keys, st, _ = batches.ecdsa_batch(0, 3000, edge_every=9, seed=5)
                for _ in range(3):
                    out, s2 = sigops.secp256k1_ecdsa.ecrecover_with_status(sigs, msgs)
                    if not (np.array_equal(out, keys) and np.array_equal(s2, st)):
                        errors.append("blocking")
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        ts = [threading.Thread(target=drive, args=([0, 1], range(0, 8))),
              threading.Thread(target=drive, args=([2, 3], range(8, 16))),
              threading.Thread(target=blocking)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    assert not errors, errors


@pytest.mark.gpu
def test_queues_on_two_devices(sigops):
    """One queue per device of the pool, driven alternately: independent shards, no exchange between devices."""
    lib = sigops.load()
    if lib.sigops_num_devices() < 2:
        pytest.skip("one visible GPU")
    sizes = [200] * 8
    reqs, exp = _requests("secp256r1", sizes, seed=0x51600103)
    with sigops.service.SigQueue("secp256r1", 200, depth=2, device_index=0) as q0, \
            sigops.service.SigQueue("secp256r1", 200, depth=2, device_index=1) as q1:
        assert q0.info()["device_index"] == 0 and q1.info()["device_index"] == 1
        qs = (q0, q1)
        for i, (sg, m) in enumerate(reqs):
            q, slot = qs[i % 2], (i // 2) % 2
            q.wait(slot)
            q.sigs(slot)[:200], q.msgs(slot)[:200] = sg, m
            q.submit(slot, 200)
            out, st = q.wait(slot)
            _check("secp256r1", (out, st), exp[i])
    with pytest.raises(sigops.ShaderFailureError):
        sigops.service.SigQueue("secp256k1", 64, 1, device_index=lib.sigops_num_devices())
