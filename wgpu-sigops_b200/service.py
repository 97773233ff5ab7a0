"""Streaming service mode (not part of the reference's API): a persistent ring of pinned request
slots on one device instead of one blocking call per batch.

The reference creates a wgpu device, allocates every buffer and blocks in `device.poll(Maintain::Wait)` on each call
(src/gpu.rs:5-35,129-170).  A `SigQueue` keeps `depth` slots alive -- pinned host arrays, device buffers, a stream and a
replayable CUDA graph each -- and lets `depth` requests be in flight at once, which is what multiplies the throughput of
small requests (they occupy only a fraction of each SM).

    q = SigQueue("secp256k1", max_batch=1024, depth=8)
    slot = 0
    q.sigs(slot)[:n] = ...; q.msgs(slot)[:n] = ...        # write straight into pinned memory
    q.submit(slot, n)                                      # asynchronous
    keys, status = q.wait(slot)                            # copies of the slot's results (n rows)

`device_index=-1` spreads the slots round-robin over every GPU of the pool.  `wait(slot, copy=False)` returns views of
the slot's pinned output arrays instead of copies: valid only until the slot is submitted again or the queue is closed
(close() frees the pinned memory -- the queue refuses to close while such views are handed out unless force=True).
"""
import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib

CURVES = {"secp256k1": 0, "secp256r1": 1, "ed25519": 2}


def _view(ptr: int, rows: int, width: int) -> np.ndarray:
    buf = (ctypes.c_uint8 * (rows * width)).from_address(ptr)
    a = np.frombuffer(buf, dtype=np.uint8)
    return a.reshape(rows, width) if width > 1 else a


class SigQueue:
    def __init__(self, curve: str, max_batch: int, depth: int = 4, device_index: int = 0):
        self.curve = curve
        self.cid = CURVES[curve]
        self.max_batch = int(max_batch)
        self.depth = int(depth)
        self._lib = _lib.load()
        self._q = ctypes.c_void_p()
        _lib.check(self._lib.sigops_queue_create(self.cid, device_index, self.max_batch, self.depth, ctypes.byref(self._q)))
        self._bufs = []
        self._views_out = 0
        for s in range(self.depth):
            p = [ctypes.c_void_p() for _ in range(5)]
            _lib.check(self._lib.sigops_queue_buffers(self._q, s, *[ctypes.byref(x) for x in p]))
            sigs = _view(p[0].value, self.max_batch, 64)
            msgs = _view(p[1].value, self.max_batch, 32)
            pks = _view(p[2].value, self.max_batch, 32) if p[2].value else None
            out = _view(p[3].value, self.max_batch, 1 if self.cid == 2 else 64)
            st = _view(p[4].value, self.max_batch, 1) if p[4].value else None
            self._bufs.append((sigs, msgs, pks, out, st))

    # pinned arrays of a slot (max_batch rows each); fill the first n rows before submit(slot, n)
    def sigs(self, slot: int) -> np.ndarray:
        return self._bufs[slot][0]

    def msgs(self, slot: int) -> np.ndarray:
        return self._bufs[slot][1]

    def pks(self, slot: int) -> Optional[np.ndarray]:
        return self._bufs[slot][2]

    def submit(self, slot: int, n: int) -> None:
        _lib.check(self._lib.sigops_queue_submit(self._q, slot, n))

    def done(self, slot: int) -> bool:
        d = ctypes.c_int(0)
        _lib.check(self._lib.sigops_queue_poll(self._q, slot, ctypes.byref(d)))
        return bool(d.value)

    def submit_device(self, slot: int, d_sigs: int, d_msgs: int, d_pks: Optional[int], n: int, src_device: int,
                      ready_event: Optional[int] = None) -> None:
        """Device-resident producer: the inputs are device pointers on CUDA device `src_device`; they reach the slot's
        device with cudaMemcpyPeerAsync (NVLink between peers).  ready_event: a cudaEvent_t handle to wait for, or None."""
        _lib.check(self._lib.sigops_queue_submit_device(self._q, slot, d_sigs, d_msgs, d_pks, n, src_device, ready_event))

    def slot_device(self, slot: int) -> int:
        """CUDA ordinal of the device the slot lives on."""
        return self._lib.sigops_queue_slot_device(self._q, slot)

    def wait(self, slot: int, copy: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """Blocks until the slot's request has completed.  Returns (keys n x 64, status n) for the recovery curves,
        (verdicts n, None) for ed25519 -- copies by default; with copy=False views of the slot's pinned arrays, valid only
        until the slot is submitted again or the queue is closed."""
        if not self._q:
            raise ValueError("queue is closed")
        n = ctypes.c_size_t(0)
        ms = ctypes.c_double(0)
        _lib.check(self._lib.sigops_queue_wait(self._q, slot, ctypes.byref(n), ctypes.byref(ms)))
        self.last_device_ms = ms.value
        out, st = self._bufs[slot][3], self._bufs[slot][4]
        out, st = out[: n.value], (st[: n.value] if st is not None else None)
        if copy:
            return out.copy(), (st.copy() if st is not None else None)
        self._views_out += 1
        return out, st

    def info(self) -> dict:
        c, d, dep = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        mb = ctypes.c_size_t()
        gl, gc = ctypes.c_uint64(), ctypes.c_uint64()
        _lib.check(self._lib.sigops_queue_info(self._q, ctypes.byref(c), ctypes.byref(d), ctypes.byref(mb), ctypes.byref(dep),
                                               ctypes.byref(gl), ctypes.byref(gc)))
        return {"curve": c.value, "device_index": d.value, "max_batch": mb.value, "depth": dep.value,
                "graph_launches": gl.value, "graph_captures": gc.value}

    def close(self, force: bool = False) -> None:
        """Destroys the queue and frees its pinned memory.  Views handed out by wait(copy=False) / sigs() / msgs() dangle
        afterwards, so closing with views outstanding needs force=True (the context manager and __del__ force)."""
        if self._q:
            if self._views_out and not force:
                raise RuntimeError("SigQueue.close(): %d result view(s) were handed out with copy=False; they would point "
                                   "at freed pinned memory -- drop them and pass force=True" % self._views_out)
            self._bufs = []
            self._lib.sigops_queue_destroy(self._q)
            self._q = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close(force=True)

    def __del__(self):
        try:
            self.close(force=True)
        except Exception:
            pass


def run_stream(curve: str, requests, max_batch: int, depth: int = 4, device_index: int = 0):
    """Convenience driver: pushes an iterable of requests -- (sigs, msgs) or (sigs, msgs, pks) arrays of at most
    max_batch rows -- through a queue with `depth` of them in flight, yielding copies of the results in request order."""
    with SigQueue(curve, max_batch, depth, device_index) as q:
        pending = []  # slots in submission order
        free = list(range(depth))

        def drain_one():
            slot = pending.pop(0)
            res = q.wait(slot)
            free.append(slot)
            return res

        for req in requests:
            if not free:
                yield drain_one()
            slot = free.pop(0)
            n = len(req[0])
            q.sigs(slot)[:n] = req[0]
            q.msgs(slot)[:n] = req[1]
            if q.cid == 2:
                q.pks(slot)[:n] = req[2]
            q.submit(slot, n)
            pending.append(slot)
        while pending:
            yield drain_one()
