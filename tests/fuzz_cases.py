"""Builders of mostly-invalid input batches (random bytes, values hugging 0 / n / p / 2^255) for differential fuzzing of
the CUDA path against the C oracle.  Used by tests/test_gpu_fuzz.py (fixed seeds) and tools/fuzz_soak.py (many seeds).
Test infrastructure."""
import numpy as np

import coracle
import sigops_oracle as o


def _be(x):
    return np.frombuffer(int(x).to_bytes(32, "big"), dtype=np.uint8)


def ecdsa_batch(cid, n, seed):
    """(sigs, msgs): a quarter random bytes, a quarter valid r with random s / parity, a quarter special values in r, s
    or z, a quarter valid (high-s allowed)."""
    c = (o.K1, o.R1)[cid]
    rng = np.random.default_rng(100 + cid + 1000 * seed)
    sigs, msgs, _ = coracle.gen_ecdsa(cid, n, seed=900 + cid + 1000 * seed, low_s=False)
    sigs, msgs = sigs.copy(), msgs.copy()
    q = n // 4
    sigs[:q] = rng.integers(0, 256, size=(q, 64), dtype=np.uint8)  # random r (half of them off the curve), s, parity
    msgs[:q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    sigs[q:2 * q, 32:] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # valid r, random s / parity
    special = [0, 1, 2, 3, c.n - 2, c.n - 1, c.n, c.n + 1, c.p - 1, c.p, c.p + 1, 2**255 - 1, 2**255, 2**256 - 1,
               2**128, 2**224, 2**192 + 2**96, c.n // 2, c.n // 2 + 1, 7, c.gx]
    nsp = min(q // 2, 6000)
    for i in range(2 * q, 2 * q + nsp):  # special values in r, s or z
        v = special[i % len(special)]
        where = (i // len(special)) % 3
        if where == 0:
            sigs[i, :32] = _be(v)
        elif where == 1:
            sigs[i, 32:] = _be(v % 2**255)
            sigs[i, 32] |= (i & 1) << 7
        else:
            msgs[i] = _be(v)
    if q > nsp:  # the rest of that quarter: a valid signature with one flipped bit in r, s or z
        idx = np.arange(2 * q + nsp, 3 * q)
        bit = rng.integers(0, 8, size=len(idx)).astype(np.uint8)
        byte = rng.integers(0, 96, size=len(idx))
        for arr, lo in ((sigs, 0), (msgs, 64)):
            width = 64 if lo == 0 else 32
            sel = (byte >= lo) & (byte < lo + width)
            arr[idx[sel], byte[sel] - lo] ^= (np.uint8(1) << bit[sel])
    return sigs, msgs


def ed25519_batch(n, seed):
    """(sigs, msgs, pks): a quarter random bytes, a quarter valid signatures under random keys, a quarter random s, then
    special encodings in the key / R / s, the rest valid."""
    rng = np.random.default_rng(200 + 1000 * seed)
    sigs, msgs, pks = coracle.gen_ed25519(n, seed=910 + 1000 * seed)
    sigs, msgs, pks = sigs.copy(), msgs.copy(), pks.copy()
    q = n // 4
    sigs[:q] = rng.integers(0, 256, size=(q, 64), dtype=np.uint8)
    pks[:q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    pks[q:2 * q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # valid signature under a random (maybe invalid) key
    sigs[2 * q:3 * q, 32:] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # random s (mostly non-canonical)
    sigs[2 * q:3 * q:2, 63] &= 0x0F  # ... half of them canonical-range
    special = [0, 1, o.ED_P - 1, o.ED_P, o.ED_P + 1, 2**255 - 1, 2**255 - 19 + 2**255, 2**256 - 1, o.ED_L, o.ED_L - 1]
    for i in range(3 * q, 3 * q + min(q // 2, 3000)):  # special y in the key / R, special s
        v = special[i % len(special)]
        le = np.frombuffer(int(v % 2**256).to_bytes(32, "little"), dtype=np.uint8)
        where = (i // len(special)) % 3
        if where == 0:
            pks[i] = le
        elif where == 1:
            sigs[i, :32] = le
        else:
            sigs[i, 32:] = le
    return sigs, msgs, pks
