"""Cases for the variable-length / strict ed25519 entry point: (label, sig, msg, pk) with messages of many lengths
(block-boundary lengths of SHA-512 included) plus the whole non-strict edge corpus re-judged under strict rules."""
import random

import coracle
import sigops_oracle as o

LENGTHS = [0, 1, 2, 31, 32, 33, 47, 48, 49, 63, 64, 65, 111, 112, 127, 128, 129, 175, 176, 177, 191, 192, 193, 255, 256, 1000, 4096]


def cases(seed=9):
    rng = random.Random(seed)
    out = []
    for ln in LENGTHS:
        sk = bytes(rng.getrandbits(8) for _ in range(32))
        msg = bytes(rng.getrandbits(8) for _ in range(ln))
        sig, pk = coracle.ed25519_sign(sk, msg)
        out.append((f"valid_len{ln}", sig, msg, pk))
        if ln:
            bad = bytearray(msg)
            bad[rng.randrange(ln)] ^= 1 << rng.randrange(8)
            out.append((f"flip_msg_len{ln}", sig, bytes(bad), pk))
        out.append((f"trunc_len{ln}", sig, msg[:-1] if ln else b"\x00", pk))
    # the 32-byte-message edge corpus (small-order A, non-canonical A / R, mixed order, ...): strict flips some verdicts
    out += o.ed25519_edge_cases()
    return out
