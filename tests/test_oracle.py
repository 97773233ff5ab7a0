"""Pins the oracle (CPU, no GPU needed).

1. oracle/sigops_oracle.py against every golden vector / known-answer value the reference's own tests hold for the hot
   path, RFC 8032 vectors and OpenSSL 3 (via `cryptography`) for valid signatures on all three curves.
2. oracle/sigops_oracle.c (the fast twin used for 65k..1M batches and as the timed CPU baseline) against the Python
   oracle on every edge class and on random inputs.
3. the committed fixtures under tests/golden/ against both.

The reference itself (Rust) cannot be executed in this image, so rejecting outcomes are pinned only to the published
algorithms of the pinned third-party versions ("parity unpinned" for those classes; see the oracle header) and, as
an independent check, to OpenSSL's verdicts on the whole edge corpus (`test_openssl_pins_*`).
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import coracle
import sigops_oracle as o
import unit_checks as uc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BEAST = b"A beast can never be as cruel as a human being, so artistically, so picturesquely cruel."
BEAST_SIG = bytes.fromhex(
    "46ec716ae185a1d43b537e9ee45e7f178841c9457b5ede4ace9efb585b8ad59f"
    "0131dd08f04930d2771de52d2e6aa3f7d12da172ba8af87e963921cd7ed39182"
)


# ---------------------------------------------------------------------------------------- reference golden vectors
def test_k1_rfc6979_golden_vector():
    """src/curve_algos/secp256k1_ecdsa.rs:135-322: secret key 1, message hash, nonce words, signature hex,
    recovery id 0, recovered key = G (compressed 0279be66...)."""
    z = hashlib.sha256(BEAST).digest()
    assert z.hex() == "52840c5594968f39c0d7994330b5638405311580b6f9c1b8b3c1f04ca80db7c3"  # :300
    words = [16516427254913592388, 13430632917597294833, 11381083295555450127, 10383357845931004348]  # :147-152
    k = sum(w << (64 * i) for i, w in enumerate(words))
    sig = o.ecdsa_sign(o.K1, 1, z, k, low_s=True)
    assert sig == BEAST_SIG  # :211-214 (recovery id 0 -> parity bit clear, :289)
    g = o.K1.gx.to_bytes(32, "big") + o.K1.gy.to_bytes(32, "big")
    assert o.K1.gx.to_bytes(32, "big").hex() == "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"  # :272
    assert o.K1.gy % 2 == 0  # compressed prefix 02
    assert o.ecrecover_k1(sig, z) == g
    out, st = coracle.ecrecover(0, [sig], [z])
    assert st[0] == 0 and out[0].tobytes() == g


def test_strauss_shamir_corner_case_vector():
    """src/tests/secp256k1_curve.rs:691-741: x*G + y*B passes through the point at infinity at bit 254 (the
    reference's incomplete addition fails there; the test is commented out upstream)."""
    x = 0x8CE48A1B5F7942ED63C3F5380D98BD57F702AA6DED0E8022B4890762ACA5FA5D
    y = 0x84023F2E9587339FE4076DE927D8F1CBFF4279A6982E1B0599221E20153F147A
    B = (57955212013049338432744149260690748736552621582696778344469660993364486735760,
         18014696949887157897072847726343716132385694929890630512424732633979399864330)
    G = (o.K1.gx, o.K1.gy)
    assert (B[1] ** 2 - B[0] ** 3 - 7) % o.K1.p == 0
    # bit-serial Strauss-Shamir (the reference's `projective_strauss_shamir_mul`): at the 254th step (bit index 1) the
    # accumulator is the negative of the addend, so the sum is the point at infinity
    GB = o.sw_add(o.K1, G, B)
    acc, hits = None, []
    for i in range(255, -1, -1):
        acc = o.sw_add(o.K1, acc, acc)
        bx, by = (x >> i) & 1, (y >> i) & 1
        ad = GB if bx and by else G if bx else B if by else None
        if acc is not None and ad is not None and acc[0] == ad[0] and acc != ad:
            hits.append(i)
        acc = o.sw_add(o.K1, acc, ad)
    assert hits == [1]
    assert acc == o.sw_add(o.K1, o.sw_mul(o.K1, x, G), o.sw_mul(o.K1, y, B)) and acc is not None


def test_host_byte_order_vector():
    """src/tests/buffers.rs:15-29: the 32 big-endian bytes of p_k1 cast to little-endian u32 words."""
    words = np.frombuffer(o.K1.p.to_bytes(32, "big"), dtype="<u4")
    assert list(words) == [4294967295] * 6 + [4278190079, 805109759]


def test_glv_constants():
    """src/curve_algos/secp256k1_curve.rs:47-68 and the split check of src/curve_algos/secp256k1_mul.rs:38-94."""
    G = (o.K1.gx, o.K1.gy)
    assert o.sw_mul(o.K1, o.K1_LAMBDA, G) == (o.K1_BETA * o.K1.gx % o.K1.p, o.K1.gy)
    assert pow(o.K1_LAMBDA, 3, o.K1.n) == 1 and pow(o.K1_BETA, 3, o.K1.p) == 1
    a1, mb1, a2 = 0x3086D221A7D46BCDE86C90E49284EB15, 0xE4437ED6010E88286F547FA90ABFE4C3, 0x114CA50F7A8E2F3F657C1108D9D44CFD8
    g1 = 0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031
    g2 = 0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71
    assert (a1 - mb1 * o.K1_LAMBDA) % o.K1.n == 0 and (a2 + a1 * o.K1_LAMBDA) % o.K1.n == 0
    rng = random.Random(5)
    for _ in range(1000):
        k = rng.getrandbits(256) % o.K1.n
        c1, c2 = (k * g1 + (1 << 383)) >> 384, (k * g2 + (1 << 383)) >> 384
        k1, k2 = k - c1 * a1 - c2 * a2, c1 * mb1 - c2 * a1
        assert (k1 + k2 * o.K1_LAMBDA - k) % o.K1.n == 0 and abs(k1) < 2**128 and abs(k2) < 2**128


def test_shader_constants():
    """src/shader.rs:420-428,526-530 and src/tests/mod.rs:38-64."""
    assert (o.ED_P - 5) // 8 == 2**252 - 3
    assert (1 << 512) // o.ED_L == 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEB2106215D086329A7ED9CE5A30A2C131B
    assert o.ED_D2 == 16295367250680780974490674513165176452449235426866156013048779062215315747161
    assert o.R1.b == 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B


# ---------------------------------------------------------------------------------------- external known answers
RFC8032 = [  # RFC 8032 section 7.1, TEST 1..3 and TEST SHA(abc): (secret, public, message, signature)
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
     "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a", "",
     "e5564300c360ac729086e2cc806e828a84877f1eb8e5d974d873e06522490155"
     "5fb8821590a33bacc61e39701cf9b46bd25bf5f0595bbe24655141438e7a100b"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
     "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c", "72",
     "92a009a9f0d4cab8720e820b5f642540a2b27b5416503f8fb3762223ebdb69da"
     "085ac1e43e15996e458f3613d0f11d8c387b2eaeb4302aeeb00d291612bb0c00"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7",
     "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025", "af82",
     "6291d657deec24024827e69c3abe01a30ce548a284743a445e3680d7db5ac3ac"
     "18ff9b538d16f290ae67f760984dc6594a7c15e9716ed28dc027beceea1ec40a"),
]


@pytest.mark.parametrize("sk,pk,msg,sig", RFC8032)
def test_rfc8032_vectors(sk, pk, msg, sig):
    sk, pk, msg, sig = (bytes.fromhex(x) for x in (sk, pk, msg, sig))
    assert o.ed25519_expand(sk)[2] == pk
    assert o.ed25519_sign(sk, msg) == (sig, pk)
    assert o.ecverify_ed25519(sig, msg, pk)
    bad = bytearray(sig)
    bad[5] ^= 1
    assert not o.ecverify_ed25519(bytes(bad), msg, pk)


def test_openssl_cross_check():
    """Valid signatures made by OpenSSL 3: the oracle must recover the signer's key (one of the two parities) and
    accept the Ed25519 signatures; signatures made by the oracle's generator must verify under OpenSSL."""
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec, ed25519, utils

    for c, cid, curve in ((o.K1, 0, ec.SECP256K1()), (o.R1, 1, ec.SECP256R1())):
        for i in range(8):
            key = ec.generate_private_key(curve)
            z = hashlib.sha256(b"msg%d" % i).digest()
            r, s = utils.decode_dss_signature(key.sign(z, ec.ECDSA(utils.Prehashed(hashes.SHA256()))))
            pub = key.public_key().public_numbers()
            want = pub.x.to_bytes(32, "big") + pub.y.to_bytes(32, "big")
            if s >= 2**255:  # not encodable in the Fuel format: use the low-s twin
                s = c.n - s
            got = [o.ecrecover(c, o._ecdsa_sig_bytes(r, s, par), z) for par in (0, 1)]
            assert want in got
            cg = [coracle.ecrecover(cid, [o._ecdsa_sig_bytes(r, s, par)], [z])[0][0].tobytes() for par in (0, 1)]
            assert cg == [g if g is not None else bytes(64) for g in got]
        # generator output verifies under OpenSSL
        sigs, msgs, pks = coracle.gen_ecdsa(cid, 16, seed=99, low_s=(cid == 0))
        for sg, m, pk in zip(sigs, msgs, pks):
            sg, m, pk = sg.tobytes(), m.tobytes(), pk.tobytes()
            r = int.from_bytes(sg[:32], "big")
            s = int.from_bytes(bytes([sg[32] & 0x7F]) + sg[33:], "big")
            pub = ec.EllipticCurvePublicNumbers(int.from_bytes(pk[:32], "big"), int.from_bytes(pk[32:], "big"), curve).public_key()
            pub.verify(utils.encode_dss_signature(r, s), m, ec.ECDSA(utils.Prehashed(hashes.SHA256())))
    for i in range(8):
        key = ed25519.Ed25519PrivateKey.generate()
        msg = hashlib.sha256(b"ed%d" % i).digest()
        sig = key.sign(msg)
        pk = key.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
        assert o.ecverify_ed25519(sig, msg, pk)
        assert coracle.ecverify_ed25519([sig], [msg], [pk])[0] == 1
    sigs, msgs, pks = coracle.gen_ed25519(16, seed=99)
    for sg, m, pk in zip(sigs, msgs, pks):
        ed25519.Ed25519PublicKey.from_public_bytes(pk.tobytes()).verify(sg.tobytes(), m.tobytes())


def test_openssl_pins_ed25519_edge_classes():
    """Every Ed25519 edge class of the corpus (all the REJECTING outcomes, and the surprising accepts: small-order and
    non-canonical A with a crafted R, x = 0 with the sign bit set) against an independent production implementation:
    OpenSSL's Ed25519 verify is cofactorless, checks s < L, does not check y < p and compares R bytewise -- the
    decision procedure of ed25519-dalek 2.1.1 `verify`.  Oracle (Python and C) == OpenSSL on all of them."""
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives.asymmetric import ed25519

    def openssl_verify(sig, msg, pk):
        try:
            ed25519.Ed25519PublicKey.from_public_bytes(pk).verify(sig, msg)
            return True
        except InvalidSignature:
            return False

    corpus = o.ed25519_edge_cases()
    assert len(corpus) > 70
    n_reject = 0
    for name, sg, m, pk in corpus:
        want = openssl_verify(sg, m, pk)
        assert o.ecverify_ed25519(sg, m, pk) == want, name
        assert bool(coracle.ecverify_ed25519([sg], [m], [pk])[0]) == want, name
        if want:  # strict may only ever be stricter
            assert o.ecverify_ed25519(sg, m, pk)
        else:
            assert not o.ecverify_ed25519_strict(sg, m, pk), name
            n_reject += 1
    assert n_reject > 50
    # random corruptions of valid signatures (bit flips in R, s, A, M): 400 rows, same verdicts
    sigs, msgs, pks = coracle.gen_ed25519(400, seed=7)
    rng = random.Random(7)
    for i in range(400):
        arr = (sigs, sigs, pks, msgs)[i % 4]
        col = rng.randrange(32) + (32 if i % 4 == 1 else 0)
        arr[i, col] ^= 1 << rng.randrange(8)
    got = coracle.ecverify_ed25519(sigs, msgs, pks)
    for i in range(400):
        assert bool(got[i]) == openssl_verify(sigs[i].tobytes(), msgs[i].tobytes(), pks[i].tobytes()), i


def test_openssl_pins_ecdsa_edge_classes():
    """ECDSA edge corpus against OpenSSL: whenever the oracle recovers a key, (r, s) must VERIFY under that key in OpenSSL
    (high-s, z >= n, z = 0, R = +-G, the doubling corner cases included -- OpenSSL has no low-s rule either); whenever
    the oracle rejects for range reasons (r or s zero, r >= n) OpenSSL rejects the signature under any key (checked with
    G); a recovered key for the other parity is a different key that verifies too."""
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec, utils

    for c, cid, curve in ((o.K1, 0, ec.SECP256K1()), (o.R1, 1, ec.SECP256R1())):
        g_pub = ec.EllipticCurvePublicNumbers(c.gx, c.gy, curve).public_key()
        n_rec = n_rej = 0
        for name, sg, m in o.ecdsa_edge_cases(c):
            got = o.ecrecover(c, sg, m)
            cg, cst = coracle.ecrecover(cid, [sg], [m])
            assert (cst[0] == 0) == (got is not None), name
            r = int.from_bytes(sg[:32], "big")
            s = int.from_bytes(bytes([sg[32] & 0x7F]) + sg[33:], "big")
            if got is not None:
                assert cg[0].tobytes() == got, name
                pub = ec.EllipticCurvePublicNumbers(
                    int.from_bytes(got[:32], "big"), int.from_bytes(got[32:], "big"), curve).public_key()
                n_rec += 1
            else:
                pub = g_pub
                n_rej += 1
            try:
                pub.verify(utils.encode_dss_signature(r, s), m, ec.ECDSA(utils.Prehashed(hashes.SHA256())))
                ok = True
            except InvalidSignature:
                ok = False
            assert ok == (got is not None), name
        assert n_rec >= 10 and n_rej >= 8


# ---------------------------------------------------------------------------------------- C oracle == Python oracle
def test_c_oracle_sha512():
    rng = random.Random(1)
    for ln in (0, 1, 32, 64, 96, 111, 112, 127, 128, 200, 256):
        m = bytes(rng.getrandbits(8) for _ in range(ln))
        assert coracle.sha512(m) == hashlib.sha512(m).digest()


@pytest.mark.parametrize("cid", [0, 1])
def test_c_oracle_ecrecover_matches_python(cid):
    c = (o.K1, o.R1)[cid]
    cases = uc.ecdsa_cases(c, nvalid=64)
    out, st = coracle.ecrecover(cid, [x[1] for x in cases], [x[2] for x in cases])
    uc.check_ecrecover_against_oracle(c, cases, out, st)
    # random garbage: mostly rejected (non-residue r) -- both must agree row by row
    rng = random.Random(11 + cid)
    sigs = [bytes(rng.getrandbits(8) for _ in range(64)) for _ in range(64)]
    msgs = [bytes(rng.getrandbits(8) for _ in range(32)) for _ in range(64)]
    out, st = coracle.ecrecover(cid, sigs, msgs)
    for sg, m, ob, sb in zip(sigs, msgs, out, st):
        exp = o.ecrecover(c, sg, m)
        assert (sb == 1 and not ob.any()) if exp is None else (sb == 0 and ob.tobytes() == exp)
    # the generator's signatures recover to the signer's key under the Python oracle too
    sigs, msgs, pks = coracle.gen_ecdsa(cid, 24, seed=3, low_s=False)
    for sg, m, pk in zip(sigs, msgs, pks):
        assert o.ecrecover(c, sg.tobytes(), m.tobytes()) == pk.tobytes()


def test_c_oracle_ed25519_matches_python():
    cases = uc.ed_cases(nvalid=48)
    v = coracle.ecverify_ed25519([x[1] for x in cases], [x[2] for x in cases], [x[3] for x in cases])
    uc.check_ed_against_oracle(cases, v)
    assert 0 < int(v.sum()) < len(cases)  # both outcomes occur
    rng = random.Random(12)
    sigs, msgs, pks = coracle.gen_ed25519(48, seed=4)
    sigs, msgs, pks = sigs.copy(), msgs.copy(), pks.copy()
    for i in range(48):  # a single bit flip somewhere in two thirds of the rows
        if i % 3:
            arr = (sigs, msgs, pks)[i % 3 - 1] if i % 2 else sigs
            arr[i, rng.randrange(arr.shape[1])] ^= 1 << rng.randrange(8)
    v = coracle.ecverify_ed25519(sigs, msgs, pks)
    for i in range(48):
        assert bool(v[i]) == o.ecverify_ed25519(sigs[i].tobytes(), msgs[i].tobytes(), pks[i].tobytes()), i


def test_c_oracle_threads_agree():
    sigs, msgs, pks = coracle.gen_ecdsa(0, 300, seed=8)
    a = coracle.ecrecover(0, sigs, msgs, threads=1)
    b = coracle.ecrecover(0, sigs, msgs, threads=5)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[0] == pks).all()


# ---------------------------------------------------------------------------------------- committed fixtures
def test_golden_fixtures_match_oracles():
    """tests/golden/*.json were produced by tests/golden/make_golden.py (Python oracle); both oracles must still agree
    with them, so an accidental change of either oracle's semantics is caught."""
    for name, cid in (("secp256k1", 0), ("secp256r1", 1)):
        rows = json.load(open(os.path.join(GOLDEN, f"{name}_ecrecover.json")))["cases"]
        c = (o.K1, o.R1)[cid]
        sigs = [bytes.fromhex(r["sig"]) for r in rows]
        msgs = [bytes.fromhex(r["msg"]) for r in rows]
        out, st = coracle.ecrecover(cid, sigs, msgs)
        for r, sg, m, ob, sb in zip(rows, sigs, msgs, out, st):
            want = bytes.fromhex(r["pubkey"]) if r["pubkey"] else None
            assert o.ecrecover(c, sg, m) == want, r["label"]
            assert (sb == 0 and ob.tobytes() == want) if want else (sb == 1 and not ob.any()), r["label"]
    rows = json.load(open(os.path.join(GOLDEN, "ed25519_ecverify.json")))["cases"]
    v = coracle.ecverify_ed25519(*[[bytes.fromhex(r[k]) for r in rows] for k in ("sig", "msg", "pk")])
    for r, got in zip(rows, v):
        assert bool(got) == r["valid"], r["label"]
        assert o.ecverify_ed25519(bytes.fromhex(r["sig"]), bytes.fromhex(r["msg"]), bytes.fromhex(r["pk"])) == r["valid"]


def test_precompute_bases_golden():
    """precompute::*_bases at log_limb_size = 13 (src/precompute.rs:36-69): sizes 640 / 640 / 960, entry 0 is G in
    Montgomery form with R = 2^260 (src/tests/mod.rs:134-149)."""
    g = json.load(open(os.path.join(GOLDEN, "precompute_bases_13.json")))
    for curve, size in (("secp256k1", 640), ("secp256r1", 640), ("ed25519", 960)):
        limbs = o.precompute_bases(curve, 13)
        assert len(limbs) == size and limbs == g[curve]
        assert all(0 <= x < (1 << 13) for x in limbs)
    limbs = o.precompute_bases("secp256k1", 13)
    x = sum(v << (13 * i) for i, v in enumerate(limbs[:20]))
    assert x == o.K1.gx * (1 << 260) % o.K1.p


def test_fast_k1_port_is_pinned_to_the_checker():
    """oracle/k1_fast.c (the timed CPU baseline: fold field, GLV, wNAF) against the checker oracle_ecrecover on random valid
    signatures (low-s and high-s), the whole secp256k1 edge corpus, single-bit corruptions and mostly-invalid fuzz rows."""
    import batches
    import fuzz_cases

    s, m, pk, st, n_edge = batches.ecdsa_batch(0, 6000, edge_every=3, seed=77, mix_high_s=True)
    o1, s1 = coracle.k1_ecrecover_fast(s, m, threads=2)
    assert n_edge > 1900 and st.sum() > 300
    assert (o1 == pk).all() and (s1 == st).all()
    rows = [(x[1], x[2]) for x in uc.ecdsa_cases(o.K1, nvalid=50)]
    sig = np.array([np.frombuffer(a, dtype=np.uint8) for a, _ in rows])
    msg = np.array([np.frombuffer(b, dtype=np.uint8) for _, b in rows])
    e_out, e_st = coracle.ecrecover(0, sig, msg)
    f_out, f_st = coracle.k1_ecrecover_fast(sig, msg, threads=1)
    assert (e_out == f_out).all() and (e_st == f_st).all()
    fs, fm = fuzz_cases.ecdsa_batch(0, 8000, seed=5)
    e_out, e_st = coracle.ecrecover(0, fs, fm)
    f_out, f_st = coracle.k1_ecrecover_fast(fs, fm, threads=2)
    assert 1000 < e_st.sum() < 7000
    assert (e_out == f_out).all() and (e_st == f_st).all()


def test_openssl_native_legs_agree_with_the_oracle():
    """oracle/openssl_ref.c (bench.py's OpenSSL baseline legs): every key the oracle recovers verifies, a corrupted message
    does not; Ed25519 verdicts equal the oracle's on the edge corpus (OpenSSL is cofactorless with the same s < L rule)."""
    import batches

    if coracle.openssl_ref() is None:
        pytest.skip("libcrypto not linkable")
    for cid in (0, 1):
        s, m, pk, st, _ = batches.ecdsa_batch(cid, 600, edge_every=0, seed=91 + cid, mix_high_s=(cid == 1))
        assert coracle.openssl_ecdsa_verify(cid, s, m, pk, threads=2).all()
        m2 = m.copy()
        m2[:, 5] ^= 1
        assert not coracle.openssl_ecdsa_verify(cid, s, m2, pk, threads=2).any()
    s, m, pk, v, _ = batches.ed25519_batch(900, edge_every=2, seed=93)
    assert (coracle.openssl_ed25519_verify(s, m, pk, threads=2) == v).all()
