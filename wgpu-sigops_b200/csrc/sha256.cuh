// SHA-256 for the steps on either side of public-key recovery: the message prehash
// `fuel_crypto::Message::new(bytes)` = SHA-256(bytes) that the reference's callers compute on the host before every call
// (src/tests/secp256k1_ecdsa.rs:21-22, src/benchmarks/secp256k1_ecdsa.rs:160-165) and the Fuel address
// `PublicKey::hash()` = SHA-256(X || Y) they derive from the recovered key afterwards.  Not part of the reference's
// device code; provided so that a block of transactions can go raw bytes -> addresses without a host pass.
#pragma once
#include "field.cuh"

namespace sigops {

#define SG_SHA256_K                                                                                              \
    {                                                                                                            \
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,  \
            0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u,           \
            0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau,           \
            0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u,           \
            0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u,           \
            0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u,           \
            0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,           \
            0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu,           \
            0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u            \
    }

#if defined(__CUDACC__)
static __device__ const u32 sha256_k_dev[64] = SG_SHA256_K;
#endif
static const u32 sha256_k_host[64] = SG_SHA256_K;

SG_HD u32 sha256_k(int i) {
#if SG_PTX
    return sha256_k_dev[i];
#else
    return sha256_k_host[i];
#endif
}

SG_HD u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }

#define SG_SHA256_IV \
    { 0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u }

// one compression with a rolling 16-word schedule; w is consumed
SG_HD void sha256_compress(u32* st, u32* w) {
    u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int t0 = 0; t0 < 64; t0 += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (t0 > 0) {
                u32 w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
            }
            u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            u32 ch = (e & f) ^ (~e & g);
            u32 t1 = h + S1 + ch + sha256_k(t0 + j) + w[j];
            u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            u32 maj = (a & b) ^ (a & c) ^ (b & c);
            u32 t2 = S0 + maj;
            h = g;
            g = f;
            f = e;
            e = d + t1;
            d = c;
            c = b;
            b = a;
            a = t1 + t2;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// digest as 8 little-endian-loaded words of the 32 big-endian digest bytes
SG_HD void sha256_digest_words(u32* out_w, const u32* st) {
#pragma unroll
    for (int i = 0; i < 8; i++) out_w[i] = bswap32(st[i]);
}

// SHA-256 of exactly 64 bytes given as 16 LE-loaded words (a recovered public key X || Y)
SG_HD void sha256_64(u32* out_w, const u32* in_w) {
    u32 st[8] = SG_SHA256_IV;
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = bswap32(in_w[i]);
    sha256_compress(st, w);
    w[0] = 0x80000000u;
#pragma unroll
    for (int i = 1; i < 15; i++) w[i] = 0;
    w[15] = 512;
    sha256_compress(st, w);
    sha256_digest_words(out_w, st);
}

// SHA-256 of a message of any length read bytewise from memory
SG_HD void sha256_ram(u32* out_w, const uint8_t* msg, size_t len) {
    u32 st[8] = SG_SHA256_IV;
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    size_t mi = 0;
    bool pad_done = false, len_done = false;
    while (!len_done) {
        int idx = 0;
        while (idx < 64 && mi < len) {
            w[idx >> 2] |= (u32)msg[mi] << (24 - 8 * (idx & 3));
            idx++;
            mi++;
        }
        if (idx < 64 && !pad_done) {
            w[idx >> 2] |= 0x80u << (24 - 8 * (idx & 3));
            idx++;
            pad_done = true;
        }
        if (pad_done && idx <= 56) {
            w[14] = (u32)(((u64)len << 3) >> 32);
            w[15] = (u32)((u64)len << 3);
            len_done = true;
        }
        sha256_compress(st, w);
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
    }
    sha256_digest_words(out_w, st);
}

}  // namespace sigops
