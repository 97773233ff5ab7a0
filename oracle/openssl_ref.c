/*
 * OpenSSL 3 (libcrypto) verify legs, multi-threaded -- BENCHMARK INFRASTRUCTURE ONLY (never part of the product).
 *
 * an independent, production-grade CPU reference point on the same inputs and the same core
 * count as the port: ECDSA *verify* (not recover) on secp256k1 / P-256 with the signer's key, and Ed25519 verify.  The
 * reference's own CPU column is fuel-crypto / ed25519-dalek (src/benchmarks/secp256k1_ecdsa.rs:117-122), which cannot be
 * built in this image; these numbers are labelled "verify, not recover; OpenSSL, not fuel-crypto" wherever they appear.
 */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/ecdsa.h>
#include <openssl/evp.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int curve;
    const uint8_t *sigs, *msgs, *pks;
    uint8_t* ok;
    size_t lo, hi;
} job;

static void* ecdsa_worker(void* p) {
    job* j = p;
    EC_KEY* key = EC_KEY_new_by_curve_name(j->curve == 0 ? NID_secp256k1 : NID_X9_62_prime256v1);
    const EC_GROUP* grp = EC_KEY_get0_group(key);
    EC_POINT* pt = EC_POINT_new(grp);
    BN_CTX* ctx = BN_CTX_new();
    for (size_t i = j->lo; i < j->hi; i++) {
        uint8_t oct[65], sb[32];
        oct[0] = 4;
        memcpy(oct + 1, j->pks + 64 * i, 64);
        memcpy(sb, j->sigs + 64 * i + 32, 32);
        sb[0] &= 0x7f; /* Fuel encoding: the y parity of R rides in bit 255 of s (src/wgsl/signature.wgsl:6-21) */
        int v = 0;
        if (EC_POINT_oct2point(grp, pt, oct, 65, ctx) == 1 && EC_KEY_set_public_key(key, pt) == 1) {
            ECDSA_SIG* sig = ECDSA_SIG_new();
            BIGNUM* r = BN_bin2bn(j->sigs + 64 * i, 32, NULL);
            BIGNUM* s = BN_bin2bn(sb, 32, NULL);
            ECDSA_SIG_set0(sig, r, s);
            v = ECDSA_do_verify(j->msgs + 32 * i, 32, sig, key) == 1;
            ECDSA_SIG_free(sig);
        }
        j->ok[i] = (uint8_t)v;
    }
    BN_CTX_free(ctx);
    EC_POINT_free(pt);
    EC_KEY_free(key);
    return NULL;
}

static void* ed_worker(void* p) {
    job* j = p;
    EVP_MD_CTX* ctx = EVP_MD_CTX_new();
    for (size_t i = j->lo; i < j->hi; i++) {
        int v = 0;
        EVP_PKEY* pk = EVP_PKEY_new_raw_public_key(EVP_PKEY_ED25519, NULL, j->pks + 32 * i, 32);
        if (pk) {
            EVP_MD_CTX_reset(ctx);
            if (EVP_DigestVerifyInit(ctx, NULL, NULL, NULL, pk) == 1)
                v = EVP_DigestVerify(ctx, j->sigs + 64 * i, 64, j->msgs + 32 * i, 32) == 1;
            EVP_PKEY_free(pk);
        }
        j->ok[i] = (uint8_t)v;
    }
    EVP_MD_CTX_free(ctx);
    return NULL;
}

static int run(void* (*fn)(void*), job proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    job* jobs = malloc(sizeof(job) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].lo = n * t / threads;
        jobs[t].hi = n * (t + 1) / threads;
        if (t > 0) pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    fn(&jobs[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}

/* curve 0 = secp256k1, 1 = P-256; sigs Fuel-encoded (n x 64), msgs n x 32 prehashes, pks n x 64 (X || Y); ok[i] = verified */
int openssl_ecdsa_verify(int curve, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* ok, int threads) {
    job j = {curve, sigs, msgs, pks, ok, 0, 0};
    return run(ecdsa_worker, j, n, threads);
}

/* sigs n x 64, msgs n x 32, pks n x 32 */
int openssl_ed25519_verify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* ok, int threads) {
    job j = {2, sigs, msgs, pks, ok, 0, 0};
    return run(ed_worker, j, n, threads);
}
