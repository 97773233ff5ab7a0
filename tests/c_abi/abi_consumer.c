/* A compiled C99 consumer of include/sigops.h (TEST INFRASTRUCTURE).
 *
 * Stands in for the Rust shim crate (the files under rust/src/), which cannot be compiled in this image: it makes every call the shim
 * makes, with the buffers a Rust caller has -- plain malloc'ed (pageable) memory -- and compares the results with the
 * golden fixtures of tests/golden/ (flattened to binary by tests/test_c_abi.py).  The signatures it binds are the ones
 * the reference exposes at src/secp256k1_ecdsa.rs:61-66,203-212, src/secp256r1_ecdsa.rs:62-67, src/ed25519_eddsa.rs:67-73
 * and src/precompute.rs:36-69.
 *
 *   abi_consumer <fixture_dir> host     host-only entry points + "no device => every compute call fails loudly"
 *   abi_consumer <fixture_dir> gpu      everything, on the visible CUDA devices
 *
 * Built with: gcc -std=c99 -Wall -Wextra -Werror -pedantic abi_consumer.c -L<libdir> -lsigops
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sigops.h"

static int failures = 0;
#define CHECK(cond, ...)                           \
    do {                                           \
        if (!(cond)) {                             \
            failures++;                            \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);          \
            fprintf(stderr, "\n");                 \
        }                                          \
    } while (0)

typedef struct {
    size_t n;
    uint8_t *sigs, *msgs, *pks, *expect; /* expect: status (ecdsa: 1 = rejected) or verdict (ed25519) per row */
    size_t pk_bytes;
} fixture;

static int load_fixture(const char* dir, const char* name, size_t pk_bytes, fixture* f) {
    char path[1024];
    uint32_t n = 0;
    size_t i;
    FILE* fp;
    snprintf(path, sizeof path, "%s/%s.bin", dir, name);
    fp = fopen(path, "rb");
    if (!fp) return 1;
    if (fread(&n, 4, 1, fp) != 1) return 1;
    f->n = n;
    f->pk_bytes = pk_bytes;
    f->sigs = malloc((size_t)n * 64);
    f->msgs = malloc((size_t)n * 32);
    f->pks = malloc((size_t)n * pk_bytes);
    f->expect = malloc(n);
    for (i = 0; i < n; i++) {
        if (fread(f->sigs + i * 64, 64, 1, fp) != 1 || fread(f->msgs + i * 32, 32, 1, fp) != 1 ||
            fread(f->pks + i * pk_bytes, pk_bytes, 1, fp) != 1 || fread(f->expect + i, 1, 1, fp) != 1)
            return 1;
    }
    fclose(fp);
    return 0;
}

static int load_u32(const char* dir, const char* name, uint32_t** out, size_t* len) {
    char path[1024];
    uint32_t n = 0;
    FILE* fp;
    snprintf(path, sizeof path, "%s/%s.bin", dir, name);
    fp = fopen(path, "rb");
    if (!fp) return 1;
    if (fread(&n, 4, 1, fp) != 1) return 1;
    *out = malloc((size_t)n * 4);
    *len = n;
    if (fread(*out, 4, n, fp) != n) return 1;
    fclose(fp);
    return 0;
}

typedef int (*recover_fn)(const uint8_t*, const uint8_t*, size_t, uint8_t*, uint8_t*);

/* precompute::*_bases through the capacity protocol the Rust shim uses (rust/src/precompute.rs) */
static void check_bases(const char* dir) {
    static const char* names[3] = {"bases_secp256k1", "bases_secp256r1", "bases_ed25519"};
    int c;
    for (c = 0; c < 3; c++) {
        uint32_t *want = NULL, *got;
        size_t want_len = 0, len = 0, small;
        CHECK(load_u32(dir, names[c], &want, &want_len) == 0, "fixture %s", names[c]);
        if (!want) continue;
        CHECK(sigops_precompute_bases(c, 13, NULL, &len) == 0, "length query curve %d", c);
        CHECK(len == want_len && len == (c == 2 ? 960u : 640u), "curve %d: %lu limbs", c, (unsigned long)len);
        got = malloc(len * 4);
        small = len - 1; /* wrong-length table: must fail and report the length it needs */
        CHECK(sigops_precompute_bases(c, 13, got, &small) != 0 && small == len, "short buffer must be refused");
        CHECK(sigops_precompute_bases(c, 13, got, &len) == 0, "bases curve %d", c);
        CHECK(memcmp(got, want, len * 4) == 0, "bases curve %d differ from the golden table", c);
        free(got);
        free(want);
    }
    {
        size_t len = 0;
        CHECK(sigops_precompute_bases(0, 16, NULL, &len) != 0, "log_limb_size 16 must be refused (mont.wgsl:12,37)");
        CHECK(sigops_precompute_bases(0, 10, NULL, &len) != 0, "log_limb_size 10 must be refused");
        CHECK(sigops_precompute_bases(3, 13, NULL, &len) != 0, "unknown curve must be refused");
    }
}

static void check_plan(void) {
    size_t b[9];
    int used = 0, g;
    CHECK(sigops_plan_shards(1048576, 8, b, &used) == 0 && used == 8, "plan 1M over 8");
    for (g = 0; g <= 8; g++) CHECK(b[g] == (size_t)g * 131072, "bound %d", g);
    CHECK(sigops_plan_shards(100, 8, b, &used) == 0 && used == 1 && b[1] == 100, "small batches stay on one device");
    CHECK(sigops_plan_shards(100, 0, b, &used) != 0, "n_devices 0 is an error");
}

static void run_ecdsa(const char* what, recover_fn fn, const fixture* f) {
    const size_t big = 100003; /* not a multiple of anything: ragged tail, several waves on one device */
    uint8_t *out, *st, *ts, *tm;
    size_t i;
    /* n == 0: Ok(vec![]) and nothing is touched (src/secp256k1_ecdsa.rs:71-73) -- even with NULL buffers */
    CHECK(fn(NULL, NULL, 0, NULL, NULL) == 0, "%s n=0", what);
    /* n == 1, status not requested (out_status == NULL is allowed) */
    out = malloc(64);
    memset(out, 0xAA, 64);
    CHECK(fn(f->sigs, f->msgs, 1, out, NULL) == 0, "%s n=1: %s", what, sigops_last_error());
    CHECK(memcmp(out, f->pks, 64) == 0, "%s n=1 key", what);
    free(out);
    /* the whole fixture, with status */
    out = malloc(f->n * 64);
    st = malloc(f->n);
    memset(st, 0xAA, f->n);
    CHECK(fn(f->sigs, f->msgs, f->n, out, st) == 0, "%s fixture: %s", what, sigops_last_error());
    for (i = 0; i < f->n; i++) {
        CHECK(st[i] == f->expect[i], "%s row %lu status %d want %d", what, (unsigned long)i, st[i], f->expect[i]);
        CHECK(memcmp(out + i * 64, f->pks + i * 64, 64) == 0, "%s row %lu key", what, (unsigned long)i);
    }
    free(out);
    free(st);
    /* 100,003 rows: the fixture tiled, pageable buffers, status NULL */
    ts = malloc(big * 64);
    tm = malloc(big * 32);
    out = malloc(big * 64);
    for (i = 0; i < big; i++) {
        memcpy(ts + i * 64, f->sigs + (i % f->n) * 64, 64);
        memcpy(tm + i * 32, f->msgs + (i % f->n) * 32, 32);
    }
    CHECK(fn(ts, tm, big, out, NULL) == 0, "%s n=100003: %s", what, sigops_last_error());
    for (i = 0; i < big; i++)
        if (memcmp(out + i * 64, f->pks + (i % f->n) * 64, 64) != 0) {
            CHECK(0, "%s n=100003 row %lu", what, (unsigned long)i);
            break;
        }
    /* null buffers with n > 0 are an error, not a crash */
    CHECK(fn(NULL, tm, 4, out, NULL) != 0, "%s NULL sigs must fail", what);
    free(ts);
    free(tm);
    free(out);
}

static void run_ed(const fixture* f) {
    const size_t big = 100003;
    uint8_t *v, *ts, *tm, *tk;
    size_t i;
    CHECK(sigops_ed25519_ecverify(NULL, NULL, NULL, 0, NULL) == 0, "ed25519 n=0");
    v = malloc(f->n);
    memset(v, 0xAA, f->n);
    CHECK(sigops_ed25519_ecverify(f->sigs, f->msgs, f->pks, f->n, v) == 0, "ed25519 fixture: %s", sigops_last_error());
    for (i = 0; i < f->n; i++) CHECK(v[i] == f->expect[i], "ed25519 row %lu verdict %d want %d", (unsigned long)i, v[i], f->expect[i]);
    free(v);
    ts = malloc(big * 64);
    tm = malloc(big * 32);
    tk = malloc(big * 32);
    v = malloc(big);
    for (i = 0; i < big; i++) {
        memcpy(ts + i * 64, f->sigs + (i % f->n) * 64, 64);
        memcpy(tm + i * 32, f->msgs + (i % f->n) * 32, 32);
        memcpy(tk + i * 32, f->pks + (i % f->n) * 32, 32);
    }
    CHECK(sigops_ed25519_ecverify(ts, tm, tk, big, v) == 0, "ed25519 n=100003: %s", sigops_last_error());
    for (i = 0; i < big; i++)
        if (v[i] != f->expect[i % f->n]) {
            CHECK(0, "ed25519 n=100003 row %lu", (unsigned long)i);
            break;
        }
    /* the variable-length / strict entry point on the same rows (32-byte messages through the offsets array, non-strict:
     * must agree with the fixed-size entry point) -- rust/src/ed25519_eddsa.rs */
    {
        uint64_t* off = malloc((f->n + 1) * sizeof(uint64_t));
        uint8_t* v2 = malloc(f->n);
        for (i = 0; i <= f->n; i++) off[i] = (uint64_t)i * 32;
        CHECK(sigops_ed25519_ecverify_msgs(f->sigs, f->msgs, off, f->pks, f->n, 0, v2) == 0, "ecverify_msgs: %s", sigops_last_error());
        for (i = 0; i < f->n; i++) CHECK(v2[i] == f->expect[i], "ecverify_msgs row %lu", (unsigned long)i);
        free(off);
        free(v2);
    }
    free(ts);
    free(tm);
    free(tk);
    free(v);
}

int main(int argc, char** argv) {
    fixture k1, r1, ed;
    int gpu;
    if (argc < 3) {
        fprintf(stderr, "usage: %s <fixture_dir> host|gpu\n", argv[0]);
        return 2;
    }
    gpu = strcmp(argv[2], "gpu") == 0;
    if (load_fixture(argv[1], "secp256k1", 64, &k1) || load_fixture(argv[1], "secp256r1", 64, &r1) ||
        load_fixture(argv[1], "ed25519", 32, &ed)) {
        fprintf(stderr, "cannot read the fixtures in %s\n", argv[1]);
        return 2;
    }
    check_bases(argv[1]);
    check_plan();
    if (!gpu) {
        /* no CUDA device: every compute entry point must fail with a nonzero code and say why (-> ShaderFailureError);
         * there is no CPU fallback */
        uint8_t out[64], st[1];
        CHECK(sigops_secp256k1_ecrecover(k1.sigs, k1.msgs, 1, out, st) != 0, "k1 must fail without a device");
        CHECK(strlen(sigops_last_error()) > 0, "last_error must explain the failure");
        CHECK(sigops_secp256r1_ecrecover(r1.sigs, r1.msgs, 1, out, st) != 0, "r1 must fail without a device");
        CHECK(sigops_ed25519_ecverify(ed.sigs, ed.msgs, ed.pks, 1, out) != 0, "ed25519 must fail without a device");
        CHECK(sigops_num_devices() == 0, "no devices");
    } else {
        double h2d = -1, ker = -1, d2h = -1;
        uint64_t l0 = sigops_kernel_launches();
        CHECK(sigops_init(NULL, 0) == 0, "init: %s", sigops_last_error());
        CHECK(sigops_num_devices() >= 1, "devices");
        run_ecdsa("secp256k1", sigops_secp256k1_ecrecover, &k1);
        run_ecdsa("secp256r1", sigops_secp256r1_ecrecover, &r1);
        run_ed(&ed);
        CHECK(sigops_last_timing(&h2d, &ker, &d2h) == 0 && ker > 0, "timing of the last call");
        CHECK(sigops_kernel_launches() > l0, "kernels were launched");
        {   /* a conflicting second init is refused, a matching one is a no-op */
            int wrong[2] = {0, 0};
            CHECK(sigops_init(wrong, 2) != 0, "duplicate ids must be refused");
            CHECK(sigops_init(NULL, 0) == 0, "re-init with the defaults is a no-op");
        }
        CHECK(sigops_shutdown() == 0, "shutdown");
        /* lazy re-initialisation after shutdown */
        {
            uint8_t out[64];
            CHECK(sigops_secp256k1_ecrecover(k1.sigs, k1.msgs, 1, out, NULL) == 0 && memcmp(out, k1.pks, 64) == 0, "after shutdown");
        }
    }
    if (failures) {
        fprintf(stderr, "%d check(s) failed\n", failures);
        return 1;
    }
    printf("abi_consumer %s: ok\n", argv[2]);
    return 0;
}
