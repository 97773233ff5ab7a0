/*
 * libsigops -- C ABI of the B200-native batch signature engine.
 *
 * This header is the drop-in boundary that replaces the reference's gpu.rs / shader.rs layer
 * (wgpu device + WGSL templating + multi-shader pipelines).  The reference has no FFI today: its
 * boundary is a set of Rust `pub async fn`s; a Rust maintainer binds these C entry points inside
 * the bodies of those functions (see INTEGRATION.md and rust/src/ffi.rs).  Each entry point cites
 * the reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - all pointers are HOST pointers unless the name ends in `_device`; buffers are caller-owned;
 *   - every function is blocking (except `_device` and `sigops_queue_submit*`, which are stream-ordered) and
 *     thread-safe: calls are serialised PER DEVICE (one host pipeline or one enqueue at a time on a device), so two
 *     callers that use disjoint devices (sigops_batch_on_devices, queues, `_device`) run concurrently;
 *   - return value 0 = success; nonzero = CUDA / runtime failure, which the Rust shim maps to
 *     `Err(ShaderFailureError)` (src/lib.rs:12-14).  There is NO CPU fallback: without a usable
 *     CUDA device every compute entry point fails with a nonzero code and sigops_last_error()
 *     says why;
 *   - byte layouts are exactly the reference's:
 *       k1/r1 signature  64 B  r_be[32] || s_be[32] with the y parity of R in bit 7 of byte 32
 *                               (src/wgsl/signature.wgsl:6-21, src/tests/mod.rs:151-163)
 *       message          32 B  big-endian prehash (fuel_crypto::Message)
 *       recovered key    64 B  X_be[32] || Y_be[32]  (src/wgsl/main/secp256k1_ecdsa_main_4.wgsl:44-51)
 *       ed25519 sig      64 B  R_compressed[32] || s_le[32]   (src/ed25519_eddsa.rs:37)
 *       ed25519 key      32 B  compressed A                   (src/ed25519_eddsa.rs:38)
 */
#ifndef SIGOPS_H
#define SIGOPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGOPS_CURVE_SECP256K1 0
#define SIGOPS_CURVE_SECP256R1 1
#define SIGOPS_CURVE_ED25519 2

/* status byte values written by the ecrecover entry points */
#define SIGOPS_STATUS_OK 0
#define SIGOPS_STATUS_INVALID 1 /* r==0, s==0, r>=n, s>=n, x=r not on curve, or Q = infinity */

/* Device pool.  Replaces `get_device_and_queue` (src/gpu.rs:5-35), which creates and destroys a wgpu
 * device on every API call; here the context (streams, device buffers, constant tables) persists.
 * device_ids == NULL or n_devices <= 0: use every visible CUDA device (or SIGOPS_GPUS=<count> of them).
 * Calling a compute entry point without sigops_init() initialises lazily with the defaults. */
/* A second sigops_init with an explicit device list that differs from the pool's fails (nonzero): call
 * sigops_shutdown() first.  sigops_init(NULL, 0) on an initialised pool is a no-op.
 * Init generates the fixed-base tables of the three curves on each device (csrc/ptab.h; they replace the 16-entry tables of
 * src/precompute.rs:14-69 that the reference uploads per call): one table per window of SIGOPS_GWIN bits (environment,
 * 4..24, default 22: 5.6 GB of HBM and 0.4 s per device; 20 -> 1.5 GB, 0.1 s, 0.6 % slower; 16 -> 0.1 GB, 3 % slower).  A
 * width out of range, or tables that do not fit, fail the init. */
int sigops_init(const int* device_ids, int n_devices);
int sigops_shutdown(void);
int sigops_num_devices(void);
/* Last error message of the calling process ("" if none). */
const char* sigops_last_error(void);

/* secp256k1_ecdsa::ecrecover / ecrecover_single_shader  (src/secp256k1_ecdsa.rs:61-66,215-219).
 * sigs: n*64 B, msgs: n*32 B, out_pubkeys: n*64 B, out_status: n bytes or NULL.
 * out_pubkeys[i] is 64 zero bytes when out_status[i] != 0.  n == 0 returns 0 and touches nothing
 * (src/secp256k1_ecdsa.rs:71-73).  The batch is split into contiguous shards over the pool's devices. */
int sigops_secp256k1_ecrecover(const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out_pubkeys,
                               uint8_t* out_status);

/* secp256r1_ecdsa::ecrecover / ecrecover_single_shader  (src/secp256r1_ecdsa.rs:62-67,216-220). */
int sigops_secp256r1_ecrecover(const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out_pubkeys,
                               uint8_t* out_status);

/* ed25519_eddsa::ecverify / ecverify_single  (src/ed25519_eddsa.rs:67-73,259-264).
 * sigs: n*64 B, msgs: n*32 B, pks: n*32 B, out_valid: n bytes, each 0 or 1
 * (the reference returns Vec<bool> from a u32-per-signature buffer, src/ed25519_eddsa.rs:251-254).
 * An undecodable public key yields 0 (dalek: VerifyingKey::from_bytes fails). */
int sigops_ed25519_ecverify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n,
                            uint8_t* out_valid);

/* The same three operations restricted to a subset of the pool: device_indices are positions in the pool (0 ..
 * sigops_num_devices()-1, not CUDA ordinals), n_devices == 0 means the whole pool.  curve = SIGOPS_CURVE_*; pks is used by
 * ed25519 only; out is n*64 bytes of keys (secp curves) or n verdict bytes (ed25519); out_status may be NULL.  Lets a
 * multi-threaded verifier give every thread its own GPUs -- the reference can only serialise its callers on one adapter
 * (`#[serial_test::serial]`, src/tests/secp256k1_ecdsa.rs:11-13; src/gpu.rs:7-14). */
int sigops_batch_on_devices(int curve, const int* device_indices, int n_devices, const uint8_t* sigs, const uint8_t* msgs,
                            const uint8_t* pks, size_t n, uint8_t* out, uint8_t* out_status);

/* Failure semantics of every host-buffer entry point (src/secp256k1_ecdsa.rs:203-205: all-or-nothing): when a call returns
 * nonzero every output buffer has been reset (keys / addresses / verdicts zero, status = SIGOPS_STATUS_INVALID), nothing the
 * call enqueued is still in flight, and the pool stays usable.  SIGOPS_FAIL_DEVICE=<pool index> (environment, read per
 * call) makes that device's shard fail with work in flight -- the fault-injection hook of the tests. */

/* ed25519 with variable-length messages and optional strict semantics -- the form fuel_crypto::ed25519::verify needs
 * (ed25519-dalek `verify_strict` over arbitrary-length messages); the reference hard-wires 32-byte messages and the
 * 96-byte hash input (src/wgsl/sha512.wgsl:114-123, src/wgsl/main/ed25519_eddsa_main_0.wgsl:40-85) and tests the
 * non-strict `verify` (src/tests/ed25519_eddsa.rs:26; strictness noted at src/curve_algos/ed25519_eddsa.rs:196-198).
 * msg_bytes: the messages back to back; msg_offsets: n + 1 byte offsets, message i = [msg_offsets[i], msg_offsets[i+1]).
 * flags: SIGOPS_ED25519_STRICT additionally requires that R decompresses and that neither A nor R has small order. */
#define SIGOPS_ED25519_STRICT 1u
int sigops_ed25519_ecverify_msgs(const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                                 const uint8_t* pks, size_t n, uint32_t flags, uint8_t* out_valid);

/* The steps on either side of recovery, on the device.  The reference's callers hash every
 * transaction to a `fuel_crypto::Message` (SHA-256) on the host before calling ecrecover
 * (src/tests/secp256k1_ecdsa.rs:21-22, src/benchmarks/secp256k1_ecdsa.rs:160-165) and derive the Fuel address
 * SHA-256(X || Y) from each returned key afterwards.
 * sigops_sha256_batch: out[i] = SHA-256(data[offsets[i] .. offsets[i+1])), 32 bytes each.
 * sigops_ecrecover_addresses: curve = SIGOPS_CURVE_SECP256K1 | SIGOPS_CURVE_SECP256R1.  msg_offsets != NULL: msg_bytes
 * holds the raw messages back to back and each is hashed with SHA-256 first; msg_offsets == NULL: msg_bytes holds n
 * 32-byte prehashes.  out_addresses: n * 32 bytes (zero where out_status[i] != 0); out_pubkeys (n * 64) and out_status
 * (n) may be NULL. */
int sigops_sha256_batch(const uint8_t* data, const uint64_t* offsets, size_t n, uint8_t* out);
int sigops_ecrecover_addresses(int curve, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                               size_t n, uint8_t* out_addresses, uint8_t* out_pubkeys, uint8_t* out_status);

/* precompute::{secp256k1_bases, secp256r1_bases, ed25519_bases}  (src/precompute.rs:12,36-69).
 * CPU-only compatibility table: 16 multiples (i+1)*G, coordinates in Montgomery form with
 * R = 2^(num_limbs*log_limb_size), little-endian log_limb_size-bit limbs; x||y per entry for the secp curves
 * (src/tests/mod.rs:134-149), x||y||t for ed25519 (src/tests/mod.rs:94-112).
 * log_limb_size must be in 11..=15 (src/wgsl/mont.wgsl:12,37).  *inout_len: capacity in u32 on entry,
 * number of u32 written on return (640 / 960 at log_limb_size = 13).  The engine itself never reads
 * these limbs (its own 32-bit tables are baked into the library).  Protocol: out == NULL (or *inout_len smaller than
 * needed with out == NULL) returns 0 and stores the required length; out != NULL with too small a capacity returns nonzero
 * and stores the required length. */
int sigops_precompute_bases(int curve, uint32_t log_limb_size, uint32_t* out, size_t* inout_len);

/* Shard planner used by the host entry points (pure host logic, no device needed): the batch is cut into
 * n_used <= n_devices contiguous ranges [bounds[g], bounds[g+1]) -- one per device, no exchange between them.  Replaces
 * the single-adapter dispatch of src/gpu.rs:5-35 and the pow-2 grid lookup `compute_num_workgroups`
 * (src/benchmarks/mod.rs:10-53).  bounds must hold n_devices + 1 entries. */
int sigops_plan_shards(size_t n, int n_devices, size_t* bounds, int* n_used);

/* Observability (the reference has none: `timestamp_writes: None`, src/gpu.rs:98).  Milliseconds of the last
 * host-buffer call, measured with CUDA events; max over the devices used.  A shard flows through an upload / kernel /
 * download pipeline in pieces, so: h2d = the first piece's upload (what is exposed before the first kernel), kernel = first
 * kernel start to last kernel end (the other uploads and downloads overlap it), d2h = what remains after the last kernel. */
int sigops_last_timing(double* h2d_ms, double* kernel_ms, double* d2h_ms);
/* Number of kernels the engine has launched in this process (all devices). */
uint64_t sigops_kernel_launches(void);

/* Pinned host memory for callers that want zero-staging transfers (replaces the MAP_READ staging buffer of
 * src/gpu.rs:138-166).  Plain malloc'ed (pageable) buffers are accepted everywhere too and are first-class: shards of
 * 16,384 signatures or more go through the library's own pinned staging, filled and drained by per-device copy threads
 * piece by piece under the kernels (SIGOPS_STAGING=0 disables it, SIGOPS_COPY_THREADS sets the threads per device). */
void* sigops_host_alloc(size_t bytes);
void sigops_host_free(void* p);

/* Streaming service mode: a persistent ring of pre-registered pinned buffers instead of one
 * blocking call per batch.  Replaces, for a long-running verifier, the per-call device creation, buffer allocation and
 * blocking `device.poll(Maintain::Wait)` of src/gpu.rs:5-35,129-170.
 * A queue serves ONE operation (SIGOPS_CURVE_*) with `depth` slots of up to `max_batch` signatures each, on ONE device of the
 * pool (device_index = index into the pool, not a CUDA ordinal) or, with device_index = -1, with its slots spread
 * round-robin over every device of the pool (slot i lives on pool device i mod G: an 8-GPU verifier needs one queue).  Every slot owns pinned host input / output arrays,
 * device buffers, scratch and a CUDA stream; a slot's upload, fused kernel and download are replayed as one CUDA graph
 * while the request size repeats.  Slots in flight run concurrently on the device: small requests (a few thousand
 * signatures use one or two of the 16 resident warps per SM) overlap, so throughput scales with the number of slots in
 * flight at the latency of a single request.
 * Life cycle of a slot: write the inputs into the arrays sigops_queue_buffers() returns (layouts as in the blocking entry
 * points; `pks` is NULL unless the queue is ed25519; `status` is NULL for ed25519) -> sigops_queue_submit(slot, n)
 * (asynchronous; fails if the slot is still in flight or n > max_batch) -> sigops_queue_wait(slot) (blocks until that
 * request has completed) -> read `out` / `status` -> reuse the slot.  sigops_queue_poll() is the non-blocking test.
 * Thread-safe: several threads may drive different slots of one queue; one slot must not be waited on by two threads at
 * once (the second wait fails).  A request that fails on the device is reported by sigops_queue_wait (nonzero) and the slot
 * is released -- the queue is never wedged.  All queues must be destroyed before
 * sigops_shutdown().  SIGOPS_QUEUE_GRAPHS=0 disables graph replay (plain stream launches). */
typedef struct sigops_queue sigops_queue;
#define SIGOPS_QUEUE_MAX_DEPTH 64
int sigops_queue_create(int curve, int device_index, size_t max_batch, int depth, sigops_queue** out);
int sigops_queue_destroy(sigops_queue* q);
int sigops_queue_buffers(sigops_queue* q, int slot, uint8_t** sigs, uint8_t** msgs, uint8_t** pks, uint8_t** out,
                         uint8_t** status);
int sigops_queue_submit(sigops_queue* q, int slot, size_t n);
/* Device-resident producer: the request's inputs already live in the memory of CUDA device
 * `src_device` (any device of the box, e.g. the GPU that produced or received the transactions).  They are moved into the
 * slot's device buffers with cudaMemcpyPeerAsync -- over NVLink when the devices are peers -- on the slot's stream, followed
 * by the kernel and the download of the results into the slot's pinned output arrays; the host never touches the inputs.
 * ready_event (a cudaEvent_t, may be NULL): the slot's stream waits for it before copying (the producer's completion).
 * Replaces the host-only upload path of src/gpu.rs:41-49. */
int sigops_queue_submit_device(sigops_queue* q, int slot, const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n,
                               int src_device, void* ready_event);
/* CUDA ordinal of the device a slot lives on (-1 on bad arguments). */
int sigops_queue_slot_device(sigops_queue* q, int slot);
int sigops_queue_poll(sigops_queue* q, int slot, int* done);
/* n_done (may be NULL): size of the request that completed; device_ms (may be NULL): upload + kernel + download time of
 * that request on the device (CUDA events). */
int sigops_queue_wait(sigops_queue* q, int slot, size_t* n_done, double* device_ms);
int sigops_queue_info(sigops_queue* q, int* curve, int* device_index, size_t* max_batch, int* depth,
                      uint64_t* graph_launches, uint64_t* graph_captures);

/* Device-resident variants: inputs and outputs already live in the CURRENT CUDA device's memory
 * (16-byte aligned); the kernel is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream) and the
 * call returns without synchronising.  Launches that share a device's work tables are ordered behind one another with
 * events, whatever streams they were enqueued on (two `_device` calls on different streams, or a `_device` call and a
 * host-buffer call, never run on the same tables at once).  Used by bench.py for the kernel-only figure. */
int sigops_secp256k1_ecrecover_device(const void* d_sigs, const void* d_msgs, size_t n, void* d_out_pubkeys,
                                      void* d_out_status, void* cuda_stream);
int sigops_secp256r1_ecrecover_device(const void* d_sigs, const void* d_msgs, size_t n, void* d_out_pubkeys,
                                      void* d_out_status, void* cuda_stream);
int sigops_ed25519_ecverify_device(const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n,
                                   void* d_out_valid, void* cuda_stream);

/* ---- test and measurement shims (mirror the reference's single-invocation test shaders, src/wgsl/tests/) ---- */

/* Runs device function `op` on n_items independent inputs, one per GPU thread.  Items are fixed-size groups of
 * 32-bit words (in_words / out_words per item, see SIGOPS_UNIT_* below).  Host pointers. */
int sigops_test_unit(int op, const uint32_t* in, size_t n_items, uint32_t* out);
int sigops_test_unit_shape(int op, int* in_words, int* out_words);

/* Integer-pipe micro-benchmark: independent multiply-add chains at full occupancy on every SM of the current
 * device.  kind 0 = IMAD (32-bit mad.lo), 1 = IMAD.WIDE.U32 (mad.wide), 2 = IMAD.WIDE.U32.X carry chains
 * (mad.lo.cc/madc.hi.cc), 3 = IADD3 (add), 4 = mixed 1:1 IMAD.WIDE + IADD3, 5 = DFMA (fma.rn.f64), 6 = IMAD.HI
 * (mad.hi.u32).  Returns the measured rate in
 * instructions (thread-level operations) per second in *ops_per_sec, elapsed GPU time in *ms. */
int sigops_imad_peak(int kind, int iters, double* ops_per_sec, double* ms);

enum {
    SIGOPS_UNIT_K1_MUL = 0,    /* in 16 (a,b plain)            out 8  : a*b mod p, canonical          */
    SIGOPS_UNIT_K1_SQR = 1,    /* in 8                         out 8                                   */
    SIGOPS_UNIT_K1_ADD = 2,    /* in 16                        out 8                                   */
    SIGOPS_UNIT_K1_SUB = 3,    /* in 16                        out 8                                   */
    SIGOPS_UNIT_K1_INV = 4,    /* in 8                         out 8  : a^-1 mod p (safegcd), 0 -> 0    */
    SIGOPS_UNIT_K1_SQRT = 5,   /* in 8                         out 8  : a^((p+1)/4)                    */
    SIGOPS_UNIT_R1_MUL = 6,    /* plain in, plain out (through Montgomery form)                        */
    SIGOPS_UNIT_R1_SQR = 7,
    SIGOPS_UNIT_R1_ADD = 8,
    SIGOPS_UNIT_R1_SUB = 9,
    SIGOPS_UNIT_R1_INV = 10,
    SIGOPS_UNIT_R1_SQRT = 11,
    SIGOPS_UNIT_ED_MUL = 12,
    SIGOPS_UNIT_ED_SQR = 13,
    SIGOPS_UNIT_ED_ADD = 14,
    SIGOPS_UNIT_ED_SUB = 15,
    SIGOPS_UNIT_ED_INV = 16,
    SIGOPS_UNIT_ED_POW_P58 = 17,
    SIGOPS_UNIT_K1N_MUL = 18,  /* in 16 (a,b < n plain)        out 8  : a*b mod n                      */
    SIGOPS_UNIT_K1N_INV = 19,  /* in 8                         out 8  : a^-1 mod n                     */
    SIGOPS_UNIT_R1N_MUL = 20,
    SIGOPS_UNIT_R1N_INV = 21,
    SIGOPS_UNIT_EDL_REDUCE512 = 22, /* in 16 (LE 512-bit)      out 8  : mod L                          */
    SIGOPS_UNIT_SHA512_96 = 23,     /* in 24 (96 bytes)        out 16 : digest bytes                   */
    SIGOPS_UNIT_K1_GLV = 24,        /* in 8 (k)                out 12 : |k1|[5], |k2|[5], neg1, neg2   */
    SIGOPS_UNIT_MUL8X8 = 25,        /* in 16                   out 16 : full 512-bit product           */
    SIGOPS_UNIT_SQR8 = 26,          /* in 8                    out 16                                  */
    SIGOPS_UNIT_K1_MULPT = 27,      /* in 24 (k, x, y plain)   out 17 : k*(x,y) affine x,y + inf flag  */
    SIGOPS_UNIT_R1_MULPT = 28,
    SIGOPS_UNIT_ED_MULPT = 29,      /* in 24 (k, x, y)         out 16 : k*(x,y) affine x,y             */
    SIGOPS_UNIT_K1_DOUBLE_MUL = 30, /* in 32 (u1,u2,x,y)       out 17 : u1*G + u2*(x,y)                */
    SIGOPS_UNIT_R1_DOUBLE_MUL = 31,
    SIGOPS_UNIT_K1N_INV_FERMAT = 32, /* in 8                    out 8  : a^(n-2) mod n (cross-check of the safegcd path) */
    SIGOPS_UNIT_K1_INV_FERMAT = 33,
    SIGOPS_UNIT_R1_INV_FERMAT = 34,
    SIGOPS_UNIT_ED_INV_FERMAT = 35,
    SIGOPS_UNIT_SHA256_64 = 36,      /* in 16 (64 bytes)        out 8  : SHA-256 digest bytes           */
    /* raw-representation probes of the base fields (field id 0 = secp256k1, 1 = P-256, 2 = 2^255-19 in word 0): the operands
     * are the INTERNAL, weakly reduced limbs (any value below 2^256), so that the once-in-2^31 fix-up paths can be hit */
    SIGOPS_UNIT_RAW_ADDSUB = 37,     /* in 17 (id, a, b)        out 16 : a+b, a-b as internal limbs     */
    SIGOPS_UNIT_RAW_REDUCE16 = 38,   /* in 17 (id, t[16])       out 8  : reduce16(t), ids 0 and 2 only  */
    SIGOPS_UNIT_RAW_SHL = 39,        /* in 10 (id, K = 2|3, a)  out 8  : a * 2^K as internal limbs      */
    /* the lane-group kernels' twins (several cooperating threads per item, complete projective / four-way Edwards formulas) */
    SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL = 40, /* in 32 (u1,u2,x,y)  out 17 : u1*G + u2*(x,y)                */
    SIGOPS_UNIT_R1_GROUP_DOUBLE_MUL = 41,
    SIGOPS_UNIT_ED_GROUP_MULPT = 42,      /* in 24 (k, x, y)    out 16 : k*(x,y) affine x,y             */
    SIGOPS_UNIT_ED_FIXED_MUL = 43,        /* in 8 (s)           out 16 : s*B affine x,y, through the positional table only */
    SIGOPS_UNIT_COUNT = 44
};

#ifdef __cplusplus
}
#endif
#endif /* SIGOPS_H */
