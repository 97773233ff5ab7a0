// libsigops host layer: persistent device pool, contiguous sharding across GPUs, pinned-aware transfers and the
// extern "C" entry points declared in include/sigops.h.
//
// Replaces src/gpu.rs:5-170 (per-call wgpu device creation, storage/uniform buffers, pipeline creation from a WGSL
// string, dispatch, staging-buffer read-back and device.destroy()) and the host halves of
// src/secp256k1_ecdsa.rs:11-213, src/secp256r1_ecdsa.rs:12-214, src/ed25519_eddsa.rs:12-257 (flatten, zero-pad to a
// power of two, five to seven dispatches, slice the padded result).  Here: no padding, one kernel launch per shard,
// device context and buffers persist across calls, and there is no collective -- shards are independent
// (SURVEY.md 8e).  There is no CPU fallback: without a CUDA device every compute entry point returns nonzero.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"

using namespace sigops;

namespace {

std::mutex g_mu;      // serialises the entry points
std::mutex g_err_mu;  // shard threads may report errors concurrently
std::string g_err;
std::atomic<uint64_t> g_launches{0};
thread_local bool t_capturing = false;  // stream capture records a launch, it does not perform one

void set_err(const std::string& s) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = s;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char b_[512];                                                                              \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            set_err(b_);                                                                               \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

struct Device {
    int id = -1;
    int sms = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;  // kernels / uploads / downloads
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[8] = {}, ev_k[8] = {};
    // device buffers (grow-only)
    uint8_t* d_in = nullptr;
    size_t in_cap = 0;
    uint8_t* d_out = nullptr;
    size_t out_cap = 0;
    Q4* scratch = nullptr;
    size_t scratch_cap = 0;  // in Q4
    const u32 *k1g = nullptr, *r1g = nullptr, *edb = nullptr;
    int grid_k1 = 0, grid_r1 = 0, grid_ed = 0, grid_edm = 0, grid_unit = 0;
    float ms_h2d = 0, ms_kernel = 0, ms_d2h = 0;
    bool smem_tables = true;  // SIGOPS_SMEM_TABLES=0 leaves the fixed-base tables in L2
};

std::vector<Device> g_dev;
bool g_inited = false;

int ensure_buf(uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 4096;
    CK(cudaMalloc((void**)p, want));
    *cap = want;
    return 0;
}

int ensure_scratch(Device& d, size_t q4s) {
    if (q4s <= d.scratch_cap) return 0;
    if (d.scratch) CK(cudaFree(d.scratch));
    d.scratch = nullptr;
    d.scratch_cap = 0;
    CK(cudaMalloc((void**)&d.scratch, q4s * sizeof(Q4)));
    d.scratch_cap = q4s;
    return 0;
}

template <class K>
int max_grid(Device& d, K kernel, int* out) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0));
    if (per_sm < 1) per_sm = 1;
    *out = per_sm * d.sms;
    return 0;
}

int init_device(Device& d, int id) {
    d.id = id;
    CK(cudaSetDevice(id));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, id));
    if (prop.major < 10) {
        char b[256];
        snprintf(b, sizeof b, "device %d (%s) is sm_%d%d; libsigops is built for sm_100a only", id, prop.name, prop.major,
                 prop.minor);
        set_err(b);
        return 1;
    }
    d.sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking));
    for (auto& e : d.ev) CK(cudaEventCreate(&e));
    for (auto& e : d.ev_in) CK(cudaEventCreateWithFlags(&e, cudaEventDefault));
    for (auto& e : d.ev_k) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // fixed-base tables: computed on the device once, resident for the life of the context (L2-sized: 576 KB)
    u32 *k1t = nullptr, *r1t = nullptr, *edt = nullptr;
    CK(cudaMalloc((void**)&k1t, (size_t)2 * kGTabEntries * 16 * sizeof(u32)));
    CK(cudaMalloc((void**)&r1t, (size_t)kGTabEntries * 16 * sizeof(u32)));
    CK(cudaMalloc((void**)&edt, (size_t)kGTabEntries * 24 * sizeof(u32)));
    gen_tables_kernel<<<(kGTabEntries + 63) / 64, 64, 0, d.stream>>>(k1t, r1t, edt);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d.stream));
    g_launches++;
    d.k1g = k1t;
    d.r1g = r1t;
    d.edb = edt;
    if (const char* e = getenv("SIGOPS_SMEM_TABLES")) d.smem_tables = atoi(e) != 0;
    CK(cudaFuncSetAttribute(ecrecover_kernel<CurveK1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGTabEntries * 16 * 4));
    CK(cudaFuncSetAttribute(ecrecover_kernel<CurveR1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGTabEntries * 16 * 4));
    CK(cudaFuncSetAttribute(ed25519_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGTabEntries * 24 * 4));
    if (max_grid(d, ecrecover_kernel<CurveK1>, &d.grid_k1)) return 1;
    if (max_grid(d, ecrecover_kernel<CurveR1>, &d.grid_r1)) return 1;
    if (max_grid(d, ed25519_verify_kernel, &d.grid_ed)) return 1;
    if (max_grid(d, ed25519_verify_msgs_kernel, &d.grid_edm)) return 1;
    if (max_grid(d, unit_kernel, &d.grid_unit)) return 1;
    return 0;
}

int do_init(const int* ids, int n) {
    if (g_inited) return 0;
    // The streaming mode keeps many small launches in flight on separate streams; with the default of 8 hardware work
    // queues no more than ~4 requests overlap and further submits block in the driver (measured: profiles/r01_queue_sweep.json).
    // Only effective if this process has not created its CUDA context yet; an explicit setting wins.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_err(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                "); libsigops has no CPU fallback");
        return 2;
    }
    std::vector<int> use;
    if (ids && n > 0) {
        for (int i = 0; i < n; i++) {
            if (ids[i] < 0 || ids[i] >= count) {
                set_err("sigops_init: device id out of range");
                return 1;
            }
            use.push_back(ids[i]);
        }
    } else {
        int want = count;
        if (const char* s = getenv("SIGOPS_GPUS")) {
            int v = atoi(s);
            if (v >= 1 && v < want) want = v;
        }
        for (int i = 0; i < want; i++) use.push_back(i);
    }
    g_dev.assign(use.size(), Device());
    for (size_t i = 0; i < use.size(); i++)
        if (init_device(g_dev[i], use[i])) {
            g_dev.clear();
            return 1;
        }
    g_inited = true;
    return 0;
}

// device context for the CURRENT device (used by the *_device entry points and the test shims)
Device* current_device() {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return nullptr;
    for (auto& d : g_dev)
        if (d.id == cur) return &d;
    return nullptr;
}

enum Op { OP_K1 = 0, OP_R1 = 1, OP_ED = 2 };

// Launch geometry: full 512-thread blocks (one per SM, all 16 warps phase-locked by the kernels' barriers) once the batch
// covers the device; smaller batches are spread over all SMs with proportionally smaller blocks instead of filling
// a few SMs to the brim (latency of the 64 ... 64k end of the batch-size sweep).
void launch_geometry(const Device& d, Op op, size_t n, int* grid, int* tpb_out) {
    const int max_g = op == OP_K1 ? d.grid_k1 : op == OP_R1 ? d.grid_r1 : d.grid_ed;
    int tpb = kBlock;
    if (n < (size_t)d.sms * kBlock) {
        size_t per_sm = (n + d.sms - 1) / d.sms;
        tpb = (int)std::min<size_t>(kBlock, std::max<size_t>(32, (per_sm + 31) / 32 * 32));
    }
    size_t blocks = (n + tpb - 1) / tpb;
    *grid = (int)std::min<size_t>(blocks, (size_t)max_g);
    *tpb_out = tpb;
}

// The kernels index their per-thread scratch by global thread id: `scratch` must hold chunks x grid x tpb Q4 and must not
// be shared by two launches that may run concurrently.
// A launch of w.f waves (one wave = SMs x 512 signatures, one per thread) costs ceil(w.f) full passes of ~1.65 ms: the last,
// partial pass runs every thread (rows past the end are clamped) at 16 warps per SM.  When the tail is at most
// kTailSplitNum/kTailSplitDen of a wave it is launched on its own right behind the main part, with the small-batch
// geometry: fewer warps per SM finish a pass sooner (0.85 ms at one warp per scheduler).  Matters for shards of one to a
// few waves -- the 1M block cut over 4 or 8 GPUs, the middle of the batch-size sweep.
constexpr size_t kTailSplitNum = 3, kTailSplitDen = 4;
size_t tail_split(const Device& d, size_t n) {  // signatures in the main launch (n if no split)
    const size_t wave = (size_t)d.sms * kBlock;
    if (n < wave) return n;
    const size_t tail = n % wave;
    if (tail == 0 || tail * kTailSplitDen > wave * kTailSplitNum) return n;
    if (const char* e = getenv("SIGOPS_TAIL_SPLIT"))
        if (atoi(e) == 0) return n;
    return n - tail;
}

int launch_op_scratch(Device& d, Op op, const uint8_t* d_sigs, const uint8_t* d_msgs, const uint8_t* d_pks, size_t n,
                      uint8_t* d_out, uint8_t* d_status, Q4* scratch, cudaStream_t st) {
    if (n == 0) return 0;
    const size_t main_n = tail_split(d, n);
    if (main_n < n) {
        const size_t os = op == OP_ED ? 1 : 64;
        if (launch_op_scratch(d, op, d_sigs, d_msgs, d_pks, main_n, d_out, d_status, scratch, st)) return 1;
        return launch_op_scratch(d, op, d_sigs + main_n * 64, d_msgs + main_n * 32, d_pks ? d_pks + main_n * 32 : nullptr,
                                 n - main_n, d_out + main_n * os, d_status ? d_status + main_n : nullptr, scratch, st);
    }
    int grid, tpb;
    launch_geometry(d, op, n, &grid, &tpb);
    // stage the fixed-base table in shared memory when the launch is big enough to amortise the copy (>= 1 full pass)
    const bool stage = d.smem_tables && n >= (size_t)d.sms * kBlock && tpb == kBlock;
    const u32 sw_words = stage ? (u32)kGTabEntries * 16 : 0, ed_words = stage ? (u32)kGTabEntries * 24 : 0;
    switch (op) {
        case OP_K1:
            ecrecover_kernel<CurveK1><<<grid, tpb, sw_words * 4, st>>>((const Q4*)d_sigs, (const Q4*)d_msgs, n, (Q4*)d_out,
                                                                       d_status, scratch, d.k1g, sw_words);
            break;
        case OP_R1:
            ecrecover_kernel<CurveR1><<<grid, tpb, sw_words * 4, st>>>((const Q4*)d_sigs, (const Q4*)d_msgs, n, (Q4*)d_out,
                                                                       d_status, scratch, d.r1g, sw_words);
            break;
        case OP_ED:
            ed25519_verify_kernel<<<grid, tpb, ed_words * 4, st>>>((const Q4*)d_sigs, (const Q4*)d_msgs, (const Q4*)d_pks, n,
                                                                   d_out, scratch, d.edb, ed_words);
            break;
    }
    CK(cudaGetLastError());
    if (!t_capturing) g_launches++;
    return 0;
}

int launch_op(Device& d, Op op, const uint8_t* d_sigs, const uint8_t* d_msgs, const uint8_t* d_pks, size_t n,
              uint8_t* d_out, uint8_t* d_status, cudaStream_t st) {
    if (n == 0) return 0;
    const int max_g = op == OP_K1 ? d.grid_k1 : op == OP_R1 ? d.grid_r1 : d.grid_ed;
    const size_t chunks = op == OP_ED ? kEdBatchChunks : kSwBatchChunks;
    if (ensure_scratch(d, chunks * (size_t)max_g * kBlock)) return 1;
    return launch_op_scratch(d, op, d_sigs, d_msgs, d_pks, n, d_out, d_status, d.scratch, st);
}

// One shard on one device.  The shard is cut into up to kMaxChunks pieces that flow through a three-stage pipeline
// (H2D on s_in, kernel on stream, D2H on s_out, chained with events), so that with pinned host buffers only the first
// piece's upload and the last piece's download are exposed.  Kernels run back to back on ONE stream: they share the
// per-thread scratch tables.
constexpr int kMaxChunks = 8;
constexpr size_t kLeadWaves = 2;  // passes in the first, short piece (~4 ms of kernel): exposed upload ~0.3 ms

// Piece plan: pieces are whole multiples of one wave (SMs x 512 threads, one signature per thread per pass) so that only
// the last piece ends on a partial wave.  The first piece is short (kLeadWaves) so that the kernels start early; the rest
// is cut into equal pieces of at most kSwBatch waves -- one full shared-inversion batch per thread (curve_sw.cuh).
int plan_pieces(const Device& d, size_t n, size_t* bounds) {
    int max_chunks = kMaxChunks;
    if (const char* e = getenv("SIGOPS_MAX_CHUNKS")) max_chunks = std::max(1, std::min(kMaxChunks, atoi(e)));
    const size_t wave = (size_t)d.sms * kBlock;
    const size_t n_waves = (n + wave - 1) / wave;
    int chunks = 1;
    bounds[0] = 0;
    if (max_chunks > 1 && n_waves >= 2 * kLeadWaves + 1) {
        const size_t rest = n_waves - kLeadWaves;
        const int rest_chunks = (int)std::min<size_t>((size_t)max_chunks - 1, (rest + kSwBatch - 1) / kSwBatch);
        chunks = 1 + rest_chunks;
        bounds[1] = kLeadWaves * wave;
        for (int c = 1; c <= rest_chunks; c++) bounds[1 + c] = std::min(n, (kLeadWaves + rest * (size_t)c / rest_chunks) * wave);
    } else if (max_chunks > 1 && n > wave) {
        // one to four waves (the 1M block cut over 4 or 8 GPUs): a one-wave lead piece, the rest behind it -- the second
        // upload and the first download overlap a kernel instead of being exposed
        chunks = 2;
        bounds[1] = wave;
    }
    bounds[chunks] = n;
    return chunks;
}

// The three-stage pipeline over the pieces of one shard.  up(lo, m, stream) enqueues the uploads of piece [lo, lo + m),
// launch(lo, m, stream) its kernels, down(lo, m, stream) its downloads; each returns nonzero on failure.
template <class Up, class Launch, class Down>
int run_pipeline(Device& d, size_t n, Up up, Launch launch, Down down) {
    size_t bounds[kMaxChunks + 1];
    const int chunks = plan_pieces(d, n, bounds);
    CK(cudaEventRecord(d.ev[0], d.s_in));
    for (int c = 0; c < chunks; c++) {
        const size_t lo = bounds[c], m = bounds[c + 1] - lo;
        if (up(lo, m, d.s_in)) return 1;
        CK(cudaEventRecord(d.ev_in[c], d.s_in));
        CK(cudaStreamWaitEvent(d.stream, d.ev_in[c], 0));
        if (c == 0) CK(cudaEventRecord(d.ev[1], d.stream));
        if (launch(lo, m, d.stream)) return 1;
        CK(cudaEventRecord(d.ev_k[c], d.stream));
    }
    CK(cudaEventRecord(d.ev[2], d.stream));
    // downloads are enqueued after every upload and kernel: a download into pageable memory blocks the calling thread
    // until its kernel has finished, which would otherwise stall the staging of the next piece's upload
    for (int c = 0; c < chunks; c++) {
        const size_t lo = bounds[c], m = bounds[c + 1] - lo;
        CK(cudaStreamWaitEvent(d.s_out, d.ev_k[c], 0));
        if (down(lo, m, d.s_out)) return 1;
    }
    CK(cudaEventRecord(d.ev[3], d.s_out));
    CK(cudaStreamSynchronize(d.s_out));
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaStreamSynchronize(d.s_in));
    // h2d: first upload until the first kernel may start; kernel: first kernel start to last kernel end (uploads and
    // downloads of the other pieces overlap it); d2h: what remains after the last kernel
    CK(cudaEventElapsedTime(&d.ms_h2d, d.ev[0], d.ev_in[0]));
    CK(cudaEventElapsedTime(&d.ms_kernel, d.ev[1], d.ev[2]));
    CK(cudaEventElapsedTime(&d.ms_d2h, d.ev[2], d.ev[3]));
    if (d.ms_d2h < 0) d.ms_d2h = 0;
    return 0;
}

int run_shard(Device& d, Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* out,
              uint8_t* status) {
    CK(cudaSetDevice(d.id));
    const size_t in_bytes = n * (op == OP_ED ? 128 : 96);
    const size_t out_stride = op == OP_ED ? 1 : 64;
    const size_t out_bytes = op == OP_ED ? n : n * 65;
    if (ensure_buf(&d.d_in, &d.in_cap, in_bytes + 64)) return 1;
    if (ensure_buf(&d.d_out, &d.out_cap, out_bytes + 64)) return 1;
    uint8_t* d_sigs = d.d_in;
    uint8_t* d_msgs = d.d_in + n * 64;
    uint8_t* d_pks = d.d_in + n * 96;
    uint8_t* d_status = op == OP_ED ? nullptr : d.d_out + n * 64;
    auto up = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_msgs + lo * 32, msgs + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        if (op == OP_ED) CK(cudaMemcpyAsync(d_pks + lo * 32, pks + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto launch = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        return launch_op(d, op, d_sigs + lo * 64, d_msgs + lo * 32, d_pks + lo * 32, m, d.d_out + lo * out_stride,
                         d_status ? d_status + lo : nullptr, st);
    };
    auto down = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(out + lo * out_stride, d.d_out + lo * out_stride, m * out_stride, cudaMemcpyDeviceToHost, st));
        if (op != OP_ED && status) CK(cudaMemcpyAsync(status + lo, d_status + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    return run_pipeline(d, n, up, launch, down);
}

// A device's shard is processed in sub-shards of at most kMaxSubShard signatures so that the device buffers stay
// bounded (2 GiB in, 1 GiB out) however large the batch (the reference allows up to 2^30 signatures per call,
// src/secp256k1_ecdsa.rs:22).  Timings accumulate over the sub-shards.
constexpr size_t kMaxSubShard = (size_t)1 << 24;

int run_shard_bounded(Device& d, Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n,
                      uint8_t* out, uint8_t* status) {
    const size_t out_stride = op == OP_ED ? 1 : 64;
    float h2d = 0, ker = 0, d2h = 0;
    size_t sub = kMaxSubShard;
    if (const char* e = getenv("SIGOPS_MAX_SUBSHARD")) sub = std::max<size_t>(1, (size_t)atoll(e));  // test hook
    for (size_t lo = 0; lo < n; lo += sub) {
        const size_t m = std::min(sub, n - lo);
        if (int rc = run_shard(d, op, sigs + lo * 64, msgs + lo * 32, pks ? pks + lo * 32 : nullptr, m,
                               out + lo * out_stride, status ? status + lo : nullptr))
            return rc;
        h2d += d.ms_h2d;
        ker += d.ms_kernel;
        d2h += d.ms_d2h;
    }
    d.ms_h2d = h2d;
    d.ms_kernel = ker;
    d.ms_d2h = d2h;
    return 0;
}

// below this many signatures per device a shard is not worth a GPU of its own (launch + sync latency dominates)
constexpr size_t kMinShard = 4096;

// Shard planner: G_eff = min(G, ceil(n / kMinShard)) devices, device g gets the contiguous range
// [floor(g*n/G_eff), floor((g+1)*n/G_eff)).  bounds receives G_eff + 1 entries.  Pure host logic.
int plan_shards(size_t n, int n_devices, size_t* bounds) {
    size_t G = std::min<size_t>((size_t)std::max(n_devices, 1), (n + kMinShard - 1) / kMinShard);
    if (G < 1) G = 1;
    for (size_t g = 0; g <= G; g++) bounds[g] = g * n / G;  // n <= 2^30 (src/secp256k1_ecdsa.rs:22): no overflow
    return (int)G;
}

int run_batch(Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* out,
              uint8_t* status) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n == 0) return 0;
    if (!sigs || !msgs || !out || (op == OP_ED && !pks)) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = do_init(nullptr, 0)) return rc;
    for (auto& d : g_dev) d.ms_h2d = d.ms_kernel = d.ms_d2h = 0;
    std::vector<size_t> bounds(g_dev.size() + 1);
    const size_t G = (size_t)plan_shards(n, (int)g_dev.size(), bounds.data());
    const size_t out_stride = op == OP_ED ? 1 : 64;
    if (G == 1) return run_shard_bounded(g_dev[0], op, sigs, msgs, pks, n, out, status);
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        const size_t lo = bounds[g], hi = bounds[g + 1];  // contiguous shard [lo, hi)
        th.emplace_back([&, g, lo, hi]() {
            rcs[g] = run_shard_bounded(g_dev[g], op, sigs + lo * 64, msgs + lo * 32, pks ? pks + lo * 32 : nullptr, hi - lo,
                               out + lo * out_stride, status ? status + lo : nullptr);
        });
    }
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return rcs[g];  // one failed shard fails the whole call (all-or-nothing, like ShaderFailureError)
    return 0;
}

int run_device(Op op, const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n, void* d_out, void* d_status,
               void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = do_init(nullptr, 0)) return rc;
    Device* d = current_device();
    if (!d) {
        set_err("current CUDA device is not part of the sigops pool");
        return 1;
    }
    return launch_op(*d, op, (const uint8_t*)d_sigs, (const uint8_t*)d_msgs, (const uint8_t*)d_pks, n, (uint8_t*)d_out,
                     (uint8_t*)d_status, (cudaStream_t)stream);
}

// ---- precompute::*_bases (CPU only, as in the reference) ------------------------------------------------
// 256-bit modular helpers on the portable paths of field.cuh; entry i = (i+1)*G in affine coordinates.
template <class F>
void to_limbs_mont(std::vector<uint32_t>& out, const Fe& coord_plain_in_F, const u32* p_limbs, int num_limbs,
                   int log_limb_size) {
    // value v (canonical, plain) -> v * 2^(num_limbs*log_limb_size) mod p by repeated doubling mod p
    u32 v[8];
    F::to_plain(v, coord_plain_in_F);
    const int shift = num_limbs * log_limb_size;
    for (int i = 0; i < shift; i++) {
        u32 t[8], u[8];
        u32 c = add8(t, v, v);
        u32 bw = sub8(u, t, p_limbs);
        select8(v, c != 0 || bw == 0, t, u);
    }
    // split into log_limb_size-bit limbs, little-endian
    for (int l = 0; l < num_limbs; l++) {
        uint32_t limb = 0;
        for (int b = 0; b < log_limb_size; b++) {
            int bit = l * log_limb_size + b;
            if (bit < 256) limb |= ((v[bit >> 5] >> (bit & 31)) & 1u) << b;
        }
        out.push_back(limb);
    }
}

int calc_num_limbs(int log_limb_size) {  // multiprecision::utils::calc_num_limbs(log_limb_size, 256)
    int l = 256 / log_limb_size;
    while (l * log_limb_size <= 256) l++;
    return l;
}

template <class C>
void sw_bases(std::vector<uint32_t>& out, const u32* gxy_table_entry0, const u32* p_limbs, int log_limb_size) {
    typedef typename C::F F;
    const int num_limbs = calc_num_limbs(log_limb_size);
    Fe gx, gy;
    F::from_table(gx, gxy_table_entry0);
    F::from_table(gy, gxy_table_entry0 + 8);
    JacPoint P;
    P.inf = true;
    for (int i = 0; i < 16; i++) {
        jac_madd<C>(P, gx, gy);
        Fe zi, zi2, ax, ay;
        fe_inv((F*)0, zi, P.Z);
        F::sqr(zi2, zi);
        F::mul(ax, P.X, zi2);
        F::mul(zi2, zi2, zi);
        F::mul(ay, P.Y, zi2);
        to_limbs_mont<F>(out, ax, p_limbs, num_limbs, log_limb_size);
        to_limbs_mont<F>(out, ay, p_limbs, num_limbs, log_limb_size);
    }
}


// ---- streaming service mode -----------------------------------------------------------------------
// A queue is a ring of `depth` slots on ONE device for ONE operation.  Every slot owns pinned host staging (inputs and
// outputs), device buffers, its own per-thread scratch and its own stream, so the slots of a queue run concurrently on the
// device: a request of a few thousand signatures occupies one or two warps per SM (launch_geometry), and up to 16 warps per
// SM are resident, so `depth` small requests in flight multiply the throughput at the latency of one.  A slot's H2D copies,
// kernel and D2H copies are replayed as one CUDA graph while the request size repeats (re-captured when it changes).
// Replaces the per-call device creation / buffer allocation / blocking poll of src/gpu.rs:5-35,129-170.
struct QSlot {
    cudaStream_t st = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr;  // pinned: sigs (cap*64) | msgs (cap*32) | pks (cap*32); out (cap*64) | status (cap)
    uint8_t *d_in = nullptr, *d_out = nullptr;
    Q4* scratch = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t exec_n = 0;
    bool busy = false;
    size_t n = 0;
    float last_ms = 0;
};

}  // namespace

struct sigops_queue {
    std::mutex mu;
    Device* dev = nullptr;
    Op op = OP_K1;
    size_t cap = 0;
    bool graphs = true;
    std::vector<QSlot> slots;
    uint64_t graph_launches = 0, graph_captures = 0;
};

namespace {

std::atomic<int> g_live_queues{0};

inline size_t q_out_stride(Op op) { return op == OP_ED ? 1 : 64; }

// enqueue one request of n signatures of slot s on stream st (also used under stream capture: no allocation, no sync)
int queue_enqueue(sigops_queue* q, QSlot& s, size_t n, cudaStream_t st) {
    const size_t cap = q->cap;
    const Op op = q->op;
    CK(cudaMemcpyAsync(s.d_in, s.h_in, n * 64, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_in + cap * 64, s.h_in + cap * 64, n * 32, cudaMemcpyHostToDevice, st));
    if (op == OP_ED) CK(cudaMemcpyAsync(s.d_in + cap * 96, s.h_in + cap * 96, n * 32, cudaMemcpyHostToDevice, st));
    uint8_t* d_status = op == OP_ED ? nullptr : s.d_out + cap * 64;
    if (launch_op_scratch(*q->dev, op, s.d_in, s.d_in + cap * 64, s.d_in + cap * 96, n, s.d_out, d_status, s.scratch, st))
        return 1;
    CK(cudaMemcpyAsync(s.h_out, s.d_out, n * q_out_stride(op), cudaMemcpyDeviceToHost, st));
    if (d_status) CK(cudaMemcpyAsync(s.h_out + cap * 64, d_status, n, cudaMemcpyDeviceToHost, st));
    return 0;
}

void queue_free_slot(QSlot& s) {
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.st) cudaStreamDestroy(s.st);
    if (s.t0) cudaEventDestroy(s.t0);
    if (s.t1) cudaEventDestroy(s.t1);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.scratch) cudaFree(s.scratch);
    s = QSlot();
}

int queue_alloc_slot(sigops_queue* q, QSlot& s) {
    const Device& d = *q->dev;
    const size_t cap = q->cap;
    const int max_g = q->op == OP_K1 ? d.grid_k1 : q->op == OP_R1 ? d.grid_r1 : d.grid_ed;
    const size_t chunks = q->op == OP_ED ? kEdBatchChunks : kSwBatchChunks;
    // launch_geometry never uses more than min(max_g * 512, n + 512) threads for n signatures
    const size_t threads = std::min((size_t)max_g * kBlock, cap + (size_t)kBlock);
    CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s.t0));
    CK(cudaEventCreate(&s.t1));
    CK(cudaHostAlloc((void**)&s.h_in, cap * 128, cudaHostAllocPortable));
    CK(cudaHostAlloc((void**)&s.h_out, cap * 65, cudaHostAllocPortable));
    CK(cudaMalloc((void**)&s.d_in, cap * 128 + 64));
    CK(cudaMalloc((void**)&s.d_out, cap * 65 + 64));
    CK(cudaMalloc((void**)&s.scratch, chunks * threads * sizeof(Q4)));
    return 0;
}

}  // namespace

extern "C" {

int sigops_init(const int* device_ids, int n_devices) {
    std::lock_guard<std::mutex> lk(g_mu);
    return do_init(device_ids, n_devices);
}

int sigops_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_live_queues.load() > 0) {
        set_err("sigops_shutdown: destroy every sigops_queue first");
        return 1;
    }
    for (auto& d : g_dev) {
        cudaSetDevice(d.id);
        if (d.d_in) cudaFree(d.d_in);
        if (d.d_out) cudaFree(d.d_out);
        if (d.scratch) cudaFree(d.scratch);
        if (d.k1g) cudaFree((void*)d.k1g);
        if (d.r1g) cudaFree((void*)d.r1g);
        if (d.edb) cudaFree((void*)d.edb);
        for (auto& e : d.ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : d.ev_in)
            if (e) cudaEventDestroy(e);
        for (auto& e : d.ev_k)
            if (e) cudaEventDestroy(e);
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.s_in) cudaStreamDestroy(d.s_in);
        if (d.s_out) cudaStreamDestroy(d.s_out);
    }
    g_dev.clear();
    g_inited = false;
    return 0;
}

int sigops_num_devices(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (do_init(nullptr, 0)) return 0;
    return (int)g_dev.size();
}

const char* sigops_last_error(void) {
    // a stable copy: the shared string may be rewritten by a concurrent failing call
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mu);
    copy = g_err;
    return copy.c_str();
}

int sigops_secp256k1_ecrecover(const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out_pubkeys,
                               uint8_t* out_status) {
    return run_batch(OP_K1, sigs, msgs, nullptr, n, out_pubkeys, out_status);
}

int sigops_secp256r1_ecrecover(const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out_pubkeys,
                               uint8_t* out_status) {
    return run_batch(OP_R1, sigs, msgs, nullptr, n, out_pubkeys, out_status);
}

int sigops_ed25519_ecverify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n,
                            uint8_t* out_valid) {
    return run_batch(OP_ED, sigs, msgs, pks, n, out_valid, nullptr);
}

// Variable-length-message ed25519: one device shard = signatures [lo, hi) and the message bytes they span, through the same
// piecewise upload / kernel / download pipeline as the fixed-size entry points.
static int run_ed_msgs_shard(Device& d, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, const uint8_t* pks,
                             size_t n, uint32_t flags, uint8_t* out) {
    CK(cudaSetDevice(d.id));
    const uint64_t b0 = off[0], nbytes = off[n] - off[0];
    // layout: sigs | pks | offsets (n+1, rebased to 0) | message bytes
    const size_t off_bytes = (n + 1) * sizeof(uint64_t);
    const size_t in_bytes = n * 96 + off_bytes + (size_t)nbytes + 64;
    if (ensure_buf(&d.d_in, &d.in_cap, in_bytes)) return 1;
    if (ensure_buf(&d.d_out, &d.out_cap, n + 64)) return 1;
    if (ensure_scratch(d, (size_t)kEdTabChunks * d.grid_edm * kBlock)) return 1;
    uint8_t* d_sigs = d.d_in;
    uint8_t* d_pks = d.d_in + n * 64;
    uint8_t* d_off = d.d_in + n * 96;
    uint8_t* d_msg = d_off + off_bytes;
    std::vector<uint64_t> rebased(n + 1);
    for (size_t i = 0; i <= n; i++) {
        if (off[i] < b0 || (i && off[i] < off[i - 1])) {
            set_err("sigops_ed25519_ecverify_msgs: msg_offsets must be non-decreasing");
            return 1;
        }
        rebased[i] = off[i] - b0;
    }
    // the offsets go up first, in one piece (8 bytes per signature); `rebased` stays alive until the pipeline has drained
    CK(cudaMemcpyAsync(d_off, rebased.data(), off_bytes, cudaMemcpyHostToDevice, d.s_in));
    auto up = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_pks + lo * 32, pks + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        const uint64_t bl = rebased[lo], bh = rebased[lo + m];
        if (bh > bl) CK(cudaMemcpyAsync(d_msg + bl, msg_bytes + b0 + bl, (size_t)(bh - bl), cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto launch = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        int tpb = kBlock;
        if (m < (size_t)d.sms * kBlock) {
            size_t per_sm = (m + d.sms - 1) / d.sms;
            tpb = (int)std::min<size_t>(kBlock, std::max<size_t>(32, (per_sm + 31) / 32 * 32));
        }
        int grid = (int)std::min<size_t>((m + tpb - 1) / tpb, (size_t)d.grid_edm);
        ed25519_verify_msgs_kernel<<<grid, tpb, 0, st>>>((const Q4*)(d_sigs + lo * 64), d_msg,
                                                         (const unsigned long long*)d_off + lo, (const Q4*)(d_pks + lo * 32), m,
                                                         (int)(flags & SIGOPS_ED25519_STRICT), d.d_out + lo, d.scratch, d.edb);
        CK(cudaGetLastError());
        g_launches++;
        return 0;
    };
    auto down = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(out + lo, d.d_out + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    return run_pipeline(d, n, up, launch, down);
}

int sigops_ed25519_ecverify_msgs(const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                                 const uint8_t* pks, size_t n, uint32_t flags, uint8_t* out_valid) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n == 0) return 0;
    if (!sigs || !msg_offsets || !pks || !out_valid || (!msg_bytes && msg_offsets[n] != msg_offsets[0])) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = do_init(nullptr, 0)) return rc;
    for (auto& d : g_dev) d.ms_h2d = d.ms_kernel = d.ms_d2h = 0;
    std::vector<size_t> bounds(g_dev.size() + 1);
    const size_t G = (size_t)plan_shards(n, (int)g_dev.size(), bounds.data());
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        const size_t lo = bounds[g], hi = bounds[g + 1];
        auto work = [&, g, lo, hi]() {
            // sub-shards bound the device buffers as in run_shard_bounded
            for (size_t a = lo; a < hi && !rcs[g]; a += kMaxSubShard) {
                const size_t m = std::min(kMaxSubShard, hi - a);
                rcs[g] = run_ed_msgs_shard(g_dev[g], sigs + a * 64, msg_bytes, msg_offsets + a, pks + a * 32, m, flags,
                                           out_valid + a);
            }
        };
        if (G == 1)
            work();
        else
            th.emplace_back(work);
    }
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return rcs[g];
    return 0;
}

// raw messages -> SHA-256 -> recover -> SHA-256(pubkey), all on the device, piece by piece; no host pass in between
static int run_addresses_shard(Device& d, Op op, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, size_t n,
                               uint8_t* out_addr, uint8_t* out_pk, uint8_t* out_st) {
    CK(cudaSetDevice(d.id));
    const uint64_t b0 = off ? off[0] : 0, nbytes = off ? off[n] - off[0] : 0;
    const size_t off_bytes = off ? (n + 1) * sizeof(uint64_t) : 0;
    // d_in: sigs | prehashes (n*32) | offsets | raw bytes        d_out: pubkeys (n*64) | status (n) | pad | addresses
    const size_t in_bytes = n * 96 + off_bytes + (size_t)nbytes + 64;
    const size_t addr_off = (n * 65 + 63) / 64 * 64;
    if (ensure_buf(&d.d_in, &d.in_cap, in_bytes)) return 1;
    if (ensure_buf(&d.d_out, &d.out_cap, addr_off + n * 32 + 64)) return 1;
    uint8_t* d_sigs = d.d_in;
    uint8_t* d_msgs = d.d_in + n * 64;
    uint8_t* d_off = d.d_in + n * 96;
    uint8_t* d_raw = d_off + off_bytes;
    uint8_t* d_status = d.d_out + n * 64;
    uint8_t* d_addr = d.d_out + addr_off;
    std::vector<uint64_t> rebased(off ? n + 1 : 0);
    if (off) {
        for (size_t i = 0; i <= n; i++) {
            if (off[i] < b0 || (i && off[i] < off[i - 1])) {
                set_err("sigops_ecrecover_addresses: msg_offsets must be non-decreasing");
                return 1;
            }
            rebased[i] = off[i] - b0;
        }
        CK(cudaMemcpyAsync(d_off, rebased.data(), off_bytes, cudaMemcpyHostToDevice, d.s_in));
    }
    auto up = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        if (off) {
            const uint64_t bl = rebased[lo], bh = rebased[lo + m];
            if (bh > bl) CK(cudaMemcpyAsync(d_raw + bl, msg_bytes + b0 + bl, (size_t)(bh - bl), cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMemcpyAsync(d_msgs + lo * 32, msg_bytes + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        }
        return 0;
    };
    auto launch = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        if (off) {
            sha256_msgs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_raw, (const unsigned long long*)d_off + lo, m,
                                                                           (u32*)(d_msgs + lo * 32));
            CK(cudaGetLastError());
            g_launches++;
        }
        if (launch_op(d, op, d_sigs + lo * 64, d_msgs + lo * 32, nullptr, m, d.d_out + lo * 64, d_status + lo, st)) return 1;
        sha256_pubkeys_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>((const u32*)(d.d_out + lo * 64), d_status + lo, m,
                                                                          (u32*)(d_addr + lo * 32));
        CK(cudaGetLastError());
        g_launches++;
        return 0;
    };
    auto down = [&](size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(out_addr + lo * 32, d_addr + lo * 32, m * 32, cudaMemcpyDeviceToHost, st));
        if (out_pk) CK(cudaMemcpyAsync(out_pk + lo * 64, d.d_out + lo * 64, m * 64, cudaMemcpyDeviceToHost, st));
        if (out_st) CK(cudaMemcpyAsync(out_st + lo, d_status + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    return run_pipeline(d, n, up, launch, down);
}

int sigops_ecrecover_addresses(int curve, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                               size_t n, uint8_t* out_addresses, uint8_t* out_pubkeys, uint8_t* out_status) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n == 0) return 0;
    if (curve != SIGOPS_CURVE_SECP256K1 && curve != SIGOPS_CURVE_SECP256R1) {
        set_err("sigops_ecrecover_addresses: curve must be secp256k1 or secp256r1");
        return 1;
    }
    if (!sigs || !out_addresses || (!msg_bytes && (!msg_offsets || msg_offsets[n] != msg_offsets[0]))) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = do_init(nullptr, 0)) return rc;
    for (auto& d : g_dev) d.ms_h2d = d.ms_kernel = d.ms_d2h = 0;
    const Op op = curve == SIGOPS_CURVE_SECP256K1 ? OP_K1 : OP_R1;
    std::vector<size_t> bounds(g_dev.size() + 1);
    const size_t G = (size_t)plan_shards(n, (int)g_dev.size(), bounds.data());
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        const size_t lo = bounds[g], hi = bounds[g + 1];
        auto work = [&, g, lo, hi]() {
            for (size_t a = lo; a < hi && !rcs[g]; a += kMaxSubShard) {
                const size_t m = std::min(kMaxSubShard, hi - a);
                rcs[g] = run_addresses_shard(g_dev[g], op, sigs + a * 64, msg_offsets ? msg_bytes : msg_bytes + a * 32,
                                             msg_offsets ? msg_offsets + a : nullptr, m, out_addresses + a * 32,
                                             out_pubkeys ? out_pubkeys + a * 64 : nullptr, out_status ? out_status + a : nullptr);
            }
        };
        if (G == 1)
            work();
        else
            th.emplace_back(work);
    }
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return rcs[g];
    return 0;
}

int sigops_sha256_batch(const uint8_t* data, const uint64_t* offsets, size_t n, uint8_t* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n == 0) return 0;
    if (!offsets || !out || (!data && offsets[n] != offsets[0])) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = do_init(nullptr, 0)) return rc;
    Device& d = g_dev[0];
    CK(cudaSetDevice(d.id));
    for (size_t a = 0; a < n; a += kMaxSubShard) {
        const size_t m = std::min(kMaxSubShard, n - a);
        const uint64_t b0 = offsets[a], nbytes = offsets[a + m] - b0;
        const size_t off_bytes = (m + 1) * sizeof(uint64_t);
        if (ensure_buf(&d.d_in, &d.in_cap, off_bytes + (size_t)nbytes + 64)) return 1;
        if (ensure_buf(&d.d_out, &d.out_cap, m * 32 + 64)) return 1;
        std::vector<uint64_t> rebased(m + 1);
        for (size_t i = 0; i <= m; i++) {
            if (offsets[a + i] < b0 || (i && offsets[a + i] < offsets[a + i - 1])) {
                set_err("sigops_sha256_batch: offsets must be non-decreasing");
                return 1;
            }
            rebased[i] = offsets[a + i] - b0;
        }
        CK(cudaMemcpyAsync(d.d_in, rebased.data(), off_bytes, cudaMemcpyHostToDevice, d.stream));
        if (nbytes) CK(cudaMemcpyAsync(d.d_in + off_bytes, data + b0, (size_t)nbytes, cudaMemcpyHostToDevice, d.stream));
        sha256_msgs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, d.stream>>>(d.d_in + off_bytes,
                                                                             (const unsigned long long*)d.d_in, m, (u32*)d.d_out);
        CK(cudaGetLastError());
        g_launches++;
        CK(cudaMemcpyAsync(out + a * 32, d.d_out, m * 32, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    return 0;
}

int sigops_secp256k1_ecrecover_device(const void* d_sigs, const void* d_msgs, size_t n, void* d_out, void* d_status,
                                      void* stream) {
    return run_device(OP_K1, d_sigs, d_msgs, nullptr, n, d_out, d_status, stream);
}
int sigops_secp256r1_ecrecover_device(const void* d_sigs, const void* d_msgs, size_t n, void* d_out, void* d_status,
                                      void* stream) {
    return run_device(OP_R1, d_sigs, d_msgs, nullptr, n, d_out, d_status, stream);
}
int sigops_ed25519_ecverify_device(const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n, void* d_out,
                                   void* stream) {
    return run_device(OP_ED, d_sigs, d_msgs, d_pks, n, d_out, nullptr, stream);
}

int sigops_plan_shards(size_t n, int n_devices, size_t* bounds, int* n_used) {
    if (!bounds || !n_used || n_devices < 1) {
        set_err("sigops_plan_shards: bad arguments");
        return 1;
    }
    *n_used = plan_shards(n, n_devices, bounds);
    return 0;
}

int sigops_last_timing(double* h2d_ms, double* kernel_ms, double* d2h_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    double a = 0, b = 0, c = 0;
    for (auto& d : g_dev) {
        a = std::max(a, (double)d.ms_h2d);
        b = std::max(b, (double)d.ms_kernel);
        c = std::max(c, (double)d.ms_d2h);
    }
    if (h2d_ms) *h2d_ms = a;
    if (kernel_ms) *kernel_ms = b;
    if (d2h_ms) *d2h_ms = c;
    return 0;
}

uint64_t sigops_kernel_launches(void) { return g_launches.load(); }

void* sigops_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        set_err("cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}

void sigops_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int sigops_precompute_bases(int curve, uint32_t log_limb_size, uint32_t* out, size_t* inout_len) {
    if (log_limb_size < 11 || log_limb_size > 15 || !inout_len) {
        set_err("log_limb_size must be in 11..=15");
        return 1;
    }
    std::vector<uint32_t> v;
    const int num_limbs = calc_num_limbs((int)log_limb_size);
    if (curve == SIGOPS_CURVE_SECP256K1) {
        const u32 P[8] = SG_K1_P;
        sw_bases<CurveK1>(v, k1_g_host, P, (int)log_limb_size);
    } else if (curve == SIGOPS_CURVE_SECP256R1) {
        const u32 P[8] = SG_R1_P;
        sw_bases<CurveR1>(v, r1_g_host, P, (int)log_limb_size);
    } else if (curve == SIGOPS_CURVE_ED25519) {
        const u32 P[8] = SG_ED_P;
        Fe bx, by;
        copy8(bx.v, ed_b_host);
        copy8(by.v, ed_b_host + 8);
        EdPoint acc, B;
        B.X = bx;
        B.Y = by;
        FE::set_one(B.Z);
        FE::mul(B.T, bx, by);
        const Fe d2 = {SG_ED_D2};
        Fe ypx, ymx, t2d;
        FE::add(ypx, by, bx);
        FE::sub(ymx, by, bx);
        FE::mul(t2d, B.T, d2);
        ed_set_identity(acc);
        for (int i = 0; i < 16; i++) {
            ed_add_niels<FE>(acc, ypx, ymx, t2d, false, true);
            Fe zi, ax, ay, at;
            fe_inv((FE*)0, zi, acc.Z);
            FE::mul(ax, acc.X, zi);
            FE::mul(ay, acc.Y, zi);
            FE::mul(at, ax, ay);
            to_limbs_mont<FE>(v, ax, P, num_limbs, (int)log_limb_size);
            to_limbs_mont<FE>(v, ay, P, num_limbs, (int)log_limb_size);
            to_limbs_mont<FE>(v, at, P, num_limbs, (int)log_limb_size);
        }
    } else {
        set_err("unknown curve");
        return 1;
    }
    if (!out || *inout_len < v.size()) {
        *inout_len = v.size();
        if (out) {
            set_err("output buffer too small");
            return 1;
        }
        return 0;
    }
    memcpy(out, v.data(), v.size() * sizeof(uint32_t));
    *inout_len = v.size();
    return 0;
}

int sigops_test_unit_shape(int op, int* in_words, int* out_words) {
    if (op < 0 || op >= SIGOPS_UNIT_COUNT || !in_words || !out_words) return 1;
    unit_shape(op, *in_words, *out_words);
    return 0;
}

int sigops_test_unit(int op, const uint32_t* in, size_t n, uint32_t* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (op < 0 || op >= SIGOPS_UNIT_COUNT) {
        set_err("bad unit op");
        return 1;
    }
    if (n == 0) return 0;
    if (int rc = do_init(nullptr, 0)) return rc;
    Device& d = g_dev[0];
    CK(cudaSetDevice(d.id));
    int in_w, out_w;
    unit_shape(op, in_w, out_w);
    if (ensure_buf(&d.d_in, &d.in_cap, n * in_w * 4)) return 1;
    if (ensure_buf(&d.d_out, &d.out_cap, n * out_w * 4)) return 1;
    if (ensure_scratch(d, (size_t)kEdTabChunks * d.grid_unit * kBlock)) return 1;
    CK(cudaMemcpyAsync(d.d_in, in, n * in_w * 4, cudaMemcpyHostToDevice, d.stream));
    int grid = (int)std::min<size_t>((n + kBlock - 1) / kBlock, (size_t)d.grid_unit);
    unit_kernel<<<grid, kBlock, 0, d.stream>>>(op, (const u32*)d.d_in, n, (u32*)d.d_out, d.scratch, d.k1g, d.r1g, d.edb);
    CK(cudaGetLastError());
    g_launches++;
    CK(cudaMemcpyAsync(out, d.d_out, n * out_w * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return 0;
}

int sigops_imad_peak(int kind, int iters, double* ops_per_sec, double* ms_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = do_init(nullptr, 0)) return rc;
    Device& d = g_dev[0];
    CK(cudaSetDevice(d.id));
    if (ensure_buf(&d.d_out, &d.out_cap, (size_t)d.sms * 8 * 256 * 4)) return 1;
    const int grid = d.sms * 8, block = 256;  // 2048 threads per SM: full occupancy
    double per_thread_iter;
    for (int rep = 0; rep < 2; rep++) {  // first pass warms up
        CK(cudaEventRecord(d.ev[0], d.stream));
        switch (kind) {
            case 0: imad_peak_kernel<0><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 1: imad_peak_kernel<1><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 2: imad_peak_kernel<2><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 3: imad_peak_kernel<3><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 4: imad_peak_kernel<4><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 5: imad_peak_kernel<5><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            case 6: imad_peak_kernel<6><<<grid, block, 0, d.stream>>>((u32*)d.d_out, iters, 12345u); break;
            default: set_err("bad kind"); return 1;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(d.ev[1], d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
    // counted operations per thread per outer iteration (16 unrolled blocks):
    //  kind 0: 8 IMAD; 1: 8 IMAD.WIDE; 2: 8 IMAD.WIDE(.X); 3: 8 IADD; 4: 8 IMAD.WIDE + 8 IADD (counted: the 8 wide);
    //  5: 8 DFMA; 6: 8 IMAD.HI
    per_thread_iter = 16.0 * 8.0;
    double ops = per_thread_iter * (double)iters * (double)grid * block;
    if (ops_per_sec) *ops_per_sec = ops / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}


// ---- streaming service mode: C entry points ----------------------------------------------------------------------------
int sigops_queue_create(int curve, int device_index, size_t max_batch, int depth, sigops_queue** out) {
    if (!out) {
        set_err("sigops_queue_create: out is NULL");
        return 1;
    }
    *out = nullptr;
    if (curve < SIGOPS_CURVE_SECP256K1 || curve > SIGOPS_CURVE_ED25519 || max_batch == 0 || max_batch > kMaxSubShard ||
        depth < 1 || depth > SIGOPS_QUEUE_MAX_DEPTH) {
        set_err("sigops_queue_create: bad arguments (curve 0..2, 1 <= max_batch <= 2^24, 1 <= depth <= 64)");
        return 1;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = do_init(nullptr, 0)) return rc;
    if (device_index < 0 || device_index >= (int)g_dev.size()) {
        set_err("sigops_queue_create: device_index is not in the pool");
        return 1;
    }
    sigops_queue* q = new sigops_queue();
    q->dev = &g_dev[device_index];
    q->op = (Op)curve;
    q->cap = max_batch;
    if (const char* e = getenv("SIGOPS_QUEUE_GRAPHS")) q->graphs = atoi(e) != 0;
    q->slots.resize(depth);
    if (cudaSetDevice(q->dev->id) != cudaSuccess) {
        set_err("sigops_queue_create: cudaSetDevice failed");
        delete q;
        return 1;
    }
    for (auto& s : q->slots)
        if (queue_alloc_slot(q, s)) {
            for (auto& t : q->slots) queue_free_slot(t);
            delete q;
            return 1;
        }
    g_live_queues++;
    *out = q;
    return 0;
}

int sigops_queue_destroy(sigops_queue* q) {
    if (!q) return 0;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        cudaSetDevice(q->dev->id);
        for (auto& s : q->slots) {
            if (s.st) cudaStreamSynchronize(s.st);
            queue_free_slot(s);
        }
    }
    delete q;
    g_live_queues--;
    return 0;
}

int sigops_queue_buffers(sigops_queue* q, int slot, uint8_t** sigs, uint8_t** msgs, uint8_t** pks, uint8_t** out,
                         uint8_t** status) {
    if (!q || slot < 0 || slot >= (int)q->slots.size()) {
        set_err("sigops_queue_buffers: bad queue or slot");
        return 1;
    }
    QSlot& s = q->slots[slot];
    if (sigs) *sigs = s.h_in;
    if (msgs) *msgs = s.h_in + q->cap * 64;
    if (pks) *pks = q->op == OP_ED ? s.h_in + q->cap * 96 : nullptr;
    if (out) *out = s.h_out;
    if (status) *status = q->op == OP_ED ? nullptr : s.h_out + q->cap * 64;
    return 0;
}

int sigops_queue_submit(sigops_queue* q, int slot, size_t n) {
    if (!q || slot < 0 || slot >= (int)q->slots.size()) {
        set_err("sigops_queue_submit: bad queue or slot");
        return 1;
    }
    std::lock_guard<std::mutex> lk(q->mu);
    QSlot& s = q->slots[slot];
    if (s.busy) {
        set_err("sigops_queue_submit: slot is in flight (wait for it first)");
        return 1;
    }
    if (n > q->cap) {
        set_err("sigops_queue_submit: n exceeds the queue's max_batch");
        return 1;
    }
    s.n = n;
    s.last_ms = 0;
    if (n == 0) return 0;  // nothing to do: the slot stays free (src/secp256k1_ecdsa.rs:71-73)
    CK(cudaSetDevice(q->dev->id));
    if (q->graphs && s.exec_n != n) {
        // (re)capture the slot's copy / kernel / copy sequence for this request size
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
        t_capturing = true;
        int rc = queue_enqueue(q, s, n, s.st);
        t_capturing = false;
        cudaError_t e = cudaStreamEndCapture(s.st, &g);
        if (rc || e != cudaSuccess) {
            if (g) cudaGraphDestroy(g);
            if (!rc) set_err(std::string("cudaStreamEndCapture failed: ") + cudaGetErrorString(e));
            return 1;
        }
        bool updated = false;
        if (s.exec) {
            cudaGraphExecUpdateResultInfo info;
            updated = cudaGraphExecUpdate(s.exec, g, &info) == cudaSuccess;
            if (!updated) {
                cudaGetLastError();
                cudaGraphExecDestroy(s.exec);
                s.exec = nullptr;
            }
        }
        if (!updated) {
            e = cudaGraphInstantiate(&s.exec, g, 0);
            if (e != cudaSuccess) {
                cudaGraphDestroy(g);
                s.exec = nullptr;
                s.exec_n = 0;
                set_err(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(e));
                return 1;
            }
        }
        cudaGraphDestroy(g);
        s.exec_n = n;
        q->graph_captures++;
    }
    CK(cudaEventRecord(s.t0, s.st));
    if (q->graphs) {
        CK(cudaGraphLaunch(s.exec, s.st));
        g_launches += tail_split(*q->dev, n) < n ? 2 : 1;  // the fused kernel (twice when the tail is launched on its own)
        q->graph_launches++;
    } else if (queue_enqueue(q, s, n, s.st)) {
        return 1;
    }
    CK(cudaEventRecord(s.t1, s.st));
    s.busy = true;
    return 0;
}

int sigops_queue_poll(sigops_queue* q, int slot, int* done) {
    if (!q || slot < 0 || slot >= (int)q->slots.size() || !done) {
        set_err("sigops_queue_poll: bad arguments");
        return 1;
    }
    std::lock_guard<std::mutex> lk(q->mu);
    QSlot& s = q->slots[slot];
    if (!s.busy) {
        *done = 1;
        return 0;
    }
    cudaError_t e = cudaEventQuery(s.t1);
    if (e == cudaErrorNotReady) {
        *done = 0;
        return 0;
    }
    if (e != cudaSuccess) {
        set_err(std::string("sigops_queue_poll: ") + cudaGetErrorString(e));
        return 1;
    }
    *done = 1;
    return 0;
}

int sigops_queue_wait(sigops_queue* q, int slot, size_t* n_done, double* device_ms) {
    if (!q || slot < 0 || slot >= (int)q->slots.size()) {
        set_err("sigops_queue_wait: bad queue or slot");
        return 1;
    }
    QSlot& s = q->slots[slot];
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        if (!s.busy) {
            if (n_done) *n_done = s.n;
            if (device_ms) *device_ms = s.last_ms;
            return 0;
        }
        ev = s.t1;
    }
    CK(cudaEventSynchronize(ev));  // outside the queue lock: other threads keep submitting to other slots
    std::lock_guard<std::mutex> lk(q->mu);
    CK(cudaEventElapsedTime(&s.last_ms, s.t0, s.t1));
    s.busy = false;
    if (n_done) *n_done = s.n;
    if (device_ms) *device_ms = s.last_ms;
    return 0;
}

int sigops_queue_info(sigops_queue* q, int* curve, int* device_index, size_t* max_batch, int* depth, uint64_t* graph_launches,
                      uint64_t* graph_captures) {
    if (!q) {
        set_err("sigops_queue_info: queue is NULL");
        return 1;
    }
    std::lock_guard<std::mutex> lk(q->mu);
    if (curve) *curve = (int)q->op;
    if (device_index) *device_index = (int)(q->dev - g_dev.data());
    if (max_batch) *max_batch = q->cap;
    if (depth) *depth = (int)q->slots.size();
    if (graph_launches) *graph_launches = q->graph_launches;
    if (graph_captures) *graph_captures = q->graph_captures;
    return 0;
}

}  // extern "C"
