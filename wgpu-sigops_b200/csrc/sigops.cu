// libsigops host layer: persistent device pool, contiguous sharding across GPUs, pinned-aware transfers and the
// extern "C" entry points declared in include/sigops.h.
//
// Replaces src/gpu.rs:5-170 (per-call wgpu device creation, storage/uniform buffers, pipeline creation from a WGSL
// string, dispatch, staging-buffer read-back and device.destroy()) and the host halves of
// src/secp256k1_ecdsa.rs:11-213, src/secp256r1_ecdsa.rs:12-214, src/ed25519_eddsa.rs:12-257 (flatten, zero-pad to a
// power of two, five to seven dispatches, slice the padded result).  Here: no padding, one kernel launch per piece of a
// shard, device context and buffers persist across calls, and there is no collective -- shards are independent
// (SURVEY.md 8e).  There is no CPU fallback: without a CUDA device every compute entry point returns nonzero.
//
// Concurrency model: the pool (device list) is guarded by a reader/writer lock -- compute calls hold it shared, init /
// shutdown exclusively.  Every device has its own mutex (one host pipeline or one enqueue at a time per device), a
// persistent worker thread that runs the device's shard of a multi-device call, and a few copy threads that move
// pageable caller buffers through pinned staging.  Two callers on disjoint devices never wait for each other.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "group.cuh"
#include "launch.h"

using namespace sigops;

namespace {

std::shared_mutex g_pool_mu;  // shared: compute calls; exclusive: init / shutdown
std::mutex g_err_mu;          // shard threads may report errors concurrently
std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_call_seq{0};
thread_local uint64_t t_last_call = 0;  // sigops_last_timing reports the calling thread's last host-buffer call
thread_local bool t_capturing = false;  // stream capture records a launch, it does not perform one

void set_err(const std::string& s) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = s;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (cudaError_t)(call);                                                          \
        if (e_ != cudaSuccess) {                                                                       \
            char b_[512];                                                                              \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            set_err(b_);                                                                               \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

// counts outstanding jobs; wait() blocks until all are done
struct Counter {
    std::mutex mu;
    std::condition_variable cv;
    size_t pending = 0;
    void add(size_t k) {
        std::lock_guard<std::mutex> lk(mu);
        pending += k;
    }
    void done() {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return pending == 0; });
    }
};

struct CopyJob {
    void* dst;
    const void* src;
    size_t bytes;
    Counter* c;
};

constexpr int kMaxChunks = 8;

struct Device {
    int id = -1;     // CUDA ordinal
    int index = -1;  // position in the pool
    int sms = 0;
    std::mutex mu;   // one host pipeline / one enqueue at a time on this device
    std::atomic<int> inflight{0};
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;  // kernels / uploads / downloads
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {}, ev_down[kMaxChunks] = {};
    cudaEvent_t scratch_ev = nullptr;  // last launch that used `scratch`: the next one (any stream) waits for it
    bool scratch_used = false;
    // device buffers (grow-only)
    uint8_t* d_in = nullptr;
    size_t in_cap = 0;
    uint8_t* d_out = nullptr;
    size_t out_cap = 0;
    Q4* scratch = nullptr;
    size_t scratch_cap = 0;  // in Q4
    // pinned staging for pageable caller buffers (grow-only)
    uint8_t* h_in = nullptr;
    size_t h_in_cap = 0;
    uint8_t* h_out = nullptr;
    size_t h_out_cap = 0;
    PTab k1g{}, r1g{}, edb{};  // positional fixed-base tables (ptab.h), resident for the life of the context
    int grid_k1 = 0, grid_r1 = 0, grid_ed = 0, grid_edm = 0, grid_unit = 0;
    int ggrid_k1 = 0, ggrid_r1 = 0, ggrid_ed = 0;  // resident blocks of the lane-group kernels
    float ms_h2d = 0, ms_kernel = 0, ms_d2h = 0;
    uint64_t last_call = 0;
    size_t table_bytes = 0;
    // persistent worker: runs this device's shard of a multi-device call (no std::thread spawn / join per call)
    std::thread worker;
    std::mutex wmu;
    std::condition_variable wcv;
    std::deque<std::function<void()>> wq;
    bool wstop = false;
    // copy threads: pageable caller memory <-> pinned staging
    std::vector<std::thread> copiers;
    std::mutex cmu;
    std::condition_variable ccv;
    std::deque<CopyJob> cq;
    bool cstop = false;

    void post(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(wmu);
            wq.push_back(std::move(f));
        }
        wcv.notify_one();
    }
    void worker_main() {
        cudaSetDevice(id);
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(wmu);
                wcv.wait(lk, [&] { return wstop || !wq.empty(); });
                if (wq.empty()) return;
                f = std::move(wq.front());
                wq.pop_front();
            }
            f();
        }
    }
    // copy [src, src + bytes) to dst in slices of at most kCopySlice, spread over the copy threads
    static constexpr size_t kCopySlice = (size_t)1 << 20;
    void post_copy(void* dst, const void* src, size_t bytes, Counter* c) {
        if (bytes == 0) return;
        const size_t slices = (bytes + kCopySlice - 1) / kCopySlice;
        c->add(slices);
        {
            std::lock_guard<std::mutex> lk(cmu);
            for (size_t s = 0; s < slices; s++) {
                const size_t off = s * kCopySlice;
                cq.push_back({(uint8_t*)dst + off, (const uint8_t*)src + off, std::min(kCopySlice, bytes - off), c});
            }
        }
        ccv.notify_all();
    }
    void copier_main() {
        for (;;) {
            CopyJob j;
            {
                std::unique_lock<std::mutex> lk(cmu);
                ccv.wait(lk, [&] { return cstop || !cq.empty(); });
                if (cq.empty()) return;
                j = cq.front();
                cq.pop_front();
            }
            memcpy(j.dst, j.src, j.bytes);
            j.c->done();
        }
    }
    void stop_threads() {
        {
            std::lock_guard<std::mutex> lk(wmu);
            wstop = true;
        }
        wcv.notify_all();
        if (worker.joinable()) worker.join();
        {
            std::lock_guard<std::mutex> lk(cmu);
            cstop = true;
        }
        ccv.notify_all();
        for (auto& t : copiers)
            if (t.joinable()) t.join();
        copiers.clear();
    }
};

// The pool is heap-allocated and never destroyed at process exit (its threads may still be parked on their condition
// variables when static destructors run); sigops_shutdown() tears it down explicitly.
std::vector<std::unique_ptr<Device>>& pool() {
    static auto* p = new std::vector<std::unique_ptr<Device>>();
    return *p;
}
std::atomic<bool> g_inited{false};
std::vector<int> g_ids;  // CUDA ordinals the pool was initialised with

// Pinned allocations handed out by sigops_host_alloc (and owned by queue slots): known to be page-locked and mapped into
// every device's address space (unified addressing), so small requests can be read and written by the kernels directly --
// no driver query per call.
std::mutex g_pin_mu;
std::vector<std::pair<uintptr_t, size_t>> g_pinned;  // sorted by address
void pinned_add(void* p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto e = std::make_pair((uintptr_t)p, bytes);
    g_pinned.insert(std::upper_bound(g_pinned.begin(), g_pinned.end(), e), e);
}
void pinned_remove(void* p) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pinned.size(); i++)
        if (g_pinned[i].first == (uintptr_t)p) {
            g_pinned.erase(g_pinned.begin() + i);
            return;
        }
}
bool pinned_ours(const void* p, size_t bytes) {
    if (!p) return false;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = std::upper_bound(g_pinned.begin(), g_pinned.end(), std::make_pair((uintptr_t)p, (size_t)-1));
    if (it == g_pinned.begin()) return false;
    --it;
    return (uintptr_t)p >= it->first && (uintptr_t)p + bytes <= it->first + it->second;
}

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

int ensure_pinned(uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) CK(cudaFreeHost(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 4096;
    CK(cudaHostAlloc((void**)p, want, cudaHostAllocPortable));
    *cap = want;
    return 0;
}

int ensure_scratch(Device& d, size_t q4s) {
    if (q4s <= d.scratch_cap) return 0;
    if (d.scratch) CK(cudaFree(d.scratch));  // cudaFree synchronises the device: no launch still uses the old scratch
    d.scratch = nullptr;
    d.scratch_cap = 0;
    d.scratch_used = false;
    CK(cudaMalloc((void**)&d.scratch, q4s * sizeof(Q4)));
    d.scratch_cap = q4s;
    return 0;
}

void free_device(Device& d) {
    d.stop_threads();
    if (d.id < 0) return;
    cudaSetDevice(d.id);
    if (d.d_in) cudaFree(d.d_in);
    if (d.d_out) cudaFree(d.d_out);
    if (d.scratch) cudaFree(d.scratch);
    if (d.h_in) cudaFreeHost(d.h_in);
    if (d.h_out) cudaFreeHost(d.h_out);
    if (d.k1g.base) cudaFree((void*)d.k1g.base);
    if (d.r1g.base) cudaFree((void*)d.r1g.base);
    if (d.edb.base) cudaFree((void*)d.edb.base);
    for (auto& e : d.ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : d.ev_in)
        if (e) cudaEventDestroy(e);
    for (auto& e : d.ev_k)
        if (e) cudaEventDestroy(e);
    for (auto& e : d.ev_down)
        if (e) cudaEventDestroy(e);
    if (d.scratch_ev) cudaEventDestroy(d.scratch_ev);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.s_in) cudaStreamDestroy(d.s_in);
    if (d.s_out) cudaStreamDestroy(d.s_out);
}

int init_device(Device& d, int id, int index, int copy_threads) {
    d.id = id;
    d.index = index;
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
    for (auto& e : d.ev_down) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d.scratch_ev, cudaEventDisableTiming));
    // positional fixed-base tables (ptab.h): computed on the device once, resident for the life of the context
    {
        u32 w = kPTabDefaultWin;
        if (const char* e = getenv("SIGOPS_GWIN")) {
            const int v = atoi(e);
            if (v < (int)kPTabMinWin || v > (int)kPTabMaxWin) {
                char b[160];
                snprintf(b, sizeof b, "SIGOPS_GWIN=%s: the fixed-base window must be %u..%u bits", e, kPTabMinWin, kPTabMaxWin);
                set_err(b);
                return 1;
            }
            w = (u32)v;
        }
        const size_t entries = ptab_entries(w);
        u32* bases = nullptr;
        CK(cudaMalloc((void**)&bases, (size_t)ptab_positions(w) * 24 * sizeof(u32)));
        PTab* tabs[3] = {&d.k1g, &d.r1g, &d.edb};
        for (int c = 0; c < 3; c++) {
            u32* t = nullptr;
            const cudaError_t e = cudaMalloc((void**)&t, entries * (c == 2 ? 24 : 16) * sizeof(u32));
            if (e != cudaSuccess) {
                cudaFree(bases);
                char b[200];
                snprintf(b, sizeof b, "fixed-base table of %zu entries (SIGOPS_GWIN=%u) does not fit on device %d: %s", entries, w, id,
                         cudaGetErrorString(e));
                set_err(b);
                return 1;
            }
            ptab_describe(*tabs[c], t, w);
            CK(kl_gen_ptab(d.stream, c, w, bases, t));
            CK(cudaStreamSynchronize(d.stream));  // `bases` is reused by the next curve
            g_launches += 2;
        }
        CK(cudaFree(bases));
        d.table_bytes = entries * (16 + 16 + 24) * sizeof(u32);
    }
    int per_sm = 0, per_sm2 = 0;
    CK(kl_k1_setup(&per_sm));
    d.grid_k1 = std::max(per_sm, 1) * d.sms;
    CK(kl_r1_setup(&per_sm));
    d.grid_r1 = std::max(per_sm, 1) * d.sms;
    CK(kl_ed_setup(&per_sm, &per_sm2));
    d.grid_ed = std::max(per_sm, 1) * d.sms;
    d.grid_edm = std::max(per_sm2, 1) * d.sms;
    CK(kl_unit_setup(&per_sm));
    d.grid_unit = std::max(per_sm, 1) * d.sms;
    CK(kl_k1_group_cold_setup(&per_sm));
    CK(kl_ed_group_cold_setup(&per_sm));
    CK(kl_k1_group_setup(&per_sm));
    d.ggrid_k1 = std::max(per_sm, 1) * d.sms;
    CK(kl_r1_group_setup(&per_sm));
    d.ggrid_r1 = std::max(per_sm, 1) * d.sms;
    CK(kl_ed_group_setup(&per_sm));
    d.ggrid_ed = std::max(per_sm, 1) * d.sms;
    d.worker = std::thread([&d] { d.worker_main(); });
    for (int i = 0; i < copy_threads; i++) d.copiers.emplace_back([&d] { d.copier_main(); });
    return 0;
}

// caller holds g_pool_mu exclusively
int do_init_locked(const int* ids, int n) {
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
            if (std::find(use.begin(), use.end(), ids[i]) != use.end()) {
                set_err("sigops_init: duplicate device id");
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
    if (g_inited.load()) {
        // already up: an explicit device list must match the pool (a silent no-op would leave the caller believing it had
        // pinned devices it has not); lazy / default initialisation is always satisfied by the existing pool
        if (ids && n > 0 && use != g_ids) {
            set_err("sigops_init: the pool is already initialised with a different device set; call sigops_shutdown() first");
            return 1;
        }
        return 0;
    }
    int copy_threads = 0;
    if (const char* s = getenv("SIGOPS_COPY_THREADS")) copy_threads = atoi(s);
    if (copy_threads <= 0) {
        const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
        copy_threads = std::max(1, std::min(4, hw / (2 * (int)use.size())));
    }
    auto& P = pool();
    P.clear();
    for (size_t i = 0; i < use.size(); i++) {
        P.emplace_back(new Device());
        if (init_device(*P.back(), use[i], (int)i, copy_threads)) {
            for (auto& d : P) free_device(*d);  // release what the devices initialised so far hold
            P.clear();
            return 1;
        }
    }
    g_ids = use;
    g_inited.store(true);
    return 0;
}

// lazy initialisation with the defaults; cheap once the pool is up
int ensure_init() {
    if (g_inited.load()) return 0;
    std::unique_lock<std::shared_mutex> lk(g_pool_mu);
    return do_init_locked(nullptr, 0);
}

// device context for the CURRENT device (used by the *_device entry points); caller holds the pool lock (shared)
Device* current_device() {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return nullptr;
    for (auto& d : pool())
        if (d->id == cur) return d.get();
    return nullptr;
}

enum Op { OP_K1 = 0, OP_R1 = 1, OP_ED = 2 };

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Launch geometry: full 512-thread blocks (one per SM, all 16 warps phase-locked by the kernels' barriers) once the batch
// covers the device; smaller batches are spread over all SMs with proportionally smaller blocks instead of filling a few
// SMs to the brim (latency of the 64 ... 64k end of the batch-size sweep).
// A pass costs about 0.31 + 0.38 k ms (secp256k1) with k = warps per scheduler = ceil(warps per SM / 4): the block's
// barriers make the most loaded scheduler set the pace, so 13 and 16 warps per SM take the same time.  That is why
// balancing a shard of w.f waves over ceil(w.f) equal passes (SIGOPS_BALANCED=1: tpb = ceil(n / (SMs x P)) rounded to a
// warp) measured SLOWER than full passes plus a separate, smaller tail launch (profiles/r02_geometry.txt: 131,072
// signatures 3.41 vs 3.28 ms, 262,144 6.60 vs 6.11 ms); it stays selectable for the record, off by default.
void launch_geometry(const Device& d, Op op, size_t n, int* grid, int* tpb_out) {
    const int max_g = op == OP_K1 ? d.grid_k1 : op == OP_R1 ? d.grid_r1 : d.grid_ed;
    const size_t wave = (size_t)d.sms * kBlock;
    int tpb = kBlock;
    size_t passes = 1;
    if (n < wave) {
        size_t per_sm = (n + d.sms - 1) / d.sms;
        tpb = (int)std::min<size_t>(kBlock, std::max<size_t>(32, (per_sm + 31) / 32 * 32));
    } else if (env_int("SIGOPS_BALANCED", 0) != 0) {
        passes = (n + wave - 1) / wave;
        size_t per_sm = (n + (size_t)d.sms * passes - 1) / ((size_t)d.sms * passes);
        tpb = (int)std::min<size_t>(kBlock, std::max<size_t>(32, (per_sm + 31) / 32 * 32));
    }
    size_t blocks = (n + (size_t)tpb * passes - 1) / ((size_t)tpb * passes);
    if (passes == 1) blocks = (n + tpb - 1) / tpb;
    *grid = (int)std::min<size_t>(blocks, (size_t)max_g);
    *tpb_out = tpb;
}

// A launch of w.f waves (one wave = SMs x 512 signatures, one per thread) costs ceil(w.f) full passes: the last, partial
// pass runs every thread (rows past the end are clamped) at 16 warps per SM.  When the tail is at most
// kTailSplitNum/kTailSplitDen of a wave it is launched on its own right behind the main part, with the small-batch
// geometry: fewer warps per scheduler finish a pass sooner.  Matters for shards of one to a few waves -- the 1M block
// cut over 4 or 8 GPUs, the middle of the batch-size sweep.  SIGOPS_TAIL_SPLIT=0 disables it.
constexpr size_t kTailSplitNum = 3, kTailSplitDen = 4;
size_t tail_split(const Device& d, size_t n) {  // signatures in the main launch (n if no split)
    const size_t wave = (size_t)d.sms * kBlock;
    if (n < wave) return n;
    if (env_int("SIGOPS_TAIL_SPLIT", 1) == 0 || env_int("SIGOPS_BALANCED", 0) != 0) return n;
    const size_t tail = n % wave;
    if (tail == 0 || tail * kTailSplitDen > wave * kTailSplitNum) return n;
    return n - tail;
}

// Small batches run on the lane-group kernels (group.cuh: several cooperating warps per 32 signatures, no device scratch):
// below one block per SM the request's latency is one signature's dependent chain, which the group kernels cut roughly in
// half.  SIGOPS_LANEGROUP=0 disables them, SIGOPS_FORCE_LANEGROUP=1 uses them for every size (tests),
// SIGOPS_LANEGROUP_MAX=<n> moves the threshold.  Defaults from profiles/r02_latency*.txt: all three curves win while the
// request fits ONE 32-signature block per SM (4,736 signatures: 0.40 / 0.79 / 0.47 ms against 0.80 / 1.12 / 0.93) -- with two
// blocks per SM the roles no longer have a scheduler to themselves and the one-thread-per-signature kernel is as fast.
// Above kGroupColdMin signatures secp256k1 and ed25519 switch to the flavour with out-of-line field products (launch_group):
// inlined wins up to 3,072 signatures (0.391 / 0.456 ms against 0.396 / 0.461), out of line from 4,096 (0.400 / 0.469 against
// 0.405 / 0.476; 4,736: 0.404 / 0.471 against 0.421 / 0.497) -- profiles/r02_group_flavours.txt.
constexpr int kGroupColdMin = 3584;
bool use_group_kernel(const Device& d, Op op, size_t n) {
    if (env_int("SIGOPS_FORCE_LANEGROUP", 0) != 0) return true;
    if (env_int("SIGOPS_LANEGROUP", 1) == 0) return false;
    const int dflt = d.sms * kGroupSigs;
    const int lim = env_int("SIGOPS_LANEGROUP_MAX", dflt);
    return n <= (size_t)std::max(lim, 0);
}

int launch_group(Device& d, Op op, const uint8_t* d_sigs, const uint8_t* d_msgs, const uint8_t* d_pks, size_t n, uint8_t* d_out,
                 uint8_t* d_status, cudaStream_t st) {
    KLaunch l;
    l.stream = st;
    l.tpb = 0;
    const int resident = op == OP_K1 ? d.ggrid_k1 : op == OP_R1 ? d.ggrid_r1 : d.ggrid_ed;
    l.grid = (int)std::min<size_t>((n + kGroupSigs - 1) / kGroupSigs, (size_t)resident);
    // flavour: field products inlined into the role programs while the request occupies a minority of the SMs, out of line
    // (smaller programs, less instruction fetch through L2) beyond SIGOPS_GROUP_COLD_MIN signatures; P-256 always out of line
    const bool cold = n > (size_t)std::max(0, env_int("SIGOPS_GROUP_COLD_MIN", kGroupColdMin));
    switch (op) {
        case OP_K1:
            if (cold)
                CK(kl_k1_group_cold(l, d_sigs, d_msgs, n, d_out, d_status, d.k1g));
            else
                CK(kl_k1_group(l, d_sigs, d_msgs, n, d_out, d_status, d.k1g));
            break;
        case OP_R1: CK(kl_r1_group(l, d_sigs, d_msgs, n, d_out, d_status, d.r1g)); break;
        case OP_ED:
            if (cold)
                CK(kl_ed_group_cold(l, d_sigs, d_msgs, d_pks, n, d_out, d.edb));
            else
                CK(kl_ed_group(l, d_sigs, d_msgs, d_pks, n, d_out, d.edb));
            break;
    }
    if (!t_capturing) g_launches++;
    return 0;
}

// The kernels index their per-thread scratch by global thread id: `scratch` must hold chunks x grid x tpb Q4 and must not
// be shared by two launches that may run concurrently (launch_op orders the users of the per-device scratch with an event;
// queue slots own theirs).
int launch_op_scratch(Device& d, Op op, const uint8_t* d_sigs, const uint8_t* d_msgs, const uint8_t* d_pks, size_t n,
                      uint8_t* d_out, uint8_t* d_status, Q4* scratch, cudaStream_t st) {
    if (n == 0) return 0;
    if (use_group_kernel(d, op, n)) return launch_group(d, op, d_sigs, d_msgs, d_pks, n, d_out, d_status, st);
    const size_t main_n = tail_split(d, n);
    if (main_n < n) {
        const size_t os = op == OP_ED ? 1 : 64;
        if (launch_op_scratch(d, op, d_sigs, d_msgs, d_pks, main_n, d_out, d_status, scratch, st)) return 1;
        return launch_op_scratch(d, op, d_sigs + main_n * 64, d_msgs + main_n * 32, d_pks ? d_pks + main_n * 32 : nullptr,
                                 n - main_n, d_out + main_n * os, d_status ? d_status + main_n : nullptr, scratch, st);
    }
    KLaunch l;
    l.stream = st;
    launch_geometry(d, op, n, &l.grid, &l.tpb);
    switch (op) {
        case OP_K1: CK(kl_k1_ecrecover(l, d_sigs, d_msgs, n, d_out, d_status, scratch, d.k1g)); break;
        case OP_R1: CK(kl_r1_ecrecover(l, d_sigs, d_msgs, n, d_out, d_status, scratch, d.r1g)); break;
        case OP_ED: CK(kl_ed_verify(l, d_sigs, d_msgs, d_pks, n, d_out, scratch, d.edb)); break;
    }
    if (!t_capturing) g_launches++;
    return 0;
}

// launch on the per-device scratch: every such launch waits for the previous one (whatever stream that was enqueued on),
// so a *_device call on a caller's stream can never overlap a host-buffer call or another *_device call on the same tables
int launch_op(Device& d, Op op, const uint8_t* d_sigs, const uint8_t* d_msgs, const uint8_t* d_pks, size_t n,
              uint8_t* d_out, uint8_t* d_status, cudaStream_t st) {
    if (n == 0) return 0;
    const int max_g = op == OP_K1 ? d.grid_k1 : op == OP_R1 ? d.grid_r1 : d.grid_ed;
    const size_t chunks = op == OP_ED ? kEdBatchChunks : kSwBatchChunks;
    if (ensure_scratch(d, chunks * (size_t)max_g * kBlock)) return 1;
    if (d.scratch_used) CK(cudaStreamWaitEvent(st, d.scratch_ev, 0));
    if (launch_op_scratch(d, op, d_sigs, d_msgs, d_pks, n, d_out, d_status, d.scratch, st)) return 1;
    CK(cudaEventRecord(d.scratch_ev, st));
    d.scratch_used = true;
    return 0;
}

// One shard on one device.  The shard is cut into up to kMaxChunks pieces that flow through a three-stage pipeline
// (H2D on s_in, kernel on stream, D2H on s_out, chained with events), so that only the first piece's upload and the last
// piece's download are exposed.  Kernels run back to back on ONE stream: they share the per-thread scratch tables.
constexpr size_t kLeadWaves = 2;  // passes in the first, short piece (~4 ms of kernel): exposed upload ~0.3 ms
constexpr size_t kTailWaves = 2;  // and in the last one: exposed download ~0.2 ms instead of ~0.6 ms

// Piece plan: pieces are whole multiples of one wave (SMs x 512 threads, one signature per thread per pass) so that only
// the last piece ends on a partial wave.  The first and the last piece are short so that the kernels start early and the
// final download is small; the rest is cut into equal pieces of at most kSwBatch waves -- one full shared-inversion batch
// per thread (curve_sw.cuh).
int plan_pieces(const Device& d, size_t n, size_t* bounds) {
    int max_chunks = std::max(1, std::min(kMaxChunks, env_int("SIGOPS_MAX_CHUNKS", kMaxChunks)));
    const size_t wave = (size_t)d.sms * kBlock;
    const size_t n_waves = (n + wave - 1) / wave;
    int chunks = 1;
    bounds[0] = 0;
    if (max_chunks > 1 && n_waves >= 2 * kLeadWaves + 1) {
        size_t rest = n_waves - kLeadWaves;
        const bool tail = max_chunks > 2 && rest >= kTailWaves + 2;
        if (tail) rest -= kTailWaves;
        const int rest_chunks =
            (int)std::min<size_t>((size_t)max_chunks - 1 - (tail ? 1 : 0), (rest + kSwBatch - 1) / kSwBatch);
        bounds[1] = kLeadWaves * wave;
        for (int c = 1; c <= rest_chunks; c++) bounds[1 + c] = std::min(n, (kLeadWaves + rest * (size_t)c / rest_chunks) * wave);
        chunks = 1 + rest_chunks + (tail ? 1 : 0);
    } else if (max_chunks > 1 && n > wave) {
        // one to four waves (the 1M block cut over 4 or 8 GPUs): two pieces of whole waves, the first one the smaller --
        // the second upload and the first download overlap a kernel instead of being exposed
        chunks = 2;
        bounds[1] = std::max<size_t>(1, n_waves / 2) * wave;
    }
    bounds[chunks] = n;
    return chunks;
}

// The three-stage pipeline over the pieces of one shard.  up(c, lo, m, stream) enqueues the uploads of piece c = [lo, lo + m),
// launch(c, lo, m, stream) its kernels, down(c, lo, m, stream) its downloads; post(c, lo, m) runs on the host once piece c's
// downloads have completed (staged path: copy out of the pinned staging).  Each returns nonzero on failure.
template <class Up, class Launch, class Down, class Post>
int run_pipeline_body(Device& d, size_t n, bool inject_fail, bool has_post, Up up, Launch launch, Down down, Post post) {
    size_t bounds[kMaxChunks + 1];
    const int chunks = plan_pieces(d, n, bounds);
    CK(cudaEventRecord(d.ev[0], d.s_in));
    for (int c = 0; c < chunks; c++) {
        const size_t lo = bounds[c], m = bounds[c + 1] - lo;
        if (up(c, lo, m, d.s_in)) return 1;
        CK(cudaEventRecord(d.ev_in[c], d.s_in));
        CK(cudaStreamWaitEvent(d.stream, d.ev_in[c], 0));
        if (c == 0) CK(cudaEventRecord(d.ev[1], d.stream));
        if (launch(c, lo, m, d.stream)) return 1;
        CK(cudaEventRecord(d.ev_k[c], d.stream));
        if (inject_fail && c == 0) {  // fault injection: fail with work in flight on all three streams
            char b[128];
            snprintf(b, sizeof b, "injected failure on pool device %d (SIGOPS_FAIL_DEVICE)", d.index);
            set_err(b);
            return 1;
        }
    }
    CK(cudaEventRecord(d.ev[2], d.stream));
    // downloads are enqueued after every upload and kernel: a download into pageable memory blocks the calling thread
    // until its kernel has finished, which would otherwise stall the staging of the next piece's upload
    for (int c = 0; c < chunks; c++) {
        const size_t lo = bounds[c], m = bounds[c + 1] - lo;
        CK(cudaStreamWaitEvent(d.s_out, d.ev_k[c], 0));
        if (down(c, lo, m, d.s_out)) return 1;
        if (has_post) CK(cudaEventRecord(d.ev_down[c], d.s_out));
    }
    CK(cudaEventRecord(d.ev[3], d.s_out));
    if (has_post)
        for (int c = 0; c < chunks; c++) {
            CK(cudaEventSynchronize(d.ev_down[c]));
            if (post(c, bounds[c], bounds[c + 1] - bounds[c])) return 1;
        }
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

// Error discipline: whatever fails inside, nothing enqueued by this call may still read or write the caller's buffers (or
// the device's) once the API has returned -- drain the three streams and clear the sticky error state first.
template <class Up, class Launch, class Down, class Post>
int run_pipeline(Device& d, size_t n, bool inject_fail, bool has_post, Up up, Launch launch, Down down, Post post) {
    const int rc = run_pipeline_body(d, n, inject_fail, has_post, up, launch, down, post);
    if (rc) {
        cudaStreamSynchronize(d.s_in);
        cudaStreamSynchronize(d.stream);
        cudaStreamSynchronize(d.s_out);
        cudaGetLastError();
    }
    return rc;
}

template <class Up, class Launch, class Down>
int run_pipeline(Device& d, size_t n, bool inject_fail, Up up, Launch launch, Down down) {
    return run_pipeline(d, n, inject_fail, false, up, launch, down, [](int, size_t, size_t) { return 0; });
}

// is this host pointer ordinary pageable memory (neither cudaHostAlloc'ed nor cudaHostRegister'ed)?
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Small requests (the sizes the lane-group kernels serve): one stream, no events between streams.  When every buffer came
// from sigops_host_alloc the kernel reads the rows from and writes the results to the caller's pinned memory directly
// (zero copy: ~100 B per signature each way over PCIe inside a ~0.45 ms kernel) -- the five small copies and the stream hops
// of the pipelined path cost ~0.06 ms of a 0.49 ms request.  SIGOPS_ZERO_COPY=0 keeps the copies.
int run_small(Device& d, Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* out,
              uint8_t* status, bool inject_fail) {
    CK(cudaSetDevice(d.id));
    const size_t out_stride = op == OP_ED ? 1 : 64;
    const bool aligned = (((uintptr_t)sigs | (uintptr_t)msgs | (uintptr_t)pks | (uintptr_t)out) & 15u) == 0;
    // measured (profiles/r02_latency_sweep.json): zero copy wins while the request fits one block per SM (4,736 signatures:
    // 0.50 against 0.54 ms); at two blocks per SM the 16-byte host accesses of twice as many warps queue up on PCIe (secp256k1
    // 0.87 against 0.64 ms), so above that the rows are copied
    const bool zc = aligned && n <= (size_t)d.sms * kGroupSigs && env_int("SIGOPS_ZERO_COPY", 1) != 0 && pinned_ours(sigs, n * 64) && pinned_ours(msgs, n * 32) &&
                    (op != OP_ED || pinned_ours(pks, n * 32)) && pinned_ours(out, n * out_stride) &&
                    (op == OP_ED || !status || pinned_ours(status, n));
    int rc = 0;
    do {
        rc = 1;
        if (zc) {
            if (cudaEventRecord(d.ev[1], d.stream) != cudaSuccess) break;
            if (launch_op(d, op, sigs, msgs, pks, n, out, op == OP_ED ? nullptr : status, d.stream)) return 1;
            if (cudaEventRecord(d.ev[2], d.stream) != cudaSuccess) break;
        } else {
            const size_t in_bytes = n * (op == OP_ED ? 128 : 96), out_bytes = op == OP_ED ? n : n * 65;
            if (ensure_buf(&d.d_in, &d.in_cap, in_bytes + 64)) return 1;
            if (ensure_buf(&d.d_out, &d.out_cap, out_bytes + 64)) return 1;
            uint8_t *d_sigs = d.d_in, *d_msgs = d.d_in + n * 64, *d_pks = d.d_in + n * 96;
            uint8_t* d_status = op == OP_ED ? nullptr : d.d_out + n * 64;
            if (cudaMemcpyAsync(d_sigs, sigs, n * 64, cudaMemcpyHostToDevice, d.stream) != cudaSuccess) break;
            if (cudaMemcpyAsync(d_msgs, msgs, n * 32, cudaMemcpyHostToDevice, d.stream) != cudaSuccess) break;
            if (op == OP_ED && cudaMemcpyAsync(d_pks, pks, n * 32, cudaMemcpyHostToDevice, d.stream) != cudaSuccess) break;
            if (cudaEventRecord(d.ev[1], d.stream) != cudaSuccess) break;
            if (launch_op(d, op, d_sigs, d_msgs, d_pks, n, d.d_out, d_status, d.stream)) return 1;
            if (cudaEventRecord(d.ev[2], d.stream) != cudaSuccess) break;
            if (cudaMemcpyAsync(out, d.d_out, n * out_stride, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) break;
            if (d_status && status && cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) break;
        }
        rc = 0;
    } while (0);
    const bool injected = !rc && inject_fail;  // fault injection: fail with the request in flight
    const cudaError_t e = cudaStreamSynchronize(d.stream);  // nothing stays in flight, whatever happened
    if (rc || injected || e != cudaSuccess) {
        if (injected) {
            char b[128];
            snprintf(b, sizeof b, "injected failure on pool device %d (SIGOPS_FAIL_DEVICE)", d.index);
            set_err(b);
        } else {
            set_err(std::string("small-request path failed: ") + cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
        }
        cudaGetLastError();
        return 1;
    }
    d.ms_h2d = 0;
    d.ms_d2h = 0;
    CK(cudaEventElapsedTime(&d.ms_kernel, d.ev[1], d.ev[2]));
    return 0;
}

// Below this many signatures a pageable shard is copied directly (the driver stages small copies itself at no cost)
constexpr size_t kMinStaged = 16384;

int run_shard(Device& d, Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* out,
              uint8_t* status, bool inject_fail) {
    if (use_group_kernel(d, op, n) && env_int("SIGOPS_SMALL_PATH", 1) != 0)
        return run_small(d, op, sigs, msgs, pks, n, out, status, inject_fail);
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
    // Pageable caller buffers (what `&Vec<...>` hands over, src/secp256k1_ecdsa.rs:61-66) go through pinned staging filled
    // / drained by the device's copy threads, piece by piece, so that the DMA engines run at pinned speed and the host
    // copies overlap the kernels; replaces the create_buffer_init upload / MAP_READ staging of src/gpu.rs:41-49,138-166.
    const int staging = env_int("SIGOPS_STAGING", -1);  // -1 auto, 0 never, 1 always
    const bool big = n >= kMinStaged || staging == 1;
    const bool stage_in = staging != 0 && big && (staging == 1 || is_pageable(sigs) || is_pageable(msgs) || (pks && is_pageable(pks)));
    const bool stage_out = staging != 0 && big && (staging == 1 || is_pageable(out) || (status && is_pageable(status)));
    if (stage_in && ensure_pinned(&d.h_in, &d.h_in_cap, in_bytes)) return 1;
    if (stage_out && ensure_pinned(&d.h_out, &d.h_out_cap, out_bytes)) return 1;
    uint8_t *h_sigs = d.h_in, *h_msgs = d.h_in + n * 64, *h_pks = d.h_in + n * 96;
    uint8_t *h_o = d.h_out, *h_st = d.h_out + n * 64;
    Counter in_done[kMaxChunks], out_done;
    struct WaitAll {  // no copy job may outlive this frame (they reference the caller's buffers and the counters)
        Counter *a, *b;
        ~WaitAll() {
            for (int i = 0; i < kMaxChunks; i++) a[i].wait();
            b->wait();
        }
    } wait_all{in_done, &out_done};
    if (stage_in) {
        size_t bounds[kMaxChunks + 1];
        const int chunks = plan_pieces(d, n, bounds);
        for (int c = 0; c < chunks; c++) {  // FIFO: piece 0 is staged first
            const size_t lo = bounds[c], m = bounds[c + 1] - lo;
            d.post_copy(h_sigs + lo * 64, sigs + lo * 64, m * 64, &in_done[c]);
            d.post_copy(h_msgs + lo * 32, msgs + lo * 32, m * 32, &in_done[c]);
            if (op == OP_ED) d.post_copy(h_pks + lo * 32, pks + lo * 32, m * 32, &in_done[c]);
        }
    }
    auto up = [&](int c, size_t lo, size_t m, cudaStream_t st) -> int {
        const uint8_t *s = sigs, *g = msgs, *k = pks;
        if (stage_in) {
            in_done[c].wait();
            s = h_sigs;
            g = h_msgs;
            k = h_pks;
        }
        CK(cudaMemcpyAsync(d_sigs + lo * 64, s + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_msgs + lo * 32, g + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        if (op == OP_ED) CK(cudaMemcpyAsync(d_pks + lo * 32, k + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto launch = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        return launch_op(d, op, d_sigs + lo * 64, d_msgs + lo * 32, d_pks + lo * 32, m, d.d_out + lo * out_stride,
                         d_status ? d_status + lo : nullptr, st);
    };
    auto down = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        uint8_t* o = stage_out ? h_o : out;
        uint8_t* s = stage_out ? h_st : status;
        CK(cudaMemcpyAsync(o + lo * out_stride, d.d_out + lo * out_stride, m * out_stride, cudaMemcpyDeviceToHost, st));
        if (op != OP_ED && status) CK(cudaMemcpyAsync(s + lo, d_status + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    auto post = [&](int, size_t lo, size_t m) -> int {
        d.post_copy(out + lo * out_stride, h_o + lo * out_stride, m * out_stride, &out_done);
        if (op != OP_ED && status) d.post_copy(status + lo, h_st + lo, m, &out_done);
        return 0;
    };
    const int rc = run_pipeline(d, n, inject_fail, stage_out, up, launch, down, post);
    out_done.wait();
    return rc;
}

// A device's shard is processed in sub-shards of at most kMaxSubShard signatures so that the device buffers stay
// bounded (2 GiB in, 1 GiB out; the pinned staging of the pageable path is bounded the same way) however large the batch
// (the reference allows up to 2^30 signatures per call, src/secp256k1_ecdsa.rs:22).  Timings accumulate over the sub-shards.
constexpr size_t kMaxSubShard = (size_t)1 << 24;
constexpr size_t kMaxSubShardStaged = (size_t)1 << 22;  // 512 MiB + 260 MiB of pinned staging per device at most

int run_shard_bounded(Device& d, Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n,
                      uint8_t* out, uint8_t* status, bool inject_fail) {
    const size_t out_stride = op == OP_ED ? 1 : 64;
    float h2d = 0, ker = 0, d2h = 0;
    size_t sub = kMaxSubShard;
    if (n > kMaxSubShardStaged && env_int("SIGOPS_STAGING", -1) != 0 && (is_pageable(sigs) || is_pageable(out)))
        sub = kMaxSubShardStaged;
    if (const char* e = getenv("SIGOPS_MAX_SUBSHARD")) sub = std::max<size_t>(1, (size_t)atoll(e));  // test hook
    for (size_t lo = 0; lo < n; lo += sub) {
        const size_t m = std::min(sub, n - lo);
        if (int rc = run_shard(d, op, sigs + lo * 64, msgs + lo * 32, pks ? pks + lo * 32 : nullptr, m,
                               out + lo * out_stride, status ? status + lo : nullptr, inject_fail))
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

// Runs fn(device, g, lo, hi) for every shard of an n-item batch: shard 0 on the calling thread, the others on their devices'
// persistent workers; every shard holds its device's mutex while it runs.  `cand` (pool indices, may be empty = all) limits
// the devices; when fewer devices than candidates are needed the idle ones are preferred.  One failed shard fails the whole
// call (all-or-nothing, like ShaderFailureError).  Caller holds the pool lock (shared).
template <class Fn>
int for_each_shard(size_t n, const std::vector<int>& cand_in, Fn fn) {
    auto& P = pool();
    std::vector<int> cand = cand_in;
    if (cand.empty())
        for (size_t i = 0; i < P.size(); i++) cand.push_back((int)i);
    std::vector<size_t> bounds(cand.size() + 1);
    const size_t G = (size_t)plan_shards(n, (int)cand.size(), bounds.data());
    std::vector<int> use;
    if (G < cand.size()) {
        for (int i : cand)
            if (use.size() < G && P[i]->inflight.load() == 0) use.push_back(i);
        for (int i : cand)
            if (use.size() < G && std::find(use.begin(), use.end(), i) == use.end()) use.push_back(i);
        std::sort(use.begin(), use.end());
    } else {
        use = cand;
    }
    const uint64_t call = ++g_call_seq;
    t_last_call = call;
    const int fail_dev = env_int("SIGOPS_FAIL_DEVICE", -1);
    std::vector<int> rcs(G, 0);
    Counter done;
    auto shard = [&](size_t g) {
        Device& d = *P[use[g]];
        d.inflight++;
        {
            std::lock_guard<std::mutex> lk(d.mu);
            d.ms_h2d = d.ms_kernel = d.ms_d2h = 0;
            d.last_call = call;
            rcs[g] = fn(d, bounds[g], bounds[g + 1], d.index == fail_dev);
        }
        d.inflight--;
    };
    if (G > 1) {
        done.add(G - 1);
        for (size_t g = 1; g < G; g++)
            P[use[g]]->post([&, g] {
                shard(g);
                done.done();
            });
    }
    shard(0);
    done.wait();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return rcs[g];
    return 0;
}

int run_batch(Op op, const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* out,
              uint8_t* status, const int* dev_idx = nullptr, int n_idx = 0) {
    if (n == 0) return 0;
    if (!sigs || !msgs || !out || (op == OP_ED && !pks)) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    if (!g_inited.load()) {
        set_err("the device pool was shut down during the call");
        return 1;
    }
    std::vector<int> cand;
    for (int i = 0; i < n_idx; i++) {
        if (dev_idx[i] < 0 || dev_idx[i] >= (int)pool().size() || std::find(cand.begin(), cand.end(), dev_idx[i]) != cand.end()) {
            set_err("device index out of range or repeated (indices refer to the pool, see sigops_num_devices)");
            return 1;
        }
        cand.push_back(dev_idx[i]);
    }
    const size_t out_stride = op == OP_ED ? 1 : 64;
    const int rc = for_each_shard(n, cand, [&](Device& d, size_t lo, size_t hi, bool inject_fail) {
        return run_shard_bounded(d, op, sigs + lo * 64, msgs + lo * 32, pks ? pks + lo * 32 : nullptr, hi - lo,
                                 out + lo * out_stride, status ? status + lo : nullptr, inject_fail);
    });
    if (rc) {  // all-or-nothing: no partial results are left behind
        memset(out, 0, n * out_stride);
        if (status) memset(status, op == OP_ED ? 0 : SIGOPS_STATUS_INVALID, n);
    }
    return rc;
}

int run_device(Op op, const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n, void* d_out, void* d_status,
               void* stream) {
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    Device* d = current_device();
    if (!d) {
        set_err("current CUDA device is not part of the sigops pool");
        return 1;
    }
    std::lock_guard<std::mutex> dl(d->mu);
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
// A queue is a ring of `depth` slots for ONE operation, on one device of the pool or spread round-robin over all of them
// (device_index = -1).  Every slot owns pinned host staging (inputs and outputs), device buffers, its own per-thread
// scratch and its own stream, so the slots of a queue run concurrently on the device: a request of a few thousand
// signatures occupies one or two warps per SM (launch_geometry), and up to 16 warps per SM are resident, so `depth` small
// requests in flight multiply the throughput at the latency of one.  A slot's H2D copies, kernel and D2H copies are
// replayed as one CUDA graph while the request size repeats (re-captured when it changes).
// Replaces the per-call device creation / buffer allocation / blocking poll of src/gpu.rs:5-35,129-170.
struct QSlot {
    Device* dev = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr;  // pinned: sigs (cap*64) | msgs (cap*32) | pks (cap*32); out (cap*64) | status (cap)
    uint8_t *d_in = nullptr, *d_out = nullptr;
    Q4* scratch = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t exec_n = 0;
    bool busy = false;
    bool waiting = false;  // a thread is blocked in sigops_queue_wait on this slot
    size_t n = 0;
    float last_ms = 0;
};

}  // namespace

struct sigops_queue {
    std::mutex mu;
    int device_index = 0;  // -1: slots spread over the pool
    Op op = OP_K1;
    size_t cap = 0;
    bool graphs = true;
    std::vector<QSlot> slots;
    uint64_t graph_launches = 0, graph_captures = 0;
};

namespace {

std::atomic<int> g_live_queues{0};

inline size_t q_out_stride(Op op) { return op == OP_ED ? 1 : 64; }

// enqueue the kernel and the downloads of one request of n signatures of slot s on stream st
int queue_enqueue_tail(sigops_queue* q, QSlot& s, size_t n, cudaStream_t st) {
    const size_t cap = q->cap;
    const Op op = q->op;
    uint8_t* d_status = op == OP_ED ? nullptr : s.d_out + cap * 64;
    if (launch_op_scratch(*s.dev, op, s.d_in, s.d_in + cap * 64, s.d_in + cap * 96, n, s.d_out, d_status, s.scratch, st))
        return 1;
    CK(cudaMemcpyAsync(s.h_out, s.d_out, n * q_out_stride(op), cudaMemcpyDeviceToHost, st));
    if (d_status) CK(cudaMemcpyAsync(s.h_out + cap * 64, d_status, n, cudaMemcpyDeviceToHost, st));
    return 0;
}

// the whole request: uploads from the slot's pinned arrays, kernel, downloads (also used under stream capture: no
// allocation, no synchronisation)
int queue_enqueue(sigops_queue* q, QSlot& s, size_t n, cudaStream_t st) {
    const size_t cap = q->cap;
    if (use_group_kernel(*s.dev, q->op, n) && n <= (size_t)s.dev->sms * kGroupSigs && env_int("SIGOPS_ZERO_COPY", 1) != 0) {
        // small request: the lane-group kernel works on the slot's pinned arrays directly (zero copy)
        uint8_t* h_status = q->op == OP_ED ? nullptr : s.h_out + cap * 64;
        return launch_op_scratch(*s.dev, q->op, s.h_in, s.h_in + cap * 64, s.h_in + cap * 96, n, s.h_out, h_status, s.scratch, st);
    }
    CK(cudaMemcpyAsync(s.d_in, s.h_in, n * 64, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_in + cap * 64, s.h_in + cap * 64, n * 32, cudaMemcpyHostToDevice, st));
    if (q->op == OP_ED) CK(cudaMemcpyAsync(s.d_in + cap * 96, s.h_in + cap * 96, n * 32, cudaMemcpyHostToDevice, st));
    return queue_enqueue_tail(q, s, n, st);
}

void queue_free_slot(QSlot& s) {
    if (s.dev) cudaSetDevice(s.dev->id);
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
    const Device& d = *s.dev;
    const size_t cap = q->cap;
    const int max_g = q->op == OP_K1 ? d.grid_k1 : q->op == OP_R1 ? d.grid_r1 : d.grid_ed;
    const size_t chunks = q->op == OP_ED ? kEdBatchChunks : kSwBatchChunks;
    // launch_geometry never uses more than min(max_g * 512, n + 512) threads for n signatures
    const size_t threads = std::min((size_t)max_g * kBlock, cap + (size_t)kBlock);
    CK(cudaSetDevice(d.id));
    CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s.t0));
    CK(cudaEventCreate(&s.t1));
    CK(cudaHostAlloc((void**)&s.h_in, cap * 128, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc((void**)&s.h_out, cap * 65, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaMalloc((void**)&s.d_in, cap * 128 + 64));
    CK(cudaMalloc((void**)&s.d_out, cap * 65 + 64));
    CK(cudaMalloc((void**)&s.scratch, chunks * threads * sizeof(Q4)));
    return 0;
}

// shared head of the two submit calls: slot checks; returns the slot or nullptr
QSlot* queue_submit_check(sigops_queue* q, int slot, size_t n, const char* who) {
    if (!q || slot < 0 || slot >= (int)q->slots.size()) {
        set_err(std::string(who) + ": bad queue or slot");
        return nullptr;
    }
    QSlot& s = q->slots[slot];
    if (s.busy) {
        set_err(std::string(who) + ": slot is in flight (wait for it first)");
        return nullptr;
    }
    if (n > q->cap) {
        set_err(std::string(who) + ": n exceeds the queue's max_batch");
        return nullptr;
    }
    return &s;
}

}  // namespace

extern "C" {

int sigops_init(const int* device_ids, int n_devices) {
    std::unique_lock<std::shared_mutex> lk(g_pool_mu);
    return do_init_locked(device_ids, n_devices);
}

int sigops_shutdown(void) {
    std::unique_lock<std::shared_mutex> lk(g_pool_mu);
    if (g_live_queues.load() > 0) {
        set_err("sigops_shutdown: destroy every sigops_queue first");
        return 1;
    }
    for (auto& d : pool()) free_device(*d);
    pool().clear();
    g_ids.clear();
    g_inited.store(false);
    return 0;
}

int sigops_num_devices(void) {
    if (ensure_init()) return 0;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    return (int)pool().size();
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

int sigops_batch_on_devices(int curve, const int* device_indices, int n_devices, const uint8_t* sigs, const uint8_t* msgs,
                            const uint8_t* pks, size_t n, uint8_t* out, uint8_t* out_status) {
    if (curve < SIGOPS_CURVE_SECP256K1 || curve > SIGOPS_CURVE_ED25519 || n_devices < 0 || (n_devices > 0 && !device_indices)) {
        set_err("sigops_batch_on_devices: bad arguments");
        return 1;
    }
    return run_batch((Op)curve, sigs, msgs, pks, n, out, curve == SIGOPS_CURVE_ED25519 ? nullptr : out_status, device_indices,
                     n_devices);
}

// Variable-length-message ed25519: one device shard = signatures [lo, hi) and the message bytes they span, through the same
// piecewise upload / kernel / download pipeline as the fixed-size entry points.
static int run_ed_msgs_shard(Device& d, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, const uint8_t* pks,
                             size_t n, uint32_t flags, uint8_t* out, bool inject_fail) {
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
    auto up = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_pks + lo * 32, pks + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        const uint64_t bl = rebased[lo], bh = rebased[lo + m];
        if (bh > bl) CK(cudaMemcpyAsync(d_msg + bl, msg_bytes + b0 + bl, (size_t)(bh - bl), cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto launch = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        KLaunch l;
        l.stream = st;
        l.tpb = kBlock;
        if (m < (size_t)d.sms * kBlock) {
            size_t per_sm = (m + d.sms - 1) / d.sms;
            l.tpb = (int)std::min<size_t>(kBlock, std::max<size_t>(32, (per_sm + 31) / 32 * 32));
        }
        l.grid = (int)std::min<size_t>((m + l.tpb - 1) / l.tpb, (size_t)d.grid_edm);
        if (d.scratch_used) CK(cudaStreamWaitEvent(st, d.scratch_ev, 0));
        CK(kl_ed_verify_msgs(l, d_sigs + lo * 64, d_msg, (const unsigned long long*)d_off + lo, d_pks + lo * 32, m,
                             (int)(flags & SIGOPS_ED25519_STRICT), d.d_out + lo, d.scratch, d.edb));
        CK(cudaEventRecord(d.scratch_ev, st));
        d.scratch_used = true;
        g_launches++;
        return 0;
    };
    auto down = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(out + lo, d.d_out + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    return run_pipeline(d, n, inject_fail, up, launch, down);
}

int sigops_ed25519_ecverify_msgs(const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                                 const uint8_t* pks, size_t n, uint32_t flags, uint8_t* out_valid) {
    if (n == 0) return 0;
    if (!sigs || !msg_offsets || !pks || !out_valid || (!msg_bytes && msg_offsets[n] != msg_offsets[0])) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    const int rc = for_each_shard(n, {}, [&](Device& d, size_t lo, size_t hi, bool inject_fail) {
        // sub-shards bound the device buffers as in run_shard_bounded
        for (size_t a = lo; a < hi; a += kMaxSubShard) {
            const size_t m = std::min(kMaxSubShard, hi - a);
            if (int r = run_ed_msgs_shard(d, sigs + a * 64, msg_bytes, msg_offsets + a, pks + a * 32, m, flags, out_valid + a,
                                          inject_fail))
                return r;
        }
        return 0;
    });
    if (rc) memset(out_valid, 0, n);
    return rc;
}

// raw messages -> SHA-256 -> recover -> SHA-256(pubkey), all on the device, piece by piece; no host pass in between
static int run_addresses_shard(Device& d, Op op, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, size_t n,
                               uint8_t* out_addr, uint8_t* out_pk, uint8_t* out_st, bool inject_fail) {
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
    auto up = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(d_sigs + lo * 64, sigs + lo * 64, m * 64, cudaMemcpyHostToDevice, st));
        if (off) {
            const uint64_t bl = rebased[lo], bh = rebased[lo + m];
            if (bh > bl) CK(cudaMemcpyAsync(d_raw + bl, msg_bytes + b0 + bl, (size_t)(bh - bl), cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMemcpyAsync(d_msgs + lo * 32, msg_bytes + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
        }
        return 0;
    };
    auto launch = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        if (off) {
            CK(kl_sha256_msgs(st, d_raw, (const unsigned long long*)d_off + lo, m, (u32*)(d_msgs + lo * 32)));
            g_launches++;
        }
        if (launch_op(d, op, d_sigs + lo * 64, d_msgs + lo * 32, nullptr, m, d.d_out + lo * 64, d_status + lo, st)) return 1;
        CK(kl_sha256_pubkeys(st, (const u32*)(d.d_out + lo * 64), d_status + lo, m, (u32*)(d_addr + lo * 32)));
        g_launches++;
        return 0;
    };
    auto down = [&](int, size_t lo, size_t m, cudaStream_t st) -> int {
        CK(cudaMemcpyAsync(out_addr + lo * 32, d_addr + lo * 32, m * 32, cudaMemcpyDeviceToHost, st));
        if (out_pk) CK(cudaMemcpyAsync(out_pk + lo * 64, d.d_out + lo * 64, m * 64, cudaMemcpyDeviceToHost, st));
        if (out_st) CK(cudaMemcpyAsync(out_st + lo, d_status + lo, m, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    return run_pipeline(d, n, inject_fail, up, launch, down);
}

int sigops_ecrecover_addresses(int curve, const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* msg_offsets,
                               size_t n, uint8_t* out_addresses, uint8_t* out_pubkeys, uint8_t* out_status) {
    if (n == 0) return 0;
    if (curve != SIGOPS_CURVE_SECP256K1 && curve != SIGOPS_CURVE_SECP256R1) {
        set_err("sigops_ecrecover_addresses: curve must be secp256k1 or secp256r1");
        return 1;
    }
    if (!sigs || !out_addresses || (!msg_bytes && (!msg_offsets || msg_offsets[n] != msg_offsets[0]))) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    const Op op = curve == SIGOPS_CURVE_SECP256K1 ? OP_K1 : OP_R1;
    const int rc = for_each_shard(n, {}, [&](Device& d, size_t lo, size_t hi, bool inject_fail) {
        for (size_t a = lo; a < hi; a += kMaxSubShard) {
            const size_t m = std::min(kMaxSubShard, hi - a);
            if (int r = run_addresses_shard(d, op, sigs + a * 64, msg_offsets ? msg_bytes : msg_bytes + a * 32,
                                            msg_offsets ? msg_offsets + a : nullptr, m, out_addresses + a * 32,
                                            out_pubkeys ? out_pubkeys + a * 64 : nullptr, out_status ? out_status + a : nullptr,
                                            inject_fail))
                return r;
        }
        return 0;
    });
    if (rc) {
        memset(out_addresses, 0, n * 32);
        if (out_pubkeys) memset(out_pubkeys, 0, n * 64);
        if (out_status) memset(out_status, SIGOPS_STATUS_INVALID, n);
    }
    return rc;
}

int sigops_sha256_batch(const uint8_t* data, const uint64_t* offsets, size_t n, uint8_t* out) {
    if (n == 0) return 0;
    if (!offsets || !out || (!data && offsets[n] != offsets[0])) {
        set_err("null buffer");
        return 1;
    }
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    Device& d = *pool()[0];
    std::lock_guard<std::mutex> dl(d.mu);
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
        int rc = 0;
        do {  // one pass; any failure drains the stream before `rebased` goes out of scope
            rc = 1;
            if (cudaMemcpyAsync(d.d_in, rebased.data(), off_bytes, cudaMemcpyHostToDevice, d.stream) != cudaSuccess) break;
            if (nbytes && cudaMemcpyAsync(d.d_in + off_bytes, data + b0, (size_t)nbytes, cudaMemcpyHostToDevice, d.stream) != cudaSuccess) break;
            if (kl_sha256_msgs(d.stream, d.d_in + off_bytes, (const unsigned long long*)d.d_in, m, (u32*)d.d_out)) break;
            g_launches++;
            if (cudaMemcpyAsync(out + a * 32, d.d_out, m * 32, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) break;
            rc = 0;
        } while (0);
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (rc || e != cudaSuccess) {
            set_err(std::string("sigops_sha256_batch failed: ") + cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
            cudaGetLastError();
            return 1;
        }
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
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    double a = 0, b = 0, c = 0;
    for (auto& dp : pool()) {
        Device& d = *dp;
        std::lock_guard<std::mutex> dl(d.mu);
        if (d.last_call != t_last_call) continue;  // not part of the calling thread's last host-buffer call
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
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        set_err("cudaHostAlloc failed");
        return nullptr;
    }
    pinned_add(p, bytes ? bytes : 1);
    return p;
}

void sigops_host_free(void* p) {
    if (!p) return;
    pinned_remove(p);
    cudaFreeHost(p);
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
    if (op < 0 || op >= SIGOPS_UNIT_COUNT) {
        set_err("bad unit op");
        return 1;
    }
    if (n == 0) return 0;
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    Device& d = *pool()[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CK(cudaSetDevice(d.id));
    int in_w, out_w;
    unit_shape(op, in_w, out_w);
    if (ensure_buf(&d.d_in, &d.in_cap, n * in_w * 4)) return 1;
    if (ensure_buf(&d.d_out, &d.out_cap, n * out_w * 4)) return 1;
    if (ensure_scratch(d, (size_t)kEdTabChunks * d.grid_unit * kBlock)) return 1;
    CK(cudaMemcpyAsync(d.d_in, in, n * in_w * 4, cudaMemcpyHostToDevice, d.stream));
    KLaunch l;
    l.stream = d.stream;
    l.tpb = kBlock;
    if (op >= SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL && op <= SIGOPS_UNIT_ED_GROUP_MULPT) {  // the lane-group twins: 32 items per block, no scratch
        l.grid = (int)std::min<size_t>((n + kGroupSigs - 1) / kGroupSigs, (size_t)d.sms * 2);
        CK(kl_unit_group(l, op, (const u32*)d.d_in, n, (u32*)d.d_out, d.k1g, d.r1g));
    } else {
        l.grid = (int)std::min<size_t>((n + kBlock - 1) / kBlock, (size_t)d.grid_unit);
        if (d.scratch_used) CK(cudaStreamWaitEvent(d.stream, d.scratch_ev, 0));
        CK(kl_unit(l, op, (const u32*)d.d_in, n, (u32*)d.d_out, d.scratch, d.k1g, d.r1g, d.edb));
        CK(cudaEventRecord(d.scratch_ev, d.stream));
        d.scratch_used = true;
    }
    g_launches++;
    CK(cudaMemcpyAsync(out, d.d_out, n * out_w * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return 0;
}

int sigops_imad_peak(int kind, int iters, double* ops_per_sec, double* ms_out) {
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    Device* dp = current_device();  // the pool device the caller is on, else the first one
    Device& d = dp ? *dp : *pool()[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CK(cudaSetDevice(d.id));
    if (ensure_buf(&d.d_out, &d.out_cap, (size_t)d.sms * 8 * 256 * 4)) return 1;
    const int grid = d.sms * 8, block = 256;  // 2048 threads per SM: full occupancy
    for (int rep = 0; rep < 2; rep++) {       // first pass warms up
        CK(cudaEventRecord(d.ev[0], d.stream));
        int rc = kl_imad_peak(kind, grid, block, d.stream, (u32*)d.d_out, iters, 12345u);
        if (rc == -1) {
            set_err("bad kind");
            return 1;
        }
        CK(rc);
        CK(cudaEventRecord(d.ev[1], d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
    // counted operations per thread per outer iteration (16 unrolled blocks):
    //  kind 0: 8 IMAD; 1: 8 IMAD.WIDE; 2: 8 IMAD.WIDE(.X); 3: 8 IADD; 4: 8 IMAD.WIDE + 8 IADD (counted: the 8 wide);
    //  5: 8 DFMA; 6: 8 IMAD.HI
    const double per_thread_iter = 16.0 * 8.0;
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
    if (int rc = ensure_init()) return rc;
    std::shared_lock<std::shared_mutex> lk(g_pool_mu);
    auto& P = pool();
    if (device_index < -1 || device_index >= (int)P.size()) {
        set_err("sigops_queue_create: device_index is not in the pool (-1 = spread the slots over all devices)");
        return 1;
    }
    sigops_queue* q = new sigops_queue();
    q->device_index = device_index;
    q->op = (Op)curve;
    q->cap = max_batch;
    if (const char* e = getenv("SIGOPS_QUEUE_GRAPHS")) q->graphs = atoi(e) != 0;
    q->slots.resize(depth);
    for (int i = 0; i < depth; i++) {
        QSlot& s = q->slots[i];
        s.dev = P[device_index >= 0 ? device_index : i % (int)P.size()].get();
        if (queue_alloc_slot(q, s)) {
            for (auto& t : q->slots) queue_free_slot(t);
            delete q;
            return 1;
        }
    }
    g_live_queues++;
    *out = q;
    return 0;
}

int sigops_queue_destroy(sigops_queue* q) {
    if (!q) return 0;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        for (auto& s : q->slots) {
            if (s.dev) cudaSetDevice(s.dev->id);
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
    if (!q) {
        set_err("sigops_queue_submit: bad queue or slot");
        return 1;
    }
    std::lock_guard<std::mutex> lk(q->mu);
    QSlot* sp = queue_submit_check(q, slot, n, "sigops_queue_submit");
    if (!sp) return 1;
    QSlot& s = *sp;
    s.n = n;
    s.last_ms = 0;
    if (n == 0) return 0;  // nothing to do: the slot stays free (src/secp256k1_ecdsa.rs:71-73)
    CK(cudaSetDevice(s.dev->id));
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
            cudaGetLastError();
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
        g_launches += (!use_group_kernel(*s.dev, q->op, n) && tail_split(*s.dev, n) < n) ? 2 : 1;  // twice when the tail is launched on its own
        q->graph_launches++;
    } else if (queue_enqueue(q, s, n, s.st)) {
        return 1;
    }
    CK(cudaEventRecord(s.t1, s.st));
    s.busy = true;
    return 0;
}

int sigops_queue_submit_device(sigops_queue* q, int slot, const void* d_sigs, const void* d_msgs, const void* d_pks, size_t n,
                               int src_device, void* ready_event) {
    if (!q) {
        set_err("sigops_queue_submit_device: bad queue or slot");
        return 1;
    }
    std::lock_guard<std::mutex> lk(q->mu);
    QSlot* sp = queue_submit_check(q, slot, n, "sigops_queue_submit_device");
    if (!sp) return 1;
    QSlot& s = *sp;
    if (!d_sigs || !d_msgs || (q->op == OP_ED && !d_pks)) {
        set_err("sigops_queue_submit_device: null buffer");
        return 1;
    }
    s.n = n;
    s.last_ms = 0;
    if (n == 0) return 0;
    const int dst = s.dev->id;
    CK(cudaSetDevice(dst));
    if (src_device != dst) {  // direct NVLink path when the devices are peers; the copy still works (staged) when they are not
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dst, src_device) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(src_device, 0);
            if (e != cudaSuccess) cudaGetLastError();  // already enabled
        }
    }
    const size_t cap = q->cap;
    if (ready_event) CK(cudaStreamWaitEvent(s.st, (cudaEvent_t)ready_event, 0));
    CK(cudaEventRecord(s.t0, s.st));
    int rc = 0;
    do {
        rc = 1;
        if (cudaMemcpyPeerAsync(s.d_in, dst, d_sigs, src_device, n * 64, s.st) != cudaSuccess) break;
        if (cudaMemcpyPeerAsync(s.d_in + cap * 64, dst, d_msgs, src_device, n * 32, s.st) != cudaSuccess) break;
        if (q->op == OP_ED && cudaMemcpyPeerAsync(s.d_in + cap * 96, dst, d_pks, src_device, n * 32, s.st) != cudaSuccess) break;
        rc = queue_enqueue_tail(q, s, n, s.st);
    } while (0);
    if (rc) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) set_err(std::string("sigops_queue_submit_device: ") + cudaGetErrorString(e));
        cudaStreamSynchronize(s.st);  // nothing of the failed request stays in flight
        cudaGetLastError();
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
        if (s.waiting) {  // one waiter per slot: a second one would race on the slot's state
            set_err("sigops_queue_wait: another thread is already waiting on this slot");
            return 1;
        }
        s.waiting = true;
        ev = s.t1;
    }
    const cudaError_t e = cudaEventSynchronize(ev);  // outside the queue lock: other threads keep submitting to other slots
    std::lock_guard<std::mutex> lk(q->mu);
    // whatever happened, the slot is released: a failed request must not wedge the queue
    s.waiting = false;
    s.busy = false;
    s.last_ms = 0;
    if (e != cudaSuccess) {
        set_err(std::string("sigops_queue_wait: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
    }
    if (cudaEventElapsedTime(&s.last_ms, s.t0, s.t1) != cudaSuccess) {
        cudaGetLastError();
        s.last_ms = 0;
    }
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
    if (device_index) *device_index = q->device_index;
    if (max_batch) *max_batch = q->cap;
    if (depth) *depth = (int)q->slots.size();
    if (graph_launches) *graph_launches = q->graph_launches;
    if (graph_captures) *graph_captures = q->graph_captures;
    return 0;
}

int sigops_queue_slot_device(sigops_queue* q, int slot) {
    if (!q || slot < 0 || slot >= (int)q->slots.size()) return -1;
    return q->slots[slot].dev->id;
}

}  // extern "C"
