// Host simulation of the device code (TEST INFRASTRUCTURE ONLY).
//
// Compiles the very same headers the CUDA kernels are built from with a plain C++ compiler; the inline-PTX bodies
// are replaced by their portable uint64_t twins (arith.cuh, SG_PTX == 0).  This lets the `-m "not gpu"` tests check
// all the host-visible logic -- reductions, addition chains, group law corner cases, recoding, GLV split, the
// fused per-signature functions -- against the oracle without a GPU.  It is never linked into libsigops.so.
#include <algorithm>
#include <cstring>
#include <vector>
#include <pthread.h>

#include <thread>

#include "../../wgpu-sigops_b200/csrc/kernels.cuh"
#include "../../wgpu-sigops_b200/csrc/group.cuh"

using namespace sigops;

// positional fixed-base tables (csrc/ptab.h), generated on the host with the same base / entry functions the device's
// gen_ptab kernels run.  The window is narrow here (HOSTSIM_GWIN, default 7 bits: 37 windows of 64 entries) so that the
// tables take a fraction of a second on one CPU core; the kernels take the width at run time, the GPU tests cover the
// product's width; hostsim_set_gwin rebuilds them at another width (tests of the window arithmetic).
static u32 g_gwin = 7;
static std::vector<u32> k1_gtab_w, r1_gtab_w, ed_btab_w;
static PTab k1_gtab, r1_gtab, ed_btab;
static void ensure_tables() {
    if (!k1_gtab_w.empty()) return;
    const u32 w = g_gwin, pos = ptab_positions(w);
    const size_t per = (size_t)1 << (w - 1);
    k1_gtab_w.resize(per * pos * 16);
    r1_gtab_w.resize(per * pos * 16);
    ed_btab_w.resize(per * pos * 24);
    for (u32 j = 0; j < pos; j++) {
        u32 bk[16], br[16], be[24];
        sw_ptab_base<CurveK1>(bk, w * j);
        sw_ptab_base<CurveR1>(br, w * j);
        ed_ptab_base(be, w * j);
        for (u32 m = 1; m <= per; m++) {
            const size_t e = j * per + (m - 1);
            sw_ptab_entry<CurveK1>(&k1_gtab_w[e * 16], m, w, bk);
            sw_ptab_entry<CurveR1>(&r1_gtab_w[e * 16], m, w, br);
            ed_ptab_entry(&ed_btab_w[e * 24], m, w, be);
        }
    }
    ptab_describe(k1_gtab, k1_gtab_w.data(), w);
    ptab_describe(r1_gtab, r1_gtab_w.data(), w);
    ptab_describe(ed_btab, ed_btab_w.data(), w);
}

extern "C" int hostsim_set_gwin(int w) {
    if (w < (int)kPTabMinWin || w > 14) return 1;
    g_gwin = (u32)w;
    k1_gtab_w.clear();
    ensure_tables();
    return 0;
}

// ---- lane-group emulation: the roles of a signature are OS threads meeting at a pthread barrier (GroupCtx::sync) and
//      sharing the mailbox / table arrays exactly as the warps of a block share them in shared memory ----
struct GroupShared {
    std::vector<Q4> mb;
    std::vector<u32> sc;
    std::vector<Q4> tab;
    pthread_barrier_t bar;
    int roles;
    explicit GroupShared(int roles_, int tab_chunks)
        : mb((size_t)kMbSlots * 2 * kGroupSigs), sc((size_t)kScWords * kGroupSigs), tab((size_t)tab_chunks * kGroupSigs), roles(roles_) {
        pthread_barrier_init(&bar, nullptr, (unsigned)roles);
    }
    ~GroupShared() { pthread_barrier_destroy(&bar); }
    GroupCtx ctx(int role) {
        GroupCtx g;
        g.role = role;
        g.mb = mb.data();
        g.sc = sc.data();
        g.host_sync = [](void* b) { pthread_barrier_wait((pthread_barrier_t*)b); };
        g.host_arg = &bar;
        return g;
    }
    TabRef tabref() { return TabRef{tab.data(), (u32)kGroupSigs}; }
    template <class Fn>
    void run(Fn fn) {  // fn(role, ctx) on `roles` threads
        std::vector<std::thread> th;
        for (int r = 1; r < roles; r++) th.emplace_back([&, r] { fn(r, ctx(r)); });
        fn(0, ctx(0));
        for (auto& t : th) t.join();
    }
};

extern "C" {

int hostsim_group_unit(int op, const uint32_t* in, size_t n, uint32_t* out) {
    ensure_tables();
    int in_w, out_w;
    unit_shape(op, in_w, out_w);
    const bool ed = op == SIGOPS_UNIT_ED_GROUP_MULPT;
    GroupShared sh(ed ? kGroupRolesEd : kGroupRolesSw, kGroupSwTabChunks);
    sh.run([&](int, GroupCtx g) {
        TabRef tab = sh.tabref();
        for (size_t i = 0; i < n; i++) {
            u32 a[32], r[17];
            for (int j = 0; j < 32; j++) a[j] = j < in_w ? in[i * in_w + j] : 0u;
            for (int j = 0; j < 17; j++) r[j] = 0;
            bool writer;
            if (op == SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL)
                writer = unit_double_mul_g<CurveK1>(r, a, a + 8, a + 16, tab, k1_gtab, g);
            else if (op == SIGOPS_UNIT_R1_GROUP_DOUBLE_MUL)
                writer = unit_double_mul_g<CurveR1>(r, a, a + 8, a + 16, tab, r1_gtab, g);
            else
                writer = unit_ed_mulpt_g(r, a, tab, g);
            if (writer)
                for (int j = 0; j < out_w; j++) out[i * out_w + j] = r[j];
            g.sync();
        }
    });
    return 0;
}

int hostsim_group_ecrecover(int curve, const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out, uint8_t* status) {
    ensure_tables();
    GroupShared sh(kGroupRolesSw, kGroupSwTabChunks);
    sh.run([&](int, GroupCtx g) {
        TabRef tab = sh.tabref();
        for (size_t i = 0; i < n; i++) {
            u32 sig_w[16], msg_w[8], out_w[16], st = 0;
            memcpy(sig_w, sigs + 64 * i, 64);
            memcpy(msg_w, msgs + 32 * i, 32);
            const bool writer = curve == 0 ? sw_ecrecover_group<CurveK1>(sig_w, msg_w, out_w, &st, tab, k1_gtab, g)
                                           : sw_ecrecover_group<ColdProducts<CurveR1> >(sig_w, msg_w, out_w, &st, tab, r1_gtab, g);
            if (writer) {
                memcpy(out + 64 * i, out_w, 64);
                if (status) status[i] = (uint8_t)st;
            }
            g.sync();
        }
    });
    return 0;
}

int hostsim_group_ed25519_verify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* valid) {
    ensure_tables();
    GroupShared sh(kGroupRolesEd, kGroupEdTabChunks);
    sh.run([&](int role, GroupCtx g) {
        TabRef tab = sh.tabref();
        for (size_t i = 0; i < n; i++) {
            u32 sig_w[16], msg_w[8], pk_w[8];
            memcpy(sig_w, sigs + 64 * i, 64);
            memcpy(msg_w, msgs + 32 * i, 32);
            memcpy(pk_w, pks + 32 * i, 32);
            const u32 v = (i & 1) ? ed_verify_group<true>(sig_w, msg_w, pk_w, tab, ed_btab, g) : ed_verify_group<false>(sig_w, msg_w, pk_w, tab, ed_btab, g);
            if (role == 0) valid[i] = (uint8_t)v;
            g.sync();
        }
    });
    return 0;
}

int hostsim_unit(int op, const uint32_t* in, size_t n, uint32_t* out) {
    if (op >= SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL && op <= SIGOPS_UNIT_ED_GROUP_MULPT) return hostsim_group_unit(op, in, n, out);
    ensure_tables();
    int in_w, out_w;
    unit_shape(op, in_w, out_w);
    std::vector<Q4> scratch(kEdTabChunks);
    TabRef tab{scratch.data(), 1};
    for (size_t i = 0; i < n; i++) {
        u32 a[32], r[17];
        for (int j = 0; j < 32; j++) a[j] = j < in_w ? in[i * in_w + j] : 0u;
        for (int j = 0; j < 17; j++) r[j] = 0;
        unit_dispatch(op, r, a, tab, k1_gtab, r1_gtab, ed_btab);
        for (int j = 0; j < out_w; j++) out[i * out_w + j] = r[j];
    }
    return 0;
}

int hostsim_unit_shape(int op, int* in_w, int* out_w) {
    unit_shape(op, *in_w, *out_w);
    return 0;
}

struct HostIO {
    const uint8_t* sigs;
    const uint8_t* msgs;
    uint8_t* out;
    uint8_t* status;
    size_t n, first;
    void load(int j, u32* sig_w, u32* msg_w) const {
        size_t i = first + (size_t)j;
        if (i >= n) i = n - 1;
        memcpy(sig_w, sigs + 64 * i, 64);
        memcpy(msg_w, msgs + 32 * i, 32);
    }
    void store(int j, const u32* out_w, u32 st) const {
        const size_t i = first + (size_t)j;
        if (i >= n) return;
        memcpy(out + 64 * i, out_w, 64);
        if (status) status[i] = (uint8_t)st;
    }
};

// batch = signatures per shared-inversion batch (1..kSwBatch), so the tests cover full, partial and single batches
int hostsim_ecrecover(int curve, const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out, uint8_t* status) {
    ensure_tables();
    std::vector<Q4> scratch(kSwBatchChunks);
    TabRef tab{scratch.data(), 1};
    HostIO io{sigs, msgs, out, status, n, 0};
    int batch = kSwBatch;
    for (size_t i = 0; i < n;) {
        io.first = i;
        const int B = (int)std::min<size_t>((size_t)batch, n - i);
        i += (size_t)B;
        if (curve == 0)
            sw_ecrecover_batch<CurveK1, false>(B, io, tab, k1_gtab);
        else
            sw_ecrecover_batch<CurveR1, false>(B, io, tab, r1_gtab);
        batch = batch == 1 ? kSwBatch : batch - 1;  // 8, 7, ..., 1, 8, ...: every batch size gets exercised
    }
    return 0;
}

struct HostEdIO {
    const uint8_t *sigs, *msgs, *pks;
    uint8_t* valid;
    size_t n, first;
    void load(int j, u32* sig_w, u32* msg_w, u32* pk_w) const {
        size_t i = first + (size_t)j;
        if (i >= n) i = n - 1;
        memcpy(sig_w, sigs + 64 * i, 64);
        memcpy(msg_w, msgs + 32 * i, 32);
        memcpy(pk_w, pks + 32 * i, 32);
    }
    void store(int j, u32 v) const {
        const size_t i = first + (size_t)j;
        if (i < n) valid[i] = (uint8_t)v;
    }
};

int hostsim_ed25519_verify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* valid) {
    ensure_tables();
    std::vector<Q4> scratch(kEdBatchChunks);
    TabRef tab{scratch.data(), 1};
    HostEdIO io{sigs, msgs, pks, valid, n, 0};
    int batch = kEdBatch;
    for (size_t i = 0; i < n;) {
        io.first = i;
        const int B = (int)std::min<size_t>((size_t)batch, n - i);
        i += (size_t)B;
        ed_verify_batch<false>(B, io, tab, ed_btab);
        batch = batch == 1 ? kEdBatch : batch - 1;
    }
    return 0;
}

int hostsim_ed25519_verify_msgs(const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, const uint8_t* pks, size_t n,
                                int strict, uint8_t* valid) {
    ensure_tables();
    std::vector<Q4> scratch(kEdTabChunks);
    TabRef tab{scratch.data(), 1};
    for (size_t i = 0; i < n; i++) {
        u32 sig_w[16], pk_w[8];
        memcpy(sig_w, sigs + 64 * i, 64);
        memcpy(pk_w, pks + 32 * i, 32);
        valid[i] = (uint8_t)ed_verify_msg<false>(sig_w, msg_bytes + off[i], (size_t)(off[i + 1] - off[i]), pk_w, strict != 0, tab,
                                                 ed_btab);
    }
    return 0;
}

int hostsim_sha256(const uint8_t* data, const uint64_t* off, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) {
        u32 d[8];
        sha256_ram(d, data + off[i], (size_t)(off[i + 1] - off[i]));
        memcpy(out + 32 * i, d, 32);
    }
    return 0;
}
}
