// Host-visible launchers of the kernels in kern_*.cu (one translation unit per kernel family so that the library builds in
// parallel).  Every function returns a cudaError_t as int (0 = success).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ptab.h"

namespace sigops {

struct KLaunch {
    int grid, tpb;
    cudaStream_t stream;
};

int kl_k1_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                    const PTab& gtab);
int kl_r1_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                    const PTab& gtab);
int kl_k1_setup(int* max_blocks_per_sm);
int kl_r1_setup(int* max_blocks_per_sm);
int kl_ed_verify(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, void* scratch,
                 const PTab& btab);
int kl_ed_verify_msgs(const KLaunch& l, const void* sigs, const uint8_t* msg_bytes, const unsigned long long* msg_off, const void* pks,
                      size_t n, int strict, uint8_t* valid, void* scratch, const PTab& btab);
int kl_ed_setup(int* max_blocks_per_sm, int* max_blocks_per_sm_msgs);
int kl_sha256_msgs(cudaStream_t st, const uint8_t* bytes, const unsigned long long* off, size_t n, uint32_t* out);
int kl_sha256_pubkeys(cudaStream_t st, const uint32_t* pubkeys, const uint8_t* status, size_t n, uint32_t* out);
// positional fixed-base table of one curve (0 secp256k1, 1 secp256r1, 2 ed25519): `bases` receives the pos window bases
// (16 / 24 words each), `tab` the pos * 2^(w-1) entries
int kl_gen_ptab(cudaStream_t st, int curve, uint32_t w, uint32_t* bases, uint32_t* tab);
int kl_unit(const KLaunch& l, int op, const uint32_t* in, size_t n, uint32_t* out, void* scratch, const PTab& k1g, const PTab& r1g,
            const PTab& edb);
int kl_unit_setup(int* max_blocks_per_sm);
// lane-group kernels (group.cuh): small batches, several cooperating warps per 32 signatures
int kl_k1_group(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, const PTab& gtab);
int kl_r1_group(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, const PTab& gtab);
int kl_ed_group(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, const PTab& btab);
// the same kernels with the field products out of line (ColdProducts, curve_sw.cuh); secp256r1 always runs that flavour
int kl_k1_group_cold(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, const PTab& gtab);
int kl_ed_group_cold(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, const PTab& btab);
int kl_k1_group_cold_setup(int* max_blocks_per_sm);
int kl_ed_group_cold_setup(int* max_blocks_per_sm);
int kl_k1_group_setup(int* max_blocks_per_sm);
int kl_r1_group_setup(int* max_blocks_per_sm);
int kl_ed_group_setup(int* max_blocks_per_sm);
int kl_unit_group(const KLaunch& l, int op, const uint32_t* in, size_t n, uint32_t* out, const PTab& k1g, const PTab& r1g);
int kl_imad_peak(int kind, int grid, int block, cudaStream_t st, uint32_t* sink, int iters, uint32_t seed);

}  // namespace sigops
