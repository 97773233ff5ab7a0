// secp256k1 ecrecover, lane-group kernel with the field products out of line (ColdProducts, curve_sw.cuh): the flavour the host
// picks once a request puts a block on most SMs.
#include "kern_group_sw.cuh"

namespace sigops {
int kl_k1_group_cold(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, const PTab& gtab) {
    return launch_ecrecover_group<ColdProducts<CurveK1> >(l, sigs, msgs, n, out, status, gtab);
}
int kl_k1_group_cold_setup(int* max_blocks_per_sm) { return setup_ecrecover_group<ColdProducts<CurveK1> >(max_blocks_per_sm); }
}  // namespace sigops
