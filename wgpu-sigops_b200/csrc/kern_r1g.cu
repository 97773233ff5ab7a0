// secp256r1 ecrecover, lane-group kernel (small batches): field products out of line (ColdProducts, curve_sw.cuh)
#include "kern_group_sw.cuh"

namespace sigops {
int kl_r1_group(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, const PTab& gtab) {
    return launch_ecrecover_group<ColdProducts<CurveR1> >(l, sigs, msgs, n, out, status, gtab);
}
int kl_r1_group_setup(int* max_blocks_per_sm) { return setup_ecrecover_group<ColdProducts<CurveR1> >(max_blocks_per_sm); }
}  // namespace sigops
