// secp256r1 ecrecover: kernel instantiation + launcher (replaces src/wgsl/main/secp256r1_ecdsa_main*.wgsl).
#include "kern_sw.cuh"

namespace sigops {
int kl_r1_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                    const PTab& gtab) {
    return launch_ecrecover<CurveR1>(l, sigs, msgs, n, out, status, scratch, gtab);
}
int kl_r1_setup(int* max_blocks_per_sm) { return setup_ecrecover<CurveR1>(max_blocks_per_sm); }
}  // namespace sigops
