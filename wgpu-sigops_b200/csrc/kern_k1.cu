// secp256k1 ecrecover: kernel instantiation + launcher (replaces src/wgsl/main/secp256k1_ecdsa_main*.wgsl).
#include "kern_sw.cuh"

namespace sigops {
int kl_k1_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                    const PTab& gtab) {
    return launch_ecrecover<CurveK1>(l, sigs, msgs, n, out, status, scratch, gtab);
}
int kl_k1_setup(int* max_blocks_per_sm) { return setup_ecrecover<CurveK1>(max_blocks_per_sm); }
}  // namespace sigops
