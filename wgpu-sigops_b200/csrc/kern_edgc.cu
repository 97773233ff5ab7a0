// ed25519 ecverify, lane-group kernel with the field products out of line: the flavour the host picks once a request puts a
// block on most SMs.
#include "kern_group_ed.cuh"

namespace sigops {
int kl_ed_group_cold(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, const PTab& btab) {
    return launch_ed_group<true>(l, sigs, msgs, pks, n, valid, btab);
}
int kl_ed_group_cold_setup(int* max_blocks_per_sm) { return setup_ed_group<true>(max_blocks_per_sm); }
}  // namespace sigops
