// Positional fixed-base tables: the u1*G (ECDSA) and s*B (EdDSA) halves of the double-scalar multiplications.
//
// The reference uploads a 16-entry generator table with every call (src/precompute.rs:14-69, src/secp256k1_ecdsa.rs:108)
// and walks it inside a 256-step double-and-add (src/wgsl/secp256k1_curve.wgsl:388-447 `projective_fixed_mul`).  Here the
// scalar is cut into `pos` signed windows of `w` bits and window j has a table of its own,
//     entry (j, m - 1) = m * 2^(w j) * 2^-D * G        m = 1 .. 2^(w-1),
// so the fixed-base half is `pos` = floor(256 / w) + 1 mixed additions and NO doublings: 12 at w = 22 where 12-bit windows
// sharing the loop's doublings took 22.  The sum goes into the accumulator BEFORE the double-scalar loop, whose D
// doublings (128 secp256k1, 256 secp256r1, 252 ed25519) undo the factor 2^-D baked into the entries -- G has prime order,
// 2^-D is taken mod the group order -- so the loop itself only carries the variable-base windows.
// Size: pos * 2^(w-1) * 64 B (96 B for ed25519's affine Niels triples) per curve: 1.6 + 1.6 + 2.4 GB at the default w = 22
// (of 180 GB), resident in HBM for the life of the context and generated on the device at init (0.4 s); a lookup is one
// 64-byte (96-byte) gather per addition, ~40 GB/s at full throughput.  SIGOPS_GWIN selects w (4..24); measured on B200
// (profiles/r02_ptab_sweep.txt), secp256k1 M sigs/s / init s / GB: w = 16 47.5 / 0.04 / 0.1, 20 48.6 / 0.09 / 1.5,
// 22 48.9 / 0.38 / 5.6, 24 49.2 / 1.26 / 20.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace sigops {

struct PTab {
    const uint32_t* base;  // [pos][2^(w-1)][16 or 24 words]
    uint32_t w, pos;
    uint32_t pat[9];  // sum_j 2^(w j + w - 1): added to the scalar, it turns the unsigned windows into signed digits
};

// doublings of the double-scalar loops the tables are scaled for (tools/gen_tables.py bakes 2^-D * G)
static constexpr int kPTabShiftK1 = 128, kPTabShiftR1 = 256, kPTabShiftEd = 252;
static constexpr uint32_t kPTabMinWin = 4, kPTabMaxWin = 24, kPTabDefaultWin = 22;

static inline uint32_t ptab_positions(uint32_t w) { return 256u / w + 1u; }  // w * pos >= 257: the top window absorbs the carry
static inline size_t ptab_entries(uint32_t w) { return (size_t)ptab_positions(w) << (w - 1); }

static inline void ptab_describe(PTab& t, const uint32_t* base, uint32_t w) {
    t.base = base;
    t.w = w;
    t.pos = ptab_positions(w);
    for (int i = 0; i < 9; i++) t.pat[i] = 0;
    for (uint32_t j = 0; j < t.pos; j++) {
        const uint32_t bit = w * j + w - 1;  // < 288
        t.pat[bit >> 5] |= 1u << (bit & 31);
    }
}

}  // namespace sigops
