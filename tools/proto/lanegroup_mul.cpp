// PROTOTYPE for round 2 -- not part of libsigops, not on any product path.
// Algorithm check of the planned lane-group field product ("Round-2 plan"): four lanes cooperate on one
// 256-bit product mod p = 2^256 - c.  Inputs are replicated in the four lanes; lane j multiplies the 64-bit slice
// (a[2j], a[2j+1]) by all of b (16 wide MACs, a 10-limb partial row at limb offset 2j); two butterfly rounds of
// shfl_xor (emulated here by indexing the lane array) sum the partial rows so that every lane ends with the full
// 16-limb product, and every lane folds the high half redundantly (8 MACs by c, c < 2^33 split as 2^32 + c_lo).
// The program compares the emulation with a plain 8x8 schoolbook product + the same fold on random and extreme inputs
// for secp256k1's p (c = 2^32 + 977) and 2^255-19 handled as 2^256 - 38.
//   g++ -O2 -std=c++17 -o /tmp/lanegroup_mul tools/proto/lanegroup_mul.cpp && /tmp/lanegroup_mul
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned __int128 u128;

struct Lane {  // what one thread would hold in registers
  u32 a[8], b[8];
  u32 row[18];  // partial sums at absolute limb positions (16 product limbs + headroom)
};

// lane j: rows 2j and 2j+1 of the schoolbook product, placed at limb offset 2j
static void partial_rows(Lane& L, int j) {
  memset(L.row, 0, sizeof L.row);
  for (int r = 0; r < 2; ++r) {
    u64 carry = 0;
    u32 ai = L.a[2 * j + r];
    for (int k = 0; k < 8; ++k) {
      u64 t = (u64)ai * L.b[k] + L.row[2 * j + r + k] + carry;
      L.row[2 * j + r + k] = (u32)t;
      carry = t >> 32;
    }
    L.row[2 * j + r + 8] = (u32)carry;  // position was still zero
  }
}

// one butterfly round: every lane adds its partner's rows (what 10 shfl_xor + an add.cc chain would do; only the
// limbs the partner can have non-zero are exchanged)
static void butterfly(Lane g[4], int mask) {
  u32 got[4][18];
  for (int j = 0; j < 4; ++j) memcpy(got[j], g[j ^ mask].row, sizeof got[j]);  // shfl_xor
  for (int j = 0; j < 4; ++j) {
    u64 carry = 0;
    for (int k = 0; k < 18; ++k) {
      u64 t = (u64)g[j].row[k] + got[j][k] + carry;
      g[j].row[k] = (u32)t;
      carry = t >> 32;
    }
  }
}

// fold 512 -> 256 bits modulo 2^256 - c, c = c_hi * 2^32 + c_lo with c_hi in {0, 1}; result fully reduced
static void fold(const u32 wide[16], u32 c_lo, u32 c_hi, const u32 p[8], u32 out[8]) {
  u32 acc[10] = {0};
  for (int k = 0; k < 8; ++k) acc[k] = wide[k];
  for (int pass = 0; pass < 3; ++pass) {
    u32 hi[8];
    if (pass == 0) memcpy(hi, wide + 8, sizeof hi);
    else { memset(hi, 0, sizeof hi); hi[0] = acc[8]; hi[1] = acc[9]; acc[8] = acc[9] = 0; }
    u64 carry = 0;
    for (int k = 0; k < 10; ++k) {  // acc += hi * c_lo + (hi << 32) * c_hi
      u64 t = (u64)acc[k] + carry;
      if (k < 8) t += (u64)hi[k] * c_lo;
      u64 t2 = (u32)t + (u64)((c_hi && k >= 1 && k <= 8) ? hi[k - 1] : 0);
      acc[k] = (u32)t2;
      carry = (t >> 32) + (t2 >> 32);
    }
  }
  for (int rep = 0; rep < 2; ++rep) {  // conditional subtraction
    u32 d[8]; u64 borrow = 0;
    for (int k = 0; k < 8; ++k) { u64 t = (u64)acc[k] - p[k] - borrow; d[k] = (u32)t; borrow = (t >> 32) & 1; }
    if (!borrow) memcpy(acc, d, sizeof d);
  }
  memcpy(out, acc, 32);
}

static void schoolbook(const u32 a[8], const u32 b[8], u32 w[16]) {
  memset(w, 0, 64);
  for (int i = 0; i < 8; ++i) {
    u64 carry = 0;
    for (int k = 0; k < 8; ++k) { u64 t = (u64)a[i] * b[k] + w[i + k] + carry; w[i + k] = (u32)t; carry = t >> 32; }
    w[i + 8] = (u32)carry;
  }
}

static int check(const u32 a[8], const u32 b[8], u32 c_lo, u32 c_hi, const u32 p[8]) {
  Lane g[4];
  for (int j = 0; j < 4; ++j) { memcpy(g[j].a, a, 32); memcpy(g[j].b, b, 32); partial_rows(g[j], j); }
  butterfly(g, 1);
  butterfly(g, 2);
  u32 w[16], ref[8];
  schoolbook(a, b, w);
  fold(w, c_lo, c_hi, p, ref);
  int bad = 0;
  for (int j = 0; j < 4; ++j) {
    if (memcmp(g[j].row, w, 64) || g[j].row[16] || g[j].row[17]) ++bad;  // every lane holds the full product
    u32 out[8];
    fold(g[j].row, c_lo, c_hi, p, out);
    if (memcmp(out, ref, 32)) ++bad;
  }
  return bad;
}

static void hex(const u32 x[8]) { for (int k = 7; k >= 0; --k) printf("%08x", x[k]); }

int main(int argc, char** argv) {
  const u32 pk1[8] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  // 2^255-19 handled through 2^256 - 38 (the result is reduced below 2^256-38 here; the product kernels keep that
  // weak form and normalise once at the end)
  const u32 p38[8] = {0xFFFFFFDAu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  std::mt19937_64 rng(0x5160);
  long bad = 0, n = 0;
  u32 a[8], b[8];
  if (argc > 2 && !strcmp(argv[1], "--dump")) {  // "a b (a*b mod p_k1) (a*b mod 2^256-38)" per line, lane 3's copy: for an independent big-integer check
    for (int it = 0; it < atoi(argv[2]); ++it) {
      for (int k = 0; k < 8; ++k) { a[k] = (u32)rng(); b[k] = (u32)rng(); }
      if (it % 5 == 0) for (int k = 0; k < 8; ++k) a[k] = 0xFFFFFFFFu;
      if (it % 10 == 0) for (int k = 0; k < 8; ++k) b[k] = 0xFFFFFFFFu;
      Lane g[4];
      for (int j = 0; j < 4; ++j) { memcpy(g[j].a, a, 32); memcpy(g[j].b, b, 32); partial_rows(g[j], j); }
      butterfly(g, 1);
      butterfly(g, 2);
      u32 o1[8], o2[8];
      fold(g[3].row, 977, 1, pk1, o1);
      fold(g[3].row, 38, 0, p38, o2);
      hex(a); printf(" "); hex(b); printf(" "); hex(o1); printf(" "); hex(o2); printf("\n");
    }
    return 0;
  }
  for (int it = 0; it < 200000; ++it) {
    for (int k = 0; k < 8; ++k) { a[k] = (u32)rng(); b[k] = (u32)rng(); }
    if (it % 7 == 0) for (int k = 0; k < 8; ++k) a[k] = 0xFFFFFFFFu;
    if (it % 11 == 0) for (int k = 0; k < 8; ++k) b[k] = 0xFFFFFFFFu;
    if (it % 13 == 0) memset(a, 0, 32);
    if (it % 17 == 0) { memcpy(a, pk1, 32); a[0] -= 1; }
    bad += check(a, b, 977, 1, pk1);
    bad += check(a, b, 38, 0, p38);
    n += 2;
  }
  // the fold itself against 128-bit arithmetic on small cases: (x * 2^256) mod p == x * c for x < 2^32
  for (u32 x = 1; x < 2000; ++x) {
    u32 w[16] = {0}, out[8];
    w[8] = x;
    fold(w, 977, 1, pk1, out);
    u128 e = (u128)x * (((u128)1 << 32) + 977);
    if (out[0] != (u32)e || out[1] != (u32)(e >> 32) || out[2] != (u32)(e >> 64) || out[3]) ++bad;
    ++n;
  }
  printf("lanegroup_mul prototype: %ld checks, %ld mismatches\n", n, bad);
  return bad ? 1 : 0;
}
