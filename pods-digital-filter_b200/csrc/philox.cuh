// Philox4x32-10 counter-based generator for the generated-RNG mode of the DF path.
// The reference draws U(-sqrt3, sqrt3) from the unseeded global MT19937 (digitalfilters.py:1344,
// 1361-1366, 1460-1467); there is no reference stream to reproduce, so the product defines its own, which
// any time shard can regenerate on its own (oracle/philox.py is the CPU restatement of both kinds):
//
//   key = (seed lo, seed hi);  idx = p*Kp + q in the padded plane, pair = idx/2, half = idx & 1
//
//   kind PODSDF_RNG_U32 (default): one call serves ONE point x FOUR consecutive slabs, 32 random bits per number
//     counter = (idx lo, idx hi, sq lo, comp + 4 * sq hi),  sq = s/4 (s = slab);  word r[s & 3]
//     v = 1 + (r + 1/2) 2^-32  (bits dropped into the mantissa of [1, 2));
//     x = (v - 3/2) * (2 sqrt3)      -- v - 3/2 is exact, one rounding, symmetric about 0
//   kind PODSDF_RNG_U53 (round 1's stream, numpy's 53-bit uniform()): one call serves two points of one slab
//     counter = (pair lo, pair hi, s lo, comp + 4 * s hi);  (r0, r1) -> half 0, (r2, r3) -> half 1
//     u = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53 ;  x = -sqrt3 + (2 sqrt3) * u
//
// U32 halves the wide integer multiplies per random number (5 instead of 10 IMAD.WIDE) and needs two FP64
// operations per number instead of five.  Four slabs of one point per call is what a time filter wants: a thread
// owns one point and consumes its stream four steps at a time (df_time_ring8_kernel).
#pragma once
#include <cstdint>

namespace podsdf {

constexpr int kRngU53 = 0, kRngU32 = 1;

// round keys k + i*W are uniform across the grid: computed once on the host, passed by value
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
inline PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    k.k0[i] = a;
    k.k1[i] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& k, uint32_t (&r)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifndef PODSDF_PHILOX_ROUNDS
#define PODSDF_PHILOX_ROUNDS 10
#endif
#pragma unroll
  for (int i = 0; i < PODSDF_PHILOX_ROUNDS; ++i) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;   // IMAD.WIDE.U32: hi and lo at once
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[i];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[i];
    c3 = (uint32_t)p0;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// ---- kind U53 ------------------------------------------------------------------------------------
// n = (a >> 5) * 2^26 + (b >> 6) as an exact double without I2F (XU pipe): the two halves are dropped into
// the mantissa of 2^52 / 2^78 and the biases subtracted (exact: all values are integers below 2^53 * 2^26)
__device__ __forceinline__ double philox_to_uniform(uint32_t a, uint32_t b) {
  const double lo = __hiloint2double(0x43300000, (int)(b >> 6)) - 4503599627370496.0;              // 2^52 + lo
  const double hi = __hiloint2double(0x44D00000, (int)(a >> 5)) - 302231454903657293676544.0;      // 2^78 + hi*2^26
  const double n = hi + lo;                                                                         // < 2^53, exact
  const double s3 = 1.7320508075688772;  // np.sqrt(3.0)
  // low + (high - low) * u, u = n * 2^-53, with separate multiply and add as numpy computes it
  // ((high-low) * 2^-53) * n rounds identically to (high-low) * (n * 2^-53): scaling by 2^-53 is exact
  return __dadd_rn(-s3, __dmul_rn((s3 - (-s3)) * (1.0 / 9007199254740992.0), n));
}

// the two random numbers of points (2*pair, 2*pair+1) of component `comp`, slab `s` (kind U53)
__device__ __forceinline__ void philox_pair(const PhiloxKeys& k, uint64_t pair, uint64_t s, int comp,
                                            double& x0, double& x1) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)s, (uint32_t)comp + 4u * (uint32_t)(s >> 32), k, r);
  x0 = philox_to_uniform(r[0], r[1]);
  x1 = philox_to_uniform(r[2], r[3]);
}

// ---- kind U32 ------------------------------------------------------------------------------------
// 32 random bits -> U(-sqrt3, sqrt3): the word becomes the top of the mantissa of a double in [1, 2) (with the
// half-step bit below it, so the 2^32 levels are centred), then one exact subtraction and one multiplication
__device__ __forceinline__ double philox_u32_to_uniform(uint32_t r) {
  const double v = __hiloint2double((int)(0x3FF00000u | (r >> 12)), (int)((r << 20) | 0x80000u));
  return __dmul_rn(__dadd_rn(v, -1.5), 1.7320508075688772 + 1.7320508075688772);
}

// the four words of (point idx, slab quad sq, comp): r[s & 3] belongs to slab s = 4 sq + (s & 3)
__device__ __forceinline__ void philox_quad(const PhiloxKeys& k, uint64_t idx, uint64_t sq, int comp, uint32_t (&r)[4]) {
  philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)sq, (uint32_t)comp + 4u * (uint32_t)(sq >> 32), k, r);
}
__device__ __forceinline__ uint32_t philox_pick(const uint32_t (&r)[4], int w) {
  return w == 0 ? r[0] : (w == 1 ? r[1] : (w == 2 ? r[2] : r[3]));
}

// one random number of either kind (generic paths: wide filters, materialised slabs)
template <int KIND>
__device__ __forceinline__ double philox_value(const PhiloxKeys& k, uint64_t idx, uint64_t s, int comp) {
  if (KIND == kRngU32) {
    uint32_t r[4];
    philox_quad(k, idx, s >> 2, comp, r);
    return philox_u32_to_uniform(philox_pick(r, (int)(s & 3)));
  } else {
    double a, b;
    philox_pair(k, idx >> 1, s, comp, a, b);
    return (idx & 1) ? b : a;
  }
}

}  // namespace podsdf
