// Philox4x32-10 counter-based generator for the generated-RNG mode of the DF path.
// The reference draws U(-sqrt3, sqrt3) from the unseeded global MT19937 (digitalfilters.py:1344,
// 1361-1366, 1460-1467); there is no reference stream to reproduce, so the product defines one that
// any time shard can regenerate on its own (oracle/philox.py is its CPU restatement):
//   key     = (seed lo, seed hi)
//   counter = (pair lo, pair hi, s lo, comp + 4 * s hi)      pair = idx/2, idx = p*Kp + q, s = slab
//   words r0..r3: (r0, r1) -> point idx = 2*pair, (r2, r3) -> point 2*pair + 1
//   u = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53 ;  x = -sqrt3 + (2 sqrt3) * u   (numpy's uniform())
#pragma once
#include <cstdint>

namespace podsdf {

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
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;   // IMAD.WIDE.U32: hi and lo at once
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[i];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[i];
    c3 = (uint32_t)p0;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

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

// the two random numbers of points (2*pair, 2*pair+1) of component `comp`, slab `s`
__device__ __forceinline__ void philox_pair(const PhiloxKeys& k, uint64_t pair, uint64_t s, int comp,
                                            double& x0, double& x1) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)s, (uint32_t)comp + 4u * (uint32_t)(s >> 32), k, r);
  x0 = philox_to_uniform(r[0], r[1]);
  x1 = philox_to_uniform(r[2], r[3]);
}

}  // namespace podsdf
