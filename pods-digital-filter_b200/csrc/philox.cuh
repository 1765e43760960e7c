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

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&r)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

__device__ __forceinline__ double philox_to_uniform(uint32_t a, uint32_t b) {
  const double u = ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
  const double s3 = 1.7320508075688772;  // np.sqrt(3.0)
  // low + (high - low) * u with separate multiply and add, as numpy computes it (no contraction)
  return __dadd_rn(-s3, __dmul_rn(s3 - (-s3), u));
}

// the two random numbers of points (2*pair, 2*pair+1) of component `comp`, slab `s`
__device__ __forceinline__ void philox_pair(uint64_t seed, uint64_t pair, uint64_t s, int comp,
                                            double& x0, double& x1) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)s,
                (uint32_t)comp + 4u * (uint32_t)(s >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  x0 = philox_to_uniform(r[0], r[1]);
  x1 = philox_to_uniform(r[2], r[3]);
}

}  // namespace podsdf
