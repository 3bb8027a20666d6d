/*
 * sag_detmath.h -- deterministic double-precision sin / cos / atan2 / log.
 *
 * Why: the parity bar of this repo is "GPU == oracle, bit for bit".  +, -, *, /, sqrt are correctly rounded
 * IEEE operations on both sides (nvcc -fmad=false, gcc -ffp-contract=off), but libm's and CUDA's
 * transcendental functions differ in the last ulp, and the point robot's stiff yaw servo plus box contacts
 * amplify an ulp into 1e-4 m within a thousand steps.  These routines are written with plain IEEE
 * operations in a fixed order, so the CUDA kernels and the CPU oracle get identical bits from them.
 *
 * The algorithms are the classic fdlibm ones (Cody-Waite reduction by pi/2 + minimax kernels; argument
 * range splitting for atan; 2^k(1+f) decomposition for log).  Accuracy (< 1 ulp, checked against glibc in
 * tests/test_detmath.py) is the same as libm's, so neither side gains or loses fidelity to the reference,
 * which itself just calls numpy / MuJoCo's libm.
 *
 * Valid C99 and CUDA C++.  Used by oracle/sag_oracle.c and safe_adaptation_gym_b200/csrc/sag_core.cuh.
 */
#ifndef SAG_DETMATH_H
#define SAG_DETMATH_H
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SAG_DM __host__ __device__ __forceinline__
#else
#define SAG_DM static inline
#endif

SAG_DM double sag_dm_from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, sizeof(d));
  return d;
#endif
}
SAG_DM uint64_t sag_dm_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  memcpy(&b, &d, sizeof(b));
  return b;
#endif
}

/* sin and cos of x, |x| < ~1e5 (yaw angles and sim time); beyond that accuracy degrades gracefully */
SAG_DM void sag_sincos(double x, double* sn, double* cs) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double pio2_1t = 6.07710050650619224932e-11; /* pi/2 - pio2_1 */
  double fk = x * invpio2;
  fk = fk >= 0.0 ? fk + 0.5 : fk - 0.5;
  long long k = (long long)fk; /* round to nearest, ties away from zero */
  double dk = (double)k;
  double r = (x - dk * pio2_1) - dk * pio2_1t;
  double z = r * r;
  /* kernel sin */
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  double ks = r + r * z * (S1 + z * ps);
  /* kernel cos */
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  double kc = w + (((1.0 - w) - hz) + z * pc);
  int q = (int)(k & 3);
  double s = (q & 1) ? kc : ks;
  double c = (q & 1) ? ks : kc;
  if (q == 1 || q == 2) c = -c;
  if (q >= 2) s = -s;
  *sn = s;
  *cs = c;
}
SAG_DM double sag_sin(double x) { double s, c; sag_sincos(x, &s, &c); return s; }
SAG_DM double sag_cos(double x) { double s, c; sag_sincos(x, &s, &c); return c; }

/* atan(t), t >= 0 */
SAG_DM double sag_atan_pos(double t) {
  const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
               aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
               aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
               aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
  double hi = 0.0, lo = 0.0, x;
  int id;
  if (t < 0.4375) { id = -1; x = t; }
  else if (t < 1.1875) {
    if (t < 0.6875) { id = 0; x = (2.0 * t - 1.0) / (2.0 + t); hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17; }
    else { id = 1; x = (t - 1.0) / (t + 1.0); hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17; }
  } else {
    if (t < 2.4375) { id = 2; x = (t - 1.5) / (1.0 + 1.5 * t); hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17; }
    else { id = 3; x = -1.0 / t; hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17; }
  }
  double z = x * x;
  double w = z * z;
  double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  return hi - ((x * (s1 + s2) - lo) - x);
}

/* atan2(y, x) in (-pi, pi]; atan2(0, 0) = 0 like numpy.angle */
SAG_DM double sag_atan2(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16, pio2 = 1.57079632679489655800e+00;
  if (y == 0.0) return x >= 0.0 ? 0.0 : pi;
  if (x == 0.0) return y > 0.0 ? pio2 : -pio2;
  double ay = y < 0.0 ? -y : y, ax = x < 0.0 ? -x : x;
  double z = sag_atan_pos(ay / ax);
  if (x > 0.0) return y > 0.0 ? z : -z;
  return y > 0.0 ? pi - (z - pi_lo) : (z - pi_lo) - pi;
}

/* natural log of a positive normal double */
SAG_DM double sag_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t b = sag_dm_bits(x);
  int32_t hx = (int32_t)(b >> 32);
  uint32_t lx = (uint32_t)b;
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int32_t i = (hx + 0x95f64) & 0x100000;
  uint64_t nb = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx; /* normalise x or x/2 */
  k += i >> 20;
  double f = sag_dm_from_bits(nb) - 1.0;
  double dk = (double)k;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

#endif
