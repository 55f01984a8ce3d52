// odeb_math.cuh -- device-side small-vector math of the B200 step path.
// Operation order follows the reference (citations per function) so that a build with -fmad=false is
// bit-identical to the reference's x86-64 build for +,-,*,/,sqrt and single-precision atan2 (fdlibm's algorithm, below); sin/cos and the
// double-precision atan2 are CUDA libm (<= 2 ulp).
// dVector3 = 4 reals, dMatrix3 = 3 rows x 4 (include/ode/common.h:270-275), quaternion (w,x,y,z).
#ifndef ODEB_MATH_CUH
#define ODEB_MATH_CUH
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>


#if defined(ODEB_DOUBLE)
typedef double Real;
#define RSQRT(x) sqrt(x)
#define RFABS(x) fabs(x)
#define RSIN(x) sin(x)
#define RCOS(x) cos(x)
#define RATAN2(y, x) atan2((y), (x))
#define RCOPYSIGN(a, b) copysign(a, b)
#else
typedef float Real;
#define RSQRT(x) sqrtf(x)
#define RFABS(x) fabsf(x)
#define RSIN(x) sinf(x)
#define RCOS(x) cosf(x)
#define RATAN2(y, x) odeb_atan2f_fdlibm((y), (x))
#define RCOPYSIGN(a, b) copysignf(a, b)
#endif
#define R_(x) ((Real)(x))

// atan2f exactly as the reference's host libm computes it.  The reference calls atan2f in cullPoints (box.cpp:305) and in the hinge /
// universal / motor angle read-outs (joint.cpp:449 ...); CUDA's atan2f is within 2 ulp of it, which is enough to make cullPoints keep a
// different corner once in ~10^5 box pairs (two candidates a rounding error apart in angle).  glibc up to 2.40 computes atan2f with the
// fdlibm algorithm in plain float arithmetic (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c): argument reduction onto one of four
// intervals, an 11-term odd/even polynomial, hi/lo table constants.  The same operations in the same order (no FMA: -fmad=false, IEEE
// division) give the same bits on the device: checked against the host's atan2f on 2e8 inputs (random bit patterns, lattices with
// symmetric angles, extreme ratios) by tests/test_capi.py::test_atan2f_replica_matches_host_libm through odeb_test_atan2f.
__host__ __device__ __forceinline__ int odeb_f2i(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
__host__ __device__ __forceinline__ float odeb_i2f(int i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
__host__ __device__ inline float odeb_atanf_fdlibm(float x)
{
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const int hx = odeb_f2i(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {                                    // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {                                     // |x| < 0.4375
        if (ix < 0x31000000) return x;                         // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {                                 // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
    const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
    if (id < 0) return x - x * (s1 + s2);
    const float ahi = id == 0 ? hi0 : id == 1 ? hi1 : id == 2 ? hi2 : hi3, alo = id == 0 ? lo0 : id == 1 ? lo1 : id == 2 ? lo2 : lo3;
    const float r = ahi - ((x * (s1 + s2) - alo) - x);
    return hx < 0 ? -r : r;
}
__host__ __device__ inline float odeb_atan2f_fdlibm(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int hx = odeb_f2i(x), ix = hx & 0x7fffffff, hy = odeb_f2i(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;      // NaN
    if (hx == 0x3f800000) return odeb_atanf_fdlibm(y);         // x = 1
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);         // 2 * sign(x) + sign(y)
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : m == 1 ? -pi_o_4 - tiny : m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
        return m == 0 ? 0.0f : m == 1 ? -0.0f : m == 2 ? pi + tiny : -pi - tiny;
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;                     // |y / x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0f;                      // |y| / x < -2^60
    else z = odeb_atanf_fdlibm(fabsf(y / x));
    if (m == 0) return z;
    if (m == 1) return odeb_i2f(odeb_f2i(z) ^ (int)0x80000000u);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}
#define R_INF ((Real)INFINITY)

// include/ode/common.h:282-284 / :332-334
__host__ __device__ __forceinline__ Real rrecip(Real x) { return R_(1.0) / x; }
__host__ __device__ __forceinline__ Real rrecipsqrt(Real x) { return R_(1.0) / RSQRT(x); }

// include/ode/odemath.h:212-215  (a . b with strides)
__host__ __device__ __forceinline__ Real dot3(const Real *a, const Real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__host__ __device__ __forceinline__ Real dot3s(const Real *a, int sa, const Real *b, int sb) { return a[0] * b[0] + a[sa] * b[sb] + a[2 * sa] * b[2 * sb]; }
// include/ode/odemath.h:236-246
__host__ __device__ __forceinline__ void cross3(Real *r, const Real *a, const Real *b)
{
    Real r0 = a[1] * b[2] - a[2] * b[1], r1 = a[2] * b[0] - a[0] * b[2], r2 = a[0] * b[1] - a[1] * b[0];
    r[0] = r0; r[1] = r1; r[2] = r2;
}
// dMultiply0_331: r = A(3x4) * b          include/ode/odemath.h:323-330
__host__ __device__ __forceinline__ void mul0_331(Real *r, const Real *A, const Real *b)
{
    Real r0 = dot3(A, b), r1 = dot3(A + 4, b), r2 = dot3(A + 8, b);
    r[0] = r0; r[1] = r1; r[2] = r2;
}
// dMultiply1_331: r = A^T * b             include/ode/odemath.h:332-339
__host__ __device__ __forceinline__ void mul1_331(Real *r, const Real *A, const Real *b)
{
    Real r0 = dot3s(A, 4, b, 1), r1 = dot3s(A + 1, 4, b, 1), r2 = dot3s(A + 2, 4, b, 1);
    r[0] = r0; r[1] = r1; r[2] = r2;
}
// dMultiply0_333: R = A * B               include/ode/odemath.h:378-383 (rows via dMultiply0_133 = B^T a)
__host__ __device__ __forceinline__ void mul0_333(Real *r, const Real *A, const Real *B)
{
    for (int i = 0; i < 3; i++) mul1_331(r + 4 * i, B, A + 4 * i);
}
// dMultiply2_333: R = A * B^T             include/ode/odemath.h:392-397
__host__ __device__ __forceinline__ void mul2_333(Real *r, const Real *A, const Real *B)
{
    for (int i = 0; i < 3; i++) mul0_331(r + 4 * i, B, A + 4 * i);
}
// dInvertMatrix3                          include/ode/odemath.h:451-505
__host__ __device__ __forceinline__ Real invert3(Real *dst, const Real *ma)
{
    Real det = ma[0] * (ma[5] * ma[10] - ma[9] * ma[6]) - ma[1] * (ma[4] * ma[10] - ma[8] * ma[6]) + ma[2] * (ma[4] * ma[9] - ma[8] * ma[5]);
    if (det == 0) return 0;
    Real dr = rrecip(det);
    dst[0] = (ma[5] * ma[10] - ma[6] * ma[9]) * dr;
    dst[1] = (ma[9] * ma[2] - ma[1] * ma[10]) * dr;
    dst[2] = (ma[1] * ma[6] - ma[5] * ma[2]) * dr;
    dst[4] = (ma[6] * ma[8] - ma[4] * ma[10]) * dr;
    dst[5] = (ma[0] * ma[10] - ma[8] * ma[2]) * dr;
    dst[6] = (ma[4] * ma[2] - ma[0] * ma[6]) * dr;
    dst[8] = (ma[4] * ma[9] - ma[8] * ma[5]) * dr;
    dst[9] = (ma[8] * ma[1] - ma[0] * ma[9]) * dr;
    dst[10] = (ma[0] * ma[5] - ma[1] * ma[4]) * dr;
    return det;
}
// dxSafeNormalize3 + dxNormalize3         ode/src/odemath.cpp:95-163, ode/src/odemath.h:36-45
__host__ __device__ __forceinline__ void normalize3(Real *a)
{
    Real a0 = RFABS(a[0]), a1 = RFABS(a[1]), a2 = RFABS(a[2]);
    int idx;
    if (a1 > a0) idx = (a2 > a1) ? 2 : 1;
    else if (a2 > a0) idx = 2;
    else {
        if (!(a0 > R_(0.0))) { a[0] = R_(1.0); a[1] = R_(0.0); a[2] = R_(0.0); return; }
        idx = 0;
    }
    if (idx == 0) {
        Real rc = rrecip(a0), y = a[1] * rc, z = a[2] * rc, l = rrecipsqrt(R_(1.0) + y * y + z * z);
        a[1] = y * l; a[2] = z * l; a[0] = RCOPYSIGN(l, a[0]);
    } else if (idx == 1) {
        Real rc = rrecip(a1), x = a[0] * rc, z = a[2] * rc, l = rrecipsqrt(R_(1.0) + x * x + z * z);
        a[0] = x * l; a[2] = z * l; a[1] = RCOPYSIGN(l, a[1]);
    } else {
        Real rc = rrecip(a2), x = a[0] * rc, y = a[1] * rc, l = rrecipsqrt(R_(1.0) + x * x + y * y);
        a[0] = x * l; a[1] = y * l; a[2] = RCOPYSIGN(l, a[2]);
    }
}
// dxSafeNormalize4 + dxNormalize4         ode/src/odemath.cpp:200-220, ode/src/odemath.h:48-57
__host__ __device__ __forceinline__ void normalize4(Real *a)
{
    Real l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) { l = rrecipsqrt(l); a[0] *= l; a[1] *= l; a[2] *= l; a[3] *= l; }
    else { a[0] = R_(1.0); a[1] = a[2] = a[3] = R_(0.0); }
}
// dxPlaneSpace                            ode/src/odemath.cpp:224-251 (double compare against 0.70710678118654752440)
__host__ __device__ __forceinline__ void plane_space(const Real *n, Real *p, Real *q)
{
    if (RFABS(n[2]) > 0.70710678118654752440) {
        Real a = n[1] * n[1] + n[2] * n[2], k = rrecipsqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        Real a = n[0] * n[0] + n[1] * n[1], k = rrecipsqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}
// dRfromQ                                 ode/src/rotation.cpp:236-255
__host__ __device__ __forceinline__ void r_from_q(Real *R, const Real *q)
{
    Real qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;            R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]); R[3] = R_(0.0);
    R[4] = 2 * (q[1] * q[2] + q[0] * q[3]); R[5] = 1 - qq1 - qq3;            R[6] = 2 * (q[2] * q[3] - q[0] * q[1]); R[7] = R_(0.0);
    R[8] = 2 * (q[1] * q[3] - q[0] * q[2]); R[9] = 2 * (q[2] * q[3] + q[0] * q[1]); R[10] = 1 - qq1 - qq2;          R[11] = R_(0.0);
}
// dQfromR                                 ode/src/rotation.cpp:258-305
__host__ __device__ __forceinline__ void q_from_r(Real *q, const Real *R)
{
    Real tr = R[0] + R[5] + R[10], s;
    if (tr >= 0) {
        s = RSQRT(tr + 1); q[0] = R_(0.5) * s; s = R_(0.5) * rrecip(s);
        q[1] = (R[9] - R[6]) * s; q[2] = (R[2] - R[8]) * s; q[3] = (R[4] - R[1]) * s;
        return;
    }
    int c;
    if (R[5] > R[0]) c = (R[10] > R[5]) ? 2 : 1; else c = (R[10] > R[0]) ? 2 : 0;
    if (c == 0) {
        s = RSQRT((R[0] - (R[5] + R[10])) + 1); q[1] = R_(0.5) * s; s = R_(0.5) * rrecip(s);
        q[2] = (R[1] + R[4]) * s; q[3] = (R[8] + R[2]) * s; q[0] = (R[9] - R[6]) * s;
    } else if (c == 1) {
        s = RSQRT((R[5] - (R[10] + R[0])) + 1); q[2] = R_(0.5) * s; s = R_(0.5) * rrecip(s);
        q[3] = (R[6] + R[9]) * s; q[1] = (R[1] + R[4]) * s; q[0] = (R[2] - R[8]) * s;
    } else {
        s = RSQRT((R[10] - (R[0] + R[5])) + 1); q[3] = R_(0.5) * s; s = R_(0.5) * rrecip(s);
        q[1] = (R[8] + R[2]) * s; q[2] = (R[6] + R[9]) * s; q[0] = (R[4] - R[1]) * s;
    }
}
// dDQfromW                                ode/src/rotation.cpp:308-317
__host__ __device__ __forceinline__ void dq_from_w(Real *dq, const Real *w, const Real *q)
{
    dq[0] = R_(0.5) * (-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]);
    dq[1] = R_(0.5) * (w[0] * q[0] + w[1] * q[3] - w[2] * q[2]);
    dq[2] = R_(0.5) * (-w[0] * q[3] + w[1] * q[0] + w[2] * q[1]);
    dq[3] = R_(0.5) * (w[0] * q[2] - w[1] * q[1] + w[2] * q[0]);
}
// dQMultiply0..3                          ode/src/rotation.cpp:191-228
__host__ __device__ __forceinline__ void qmul0(Real *qa, const Real *qb, const Real *qc)
{
    qa[0] = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
    qa[1] = qb[0] * qc[1] + qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
    qa[2] = qb[0] * qc[2] + qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
    qa[3] = qb[0] * qc[3] + qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
}
__host__ __device__ __forceinline__ void qmul1(Real *qa, const Real *qb, const Real *qc)
{
    qa[0] = qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2] + qb[3] * qc[3];
    qa[1] = qb[0] * qc[1] - qb[1] * qc[0] - qb[2] * qc[3] + qb[3] * qc[2];
    qa[2] = qb[0] * qc[2] - qb[2] * qc[0] - qb[3] * qc[1] + qb[1] * qc[3];
    qa[3] = qb[0] * qc[3] - qb[3] * qc[0] - qb[1] * qc[2] + qb[2] * qc[1];
}
__host__ __device__ __forceinline__ void qmul2(Real *qa, const Real *qb, const Real *qc)
{
    qa[0] = qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2] + qb[3] * qc[3];
    qa[1] = -qb[0] * qc[1] + qb[1] * qc[0] - qb[2] * qc[3] + qb[3] * qc[2];
    qa[2] = -qb[0] * qc[2] + qb[2] * qc[0] - qb[3] * qc[1] + qb[1] * qc[3];
    qa[3] = -qb[0] * qc[3] + qb[3] * qc[0] - qb[1] * qc[2] + qb[2] * qc[1];
}
// dRand / dRandInt                        ode/src/misc.cpp:35-49, :78-139
__host__ __device__ __forceinline__ unsigned odeb_rand(unsigned *seed) { *seed = 1664525u * (*seed) + 1013904223u; return *seed; }
__host__ __device__ __forceinline__ int odeb_rand_int(unsigned *seed, int n)
{
    unsigned r = odeb_rand(seed), un = (unsigned)n;
    if (un <= 0x10u) {
        r ^= r >> 16; r ^= r >> 8; r ^= r >> 4;
        if (un <= 2u) { r ^= r >> 2; r ^= r >> 1; return (int)(r & (un >> 1)); }
        if (un <= 4u) { r ^= r >> 2; return (int)(((r & 3u) * un) >> 2); }
        return (int)(((r & 0xFu) * un) >> 4);
    }
    if (un <= 0x100u) { r ^= r >> 16; r ^= r >> 8; return (int)(((r & 0xFFu) * un) >> 8); }
    if (un <= 0x10000u) { r ^= r >> 16; return (int)(((r & 0xFFFFu) * un) >> 16); }
    return (int)(((unsigned long long)r * un) >> 32);
}
#endif
