// mxb_device.cuh — per-photon arithmetic of the MARXS hot path, fp64, sm_100a.
//
// Every routine restates one reference routine (file:line cited) in EXACTLY the
// operation order of oracle/marxs_oracle.py: explicit left-to-right sums over
// x,y,z, no re-association.  Two builds of the same source:
//   strict : nvcc -fmad=false              -> bit-identical to the oracle except libm calls
//   fast   : nvcc -DMXB_FAST (fmad on)     -> FMA contraction + reciprocal-multiply normalisation
#pragma once
#include <stdint.h>      // NVRTC builds get a stand-in from mxb_jit.cpp
#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#define MXB_DEV __device__ __forceinline__
// -DMXB_SHARE_CODE turns parallel transport / Rodrigues rotation into real functions (one copy, called)
// for very large programs; measured slower on C3 (r01: 5.9 -> 6.9 ms), so it is off by default.
#ifdef MXB_SHARE_CODE
#define MXB_DEV_BIG __device__ __noinline__
#else
#define MXB_DEV_BIG __device__ __forceinline__
#endif

namespace mxb {

// libm routines are 50-300 instructions each once inlined (argument reduction, slow paths) and a fused
// program uses them at many sites.  -DMXB_SHARE_LIBM makes them one shared copy each; measured on C2 / C3
// (r01) the calls cost 1-3 % more than the instruction-cache pressure they remove, so inline is the default.
#ifdef MXB_SHARE_LIBM
#define MXB_LIBM __device__ __noinline__
#else
#define MXB_LIBM __device__ __forceinline__     // measured: calls in the hot loop cost more than the code they save
#endif
#if defined(MXB_FAST) && !defined(MXB_LIBM_ASIN)
#define MXB_OWN_ASIN 1   // m_acos / m_asin are defined below (after fast_rsqrt)
#else
MXB_LIBM double m_acos(double x) { return acos(x); }
MXB_LIBM double m_asin(double x) { return asin(x); }
#endif
MXB_LIBM double m_sin(double x) { return sin(x); }
MXB_LIBM double m_log(double x) { return log(x); }
MXB_LIBM double m_exp(double x) { return exp(x); }
MXB_LIBM double m_pow(double x, double y) { return pow(x, y); }
MXB_LIBM double m_atan2(double y, double x) { return atan2(y, x); }
MXB_LIBM void libm_sincos(double x, double* s, double* c) { sincos(x, s, c); }
#if !defined(MXB_FAST) || defined(MXB_LIBM_SINCOS)
MXB_LIBM void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
#endif                                                  // (fast build: m_sincos is defined below, after its tables)

constexpr double kEnergy2Wave = 1.2398419292004202e-06;  // marxs/__init__.py:14 (keV -> mm)
constexpr double kHcKevNm = 1.2398419843320026;          // astropy u.spectral(): keV -> nm
constexpr double kHcMultilayer = 1.23984282;             // multiLayerMirror.py:158
constexpr double kTwoPi = 6.283185307179586;

struct V3 {
    double x, y, z;
};

// ---- division / sqrt -----------------------------------------------------------------
// strict build: IEEE-754 correctly rounded operators (bit parity with numpy).
// fast build: hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits) + two Newton steps: <= 1 ulp
// typical, no denormal / special-case slow path (inputs here are lengths and dot products of
// unit-scale vectors; zero still yields inf/NaN like the IEEE forms do).
MXB_DEV double fast_rcp(double x) {
#ifdef MXB_FAST
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}
MXB_DEV double fast_rsqrt(double x) {   // only used in the fast build
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-hx * r, r, 0.5);
    return fma(r, e, r);
}
// sqrt of a finite z >= 0 from the Newton rsqrt (0 -> 0; negative or NaN -> NaN), faithful: the IEEE sqrt sequence of
// the compiler carries a slow path for denormals / infinities at every site
MXB_DEV double sqrt_nn(double z) {
    const double r = fast_rsqrt(z);
    double s = z * r;
    s = fma(fma(-s, s, z), 0.5 * r, s);
    return (z == 0.0) ? 0.0 : s;
}
MXB_DEV double m_sqrt(double x) {     // strict build: IEEE (bit parity with np.sqrt)
#if defined(MXB_FAST) && !defined(MXB_IEEE_SQRT)
    return sqrt_nn(x);
#else
    return sqrt(x);
#endif
}
#ifdef MXB_OWN_ASIN
// asin / acos of the fast build.  libm's versions are 65 instructions per site, a third of them MOVs that build
// polynomial coefficients; here the coefficients sit in the constant bank and the large-argument branch uses the
// Newton rsqrt above.  asin(s) = s + s z P(z), z = s^2 <= 1/4, P = degree-11 interpolant of (asin(s) - s) / s^3 at
// the Chebyshev nodes of [0, 1/4] (fitted at 60 digits with mpmath, tools/fit_asin.py; <= 1.5 ulp);
// |x| > 1/2: asin(|x|) = pi/2 - 2 asin(sqrt((1 - |x|) / 2)) (fdlibm's reduction).  Errors <= 2.5 ulp.
__constant__ double kAsinP[12] = {0.1666666666666665, 0.07500000000020764, 0.044642857103423646, 0.03038194736709848,
                                  0.02237204763174451, 0.017355259955786323, 0.013929652902326633, 0.011875494382636922,
                                  0.0078029494773533175, 0.01603551434914882, -0.010749050339697808, 0.028169218060881414};
MXB_DEV double asin_poly(double s, double z) {
    double p = kAsinP[11];
#pragma unroll
    for (int k = 10; k >= 0; --k) p = fma(p, z, kAsinP[k]);
    return fma(s * z, p, s);
}
MXB_DEV double m_acos(double x) {
    const double ax = fabs(x);
    if (ax > 0.5) {
        const double z = (1.0 - ax) * 0.5;          // exact; negative for |x| > 1 -> NaN like np.arccos
        const double r = 2.0 * asin_poly(sqrt_nn(z), z);
        return (x < 0.0) ? (3.141592653589793116 - r) + 1.2246467991473532e-16 : r;
    }
    return 1.5707963267948966 - (asin_poly(x, x * x) - 6.123233995736766e-17);   // NaN propagates
}
MXB_DEV double m_asin(double x) {
    const double ax = fabs(x);
    if (ax > 0.5) {
        const double z = (1.0 - ax) * 0.5;
        const double r = 1.5707963267948966 - (2.0 * asin_poly(sqrt_nn(z), z) - 6.123233995736766e-17);
        return (x < 0.0) ? -r : r;
    }
    return asin_poly(x, x * x);
}
#endif
MXB_DEV double div(double a, double b) {
#ifdef MXB_FAST
    const double r = fast_rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);    // one correction step: faithful quotient
#else
    return a / b;
#endif
}

// two quotients by the same divisor: one reciprocal (fast build), each quotient corrected like div()
MXB_DEV void div2(double a0, double a1, double b, double& q0, double& q1) {
#ifdef MXB_FAST
    const double r = fast_rcp(b);
    const double t0 = a0 * r, t1 = a1 * r;
    q0 = fma(fma(-b, t0, a0), r, t0);
    q1 = fma(fma(-b, t1, a1), r, t1);
#else
    q0 = a0 / b;
    q1 = a1 / b;
#endif
}

MXB_DEV double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MXB_DEV V3 cross(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename P>
MXB_DEV V3 ld3(P p) { return V3{p[0], p[1], p[2]}; }

// math/utils.py:150-164 norm_vector
MXB_DEV V3 normalize(const V3& a) {
#ifdef MXB_FAST
    const double r = fast_rsqrt(dot(a, a));
    return V3{a.x * r, a.y * r, a.z * r};
#else
    const double n = sqrt(dot(a, a));
    return V3{a.x / n, a.y / n, a.z / n};
#endif
}

// sin and cos; the fast build uses the Taylor series for the tiny scatter angles (|x| < 2^-7:
// truncation error x^10/10! < 1e-28), libm otherwise
#ifdef MXB_FAST
// libm sincos for the (rare) large scatter angle: one out-of-line copy instead of ~110 inlined instructions per site
__device__ __noinline__ double2 sincos_large(double x) {
    double s, c;
    libm_sincos(x, &s, &c);
    return make_double2(s, c);
}
#endif
#ifdef MXB_FAST
// polynomial coefficients live in the constant bank (DFMA reads c[3][..] directly): a double literal costs two
// MOV instructions at every use, which is 15 % of the instruction stream of a large fused program
__constant__ double kSinT[4] = {2.7557319223985893e-06, -1.9841269841269841e-04, 8.3333333333333332e-03,
                                -1.6666666666666666e-01};
__constant__ double kCosT[3] = {2.4801587301587302e-05, -1.3888888888888889e-03, 4.1666666666666664e-02};
#endif
MXB_DEV void sincos_small(double x, double* s, double* c) {
#ifdef MXB_FAST
    if (fabs(x) < 0.0078125) {
        const double x2 = x * x;
#ifndef MXB_NO_CONST_TABLES
        *s = x * fma(x2, fma(x2, fma(x2, fma(x2, kSinT[0], kSinT[1]), kSinT[2]), kSinT[3]), 1.0);
        *c = fma(x2, fma(x2, fma(x2, fma(x2, kCosT[0], kCosT[1]), kCosT[2]), -0.5), 1.0);
#else
        *s = x * fma(x2, fma(x2, fma(x2, fma(x2, 2.7557319223985893e-06, -1.9841269841269841e-04),
                                     8.3333333333333332e-03), -1.6666666666666666e-01), 1.0);
        *c = fma(x2, fma(x2, fma(x2, fma(x2, 2.4801587301587302e-05, -1.3888888888888889e-03),
                                 4.1666666666666664e-02), -0.5), 1.0);
#endif
        return;
    }
    const double2 sc = sincos_large(x);
    *s = sc.x;
    *c = sc.y;
#else
    m_sincos(x, s, c);
#endif
}

MXB_DEV double clip01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }  // NaN propagates like np.clip

// ---------------------------------------------------------------------------
// math/geometry.py:211-261 FinitePlane.intersect (+ :376-380 CircularHole)
// g: c[3] ex[3] ey[3] ez[3] Ly Lz.   Returns hit; ip/loc are valid only on hit
// (rect_ok reports the rectangle test for the CircularHole NaN quirk).
// ---------------------------------------------------------------------------
template <typename P>
MXB_DEV bool plane_intersect(P g, const V3& pos, const V3& dir, bool circular, V3& ip, double& l0,
                             double& l1, bool* rect_ok = nullptr) {
    // 14 doubles as seven 16-byte loads (parameter blocks and facet rows are 16-byte aligned)
    const double2 g01 = g.ld2(0), g23 = g.ld2(1), g45 = g.ld2(2), g67 = g.ld2(3), g89 = g.ld2(4),
                  gab = g.ld2(5), gcd = g.ld2(6);
    const double cx = g01.x, cy = g01.y, cz = g23.x;
    const double ex = g23.y, ey = g45.x, ez = g45.y;
    const double k_num = (cx - pos.x) * ex + (cy - pos.y) * ey + (cz - pos.z) * ez;
    const double k_den = dir.x * ex + dir.y * ey + dir.z * ez;
    const double k = div(k_num, k_den);
    ip.x = pos.x + k * dir.x;
    ip.y = pos.y + k * dir.y;
    ip.z = pos.z + k * dir.z;
    const double vx = ip.x - cx, vy = ip.y - cy, vz = ip.z - cz;
    l0 = vx * g67.x + vy * g67.y + vz * g89.x;
    l1 = vx * g89.y + vy * gab.x + vz * gab.y;
    bool hit = (k_den != 0.0) & (k >= 0.0) & (fabs(l0) <= gcd.x) & (fabs(l1) <= gcd.y);
    if (rect_ok) *rect_ok = hit;
    if (circular) hit = hit & (sqrt(l0 * l0 + l1 * l1) <= 1.0);
    return hit;
}

// Unit-length tracking (fast build only).  The reference re-normalises directions in every
// routine; a vector this kernel has just produced by normalisation, rotation or the grating
// equation already has |v| = 1 to rounding, and normalising it again changes it by <= 1 ulp.  The
// fast build carries one flag per photon and skips those renormalisations; the strict build
// (bit parity) never does.
#ifdef MXB_FAST
constexpr bool kTrackUnit = true;
#else
constexpr bool kTrackUnit = false;
#endif
MXB_DEV V3 normalize_unless(bool unit, const V3& a) {
    if (kTrackUnit && unit) return a;
    return normalize(a);
}

// ---------------------------------------------------------------------------
// math/polarization.py:90-170 parallel_transport (identity when |d1 x d2| <= 1e-8)
// ---------------------------------------------------------------------------
// the reference's construction: pol in the frame {s, d1 x s, d1} re-expressed in {s, d2 x s, d2}, s = d1 x d2 / |d1 x d2|
MXB_DEV V3 parallel_transport_frame(const V3& d1, const V3& d2, const V3& pol) {
    V3 s = cross(d1, d2);
#ifdef MXB_FAST
    const double ns2 = dot(s, s);
    if (!(ns2 > 1e-16)) { if (ns2 == ns2) return pol; }   // |ns| <= 1e-8 -> identity; NaN falls through
    const double r = fast_rsqrt(ns2);
    s = V3{s.x * r, s.y * r, s.z * r};
#else
    const double ns = sqrt(dot(s, s));
    if (fabs(ns) <= 1e-8) return pol;
    s = V3{s.x / ns, s.y / ns, s.z / ns};
#endif
    const V3 p_in = cross(d1, s);
    const V3 p_out = cross(d2, s);
    const double a = dot(s, pol), b = dot(p_in, pol), c = dot(d1, pol);
    return V3{s.x * a + p_out.x * b + d2.x * c, s.y * a + p_out.y * b + d2.y * c,
              s.z * a + p_out.z * b + d2.z * c};
}
#ifdef MXB_FAST
// rays turned by more than ~170 degrees (cold): one out-of-line copy, arguments by value
__device__ __noinline__ double3 parallel_transport_reversed(double3 a, double3 b, double3 p) {
    const V3 r = parallel_transport_frame(V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z}, V3{p.x, p.y, p.z});
    return make_double3(r.x, r.y, r.z);
}
#endif
MXB_DEV_BIG V3 parallel_transport(const V3& dir_old, const V3& dir_new, const V3& pol, bool old_unit = false,
                                  bool new_unit = false) {
    const V3 d1 = normalize_unless(old_unit, dir_old);
    const V3 d2 = normalize_unless(new_unit, dir_new);
#if defined(MXB_FAST) && !defined(MXB_PT_FRAME)
    // Fast build: the frame map is the rotation about s = d1 x d2 that takes d1 to d2.  With the UNNORMALISED
    // a = d1 x d2 (|a| = sin t) and c = d1 . d2 Rodrigues' formula is
    //     R p = c p + a x p + a (a . p) / (1 + c),
    // the same map without normalising the (possibly tiny) cross product and without building the two frames:
    // 34 instead of 50 fp64 instructions, and better conditioned than the frame form for small deflections
    // (error ~ eps instead of eps / |a|).  Near a reversal (1 + c -> 0) it is the worse form: those rays take the
    // frame construction.
    const V3 a = cross(d1, d2);
    const double ns2 = dot(a, a);
    if (!(ns2 > 1e-16)) { if (ns2 == ns2) return pol; }   // |d1 x d2| <= 1e-8 -> identity; NaN falls through
    const double c = dot(d1, d2);
    if (c < -0.98) {
        const double3 r = parallel_transport_reversed(make_double3(d1.x, d1.y, d1.z), make_double3(d2.x, d2.y, d2.z),
                                                      make_double3(pol.x, pol.y, pol.z));
        return V3{r.x, r.y, r.z};
    }
    const V3 axp = cross(a, pol);
    const double k = dot(a, pol) * fast_rcp(1.0 + c);
    return V3{fma(a.x, k, fma(pol.x, c, axp.x)), fma(a.y, k, fma(pol.y, c, axp.y)), fma(a.z, k, fma(pol.z, c, axp.z))};
#else
    return parallel_transport_frame(d1, d2, pol);
#endif
}

// sin / cos of u turns, i.e. of the angle u * 2 * pi the reference forms for a uniform azimuth (scatter.py:135-137)
#if defined(MXB_FAST) && !defined(MXB_LIBM_SINCOS)
// (sin x - x) / x^3 and (cos x - 1 + x^2 / 2) / x^4 in z = x^2 on |x| <= pi / 4 (tools/fit_sincos.py: 1.2 / 1.7 ulp)
__constant__ double kSinQ[6] = {-0.16666666666666666, 0.0083333333333307, -0.00019841269836387345, 2.755731591191116e-06,
                                -2.5051092507061385e-08, 1.59153232122714e-10};
__constant__ double kCosQ[6] = {0.041666666666666664, -0.0013888888888887241, 2.4801587298533456e-05, -2.755731715246704e-07,
                                2.087612165887116e-09, -1.1380876948169717e-11};
#endif
MXB_DEV void sincos_turn(double u, double* s, double* c) {
#if defined(MXB_FAST) && !defined(MXB_LIBM_SINCOS)
    // quarter-turn reduction is EXACT in turns (no Cody-Waite / Payne-Hanek as for an angle in radians): 30 instead of
    // libm's ~110 instructions per site
    const double q = rint(4.0 * u);
    const double r = fma(-0.25, q, u);              // |r| <= 1/8, exact
    const double x = r * kTwoPi;
    const double z = x * x;
    double ps = kSinQ[5], pc = kCosQ[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        ps = fma(ps, z, kSinQ[k]);
        pc = fma(pc, z, kCosQ[k]);
    }
    const double sn = fma(x * z, ps, x);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int k = (int)q;
    const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
#else
    m_sincos(u * 2 * 3.141592653589793, s, c);
#endif
}

#if defined(MXB_FAST) && !defined(MXB_LIBM_SINCOS)
// sin / cos of a general angle in the fast build: Cody-Waite reduction by pi/2 = hi + lo with two fma (exact products;
// the remainder is good to 1e-28 |k| for |x| < 1e5), then the same constant-bank polynomials as sincos_turn: ~45
// instead of libm's ~110 instructions per site, a third of which build coefficients with MOV pairs; <= 3 ulp
// (test_device_math_vs_numpy).  Larger arguments take libm's routine out of line.
MXB_DEV void m_sincos(double x, double* s, double* c) {
    if (!(fabs(x) < 1e5)) {              // also NaN / inf
        const double2 sc = sincos_large(x);
        *s = sc.x;
        *c = sc.y;
        return;
    }
    const double q = rint(x * 0.63661977236758138);
    double r = fma(-q, 1.5707963267948966, x);
    r = fma(-q, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = kSinQ[5], pc = kCosQ[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        ps = fma(ps, z, kSinQ[k]);
        pc = fma(pc, z, kCosQ[k]);
    }
    const double sn = fma(r * z, ps, r);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int k = (int)q;
    const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
}
#endif

// math/rotations.py:50-87 axangle2mat applied TRANSPOSED (scatter.py:60,68), from sin / cos of the angle.
// PERP: axis . v == 0 by construction (axis = v x something); AXIS_UNIT: |axis| == 1 already (fast build only)
template <bool PERP = false, bool AXIS_UNIT = false>
MXB_DEV_BIG V3 rotate_T_sc(const V3& axis, double s, double c, const V3& v) {
#if defined(MXB_FAST) && !defined(MXB_ROT_MATRIX)
    // Fast build: R^T v = c v - s (a x v) + (1 - c)(a . v) a with the axis normalisation folded into the
    // coefficients: no 3x3 matrix (27-45 instead of 52 fp64 instructions)
    {
        const double inv = (kTrackUnit && AXIS_UNIT) ? 1.0 : fast_rsqrt(dot(axis, axis));
        const V3 axv = cross(axis, v);
        const double sk = -(s * inv);
        V3 out{fma(axv.x, sk, v.x * c), fma(axv.y, sk, v.y * c), fma(axv.z, sk, v.z * c)};
        if (!PERP) {
            const double k = ((1.0 - c) * inv) * (inv * dot(axis, v));
            out = V3{fma(axis.x, k, out.x), fma(axis.y, k, out.y), fma(axis.z, k, out.z)};
        }
        return out;
    }
#endif
    const V3 a = normalize(axis);  // axes / np.linalg.norm(axes): true division in strict build
    const double C = 1 - c;
    const double x = a.x, y = a.y, z = a.z;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double xC = x * C, yC = y * C, zC = z * C;
    const double xyC = x * yC, yzC = y * zC, zxC = z * xC;
    const double r00 = x * xC + c, r01 = xyC - zs, r02 = zxC + ys;
    const double r10 = xyC + zs, r11 = y * yC + c, r12 = yzC - xs;
    const double r20 = zxC - ys, r21 = yzC + xs, r22 = z * zC + c;
    return V3{r00 * v.x + r10 * v.y + r20 * v.z, r01 * v.x + r11 * v.y + r21 * v.z,
              r02 * v.x + r12 * v.y + r22 * v.z};
}
// the angle is a scatter angle, tiny in practice (Taylor series in the fast build, libm out of line for the rare
// large one)
template <bool PERP = false, bool AXIS_UNIT = false>
MXB_DEV_BIG V3 axangle_rotate_T(const V3& axis, double angle, const V3& v) {
    double s, c;
    sincos_small(angle, &s, &c);
    return rotate_T_sc<PERP, AXIS_UNIT>(axis, s, c, v);
}

// np.interp arithmetic, clamped ends (oracle interp1d_np)
template <typename P>
MXB_DEV double interp_clamped(P xp, P fp, int n, double x) {
    if (x >= xp[n - 1]) return fp[n - 1];
    if (x <= xp[0]) return fp[0];
    int lo = 0, hi = n - 1;  // invariant xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    const double slope = div(fp[lo + 1] - fp[lo], xp[lo + 1] - xp[lo]);
    return slope * (x - xp[lo]) + fp[lo];
}

// the bracket interp_clamped finds, for several value arrays over the SAME knots and query (MultiLayerEfficiency looks
// three columns up at one position): mode 0 interior (segment lo), 1 clamped to the first value, 2 to the last
template <typename P>
MXB_DEV int interp_bracket(P xp, int n, double x, int& lo) {
    lo = 0;
    if (x >= xp[n - 1]) return 2;
    if (x <= xp[0]) return 1;
    int hi = n - 1;  // invariant xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    return 0;
}
// value of one array on a bracket: the arithmetic of interp_clamped (x0 = xp[lo], dx = xp[lo + 1] - xp[lo])
template <typename P>
MXB_DEV double interp_on_bracket(P fp, int n, int mode, int lo, double x, double x0, double dx) {
    if (mode == 2) return fp[n - 1];
    if (mode == 1) return fp[0];
    const double slope = div(fp[lo + 1] - fp[lo], dx);
    return slope * (x - x0) + fp[lo];
}

// index i in [0, n-2] with xk[i] <= x < xk[i+1]  (searchsorted(side='right') - 1, clipped)
template <typename P>
MXB_DEV int bracket(P xk, int n, double x) {
    int lo = -1, hi = n;  // count of knots <= x is hi at the end
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xk[mid] <= x) lo = mid; else hi = mid;
    }
    int i = hi - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    return i;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (photon id lo, hi, slot, 0)
// ---------------------------------------------------------------------------
MXB_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                           uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

MXB_DEV double u01_from_bits(uint32_t hi, uint32_t lo) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    return (double)(b >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
}

// Box-Muller in fp64 (libm log / sincos).  In the fast build this is the EXTREME TAIL only (1 - u < 2^-20, one draw
// in a million), so it is one out-of-line routine: inlined at every draw site it was ~270 instructions of code per
// site that practically never run (instruction-cache pressure of large fused programs).  The strict build uses it
// for every normal and keeps it inline.
#ifdef MXB_FAST
__device__ __noinline__
#else
MXB_DEV
#endif
double2 box_muller_f64(double v, double u2) {     // by value: no address of a caller's variable escapes
    const double rad = sqrt(-2.0 * m_log(v));
    double s, c;
    m_sincos(kTwoPi * u2, &s, &c);
    return make_double2(rad * c, rad * s);
}

// kind 0: uniform [0,1); kind 1: standard normal (Box-Muller; the fast build evaluates it with the fp32
// SFU intrinsics like device_draw_normal_pair below, the extreme tail and the strict build in fp64)
MXB_DEV double device_draw(uint64_t seed, uint64_t photon_id, int slot, int kind) {
    uint32_t r[4];
    philox4x32_10((uint32_t)photon_id, (uint32_t)(photon_id >> 32), (uint32_t)slot, 0u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = u01_from_bits(r[0], r[1]);
    if (kind == 0) return u1;
    const double u2 = u01_from_bits(r[2], r[3]);
    const double v = 1.0 - u1;
#ifdef MXB_FAST
    if (v > 9.5367431640625e-07) {
        const float rad = sqrtf(-2.0f * __logf((float)v));
        return (double)(rad * __cosf(6.2831853f * ((float)u2 - 0.5f)));
    }
#endif
    return box_muller_f64(v, u2).x;
}

// two independent standard normals from ONE Philox call (both Box-Muller branches).
// The fast build evaluates log / sincos with the fp32 SFU intrinsics: a random variate only has
// to be a sample of N(0,1); 2^-24 resolution of the deviate is far below anything a photon
// distribution can resolve (KS at 1e9 samples sees ~3e-5).  The extreme tail (1-u < 2^-20) and
// the strict build use the fp64 libm forms.
MXB_DEV void device_draw_normal_pair(uint64_t seed, uint64_t photon_id, int slot, double& z0, double& z1) {
    uint32_t r[4];
    philox4x32_10((uint32_t)photon_id, (uint32_t)(photon_id >> 32), (uint32_t)slot, 0u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double v = 1.0 - u01_from_bits(r[0], r[1]);      // (0, 1]
    const double u2 = u01_from_bits(r[2], r[3]);
#ifdef MXB_FAST
    if (v > 9.5367431640625e-07) {
        const float rad = sqrtf(-2.0f * __logf((float)v));
        float s, c;
        __sincosf(6.2831853f * ((float)u2 - 0.5f), &s, &c);   // uniform angle in [-pi, pi)
        z0 = (double)(rad * c);
        z1 = (double)(rad * s);
        return;
    }
#endif
    const double2 z = box_muller_f64(v, u2);
    z0 = z.x;
    z1 = z.y;
}

}  // namespace mxb
