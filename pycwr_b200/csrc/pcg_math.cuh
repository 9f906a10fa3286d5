// pcg_math.cuh — device arithmetic for the beam-geometry inversion.
//
// The reference (pycwr/core/RadarGridC.pyx, compiled by gcc for baseline x86-64) evaluates
// its geometry in IEEE double with *unfused* multiply/add and glibc's libm.  The selected
// interpolation indices depend on comparisons of those doubles with table entries, so the
// index-determining chain here uses explicit round-to-nearest intrinsics (__dmul_rn /
// __dadd_rn / ... are never contracted into FMAs by nvcc, whatever -fmad says), IEEE sqrt and
// division, and transcendental functions that are correctly rounded wherever that is cheap
// (per-column quantities, amortised over the levels of the column).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pcg {

// RadarGridC.pyx:6-8
constexpr double kPi = 3.141592653589793;
constexpr double kHalfPi = kPi / 2.0;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kDeg2Rad = kPi / 180.0;
constexpr double kRad2Deg = 180.0 / kPi;

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// cos(x) for |x| < 0.25, evaluated as 1 - x^2/2 + T with the leading terms carried in
// double-double, so the single final rounding is the correct rounding of cos(x) except when
// the exact value lies within ~2^-65 of a rounding boundary.  phi = s/Re never exceeds 0.06 for
// ranges a weather radar can see; larger arguments use the CUDA libm cos.
__device__ __forceinline__ double cos_small(double x) {
    const double ax = fabs(x);
    if (!(ax < 0.25)) return cos(x);
    const double w = __dmul_rn(ax, ax);
    const double wl = __fma_rn(ax, ax, -w);
    double t = -1.0 / 87178291200.0;
    t = __fma_rn(t, w, 1.0 / 479001600.0);
    t = __fma_rn(t, w, -1.0 / 3628800.0);
    t = __fma_rn(t, w, 1.0 / 40320.0);
    t = __fma_rn(t, w, -1.0 / 720.0);
    t = __fma_rn(t, w, 1.0 / 24.0);
    t = __dmul_rn(__dmul_rn(w, w), t);                    // w^2/24 - w^3/720 + ...
    t = __fma_rn(__dmul_rn(w, wl), 1.0 / 12.0, t);        // d(w^2/24) from the low part of w
    t = __fma_rn(wl, -0.5, t);                            // -wl/2
    const double h = __dmul_rn(0.5, w);                   // exact
    const double s = __dsub_rn(1.0, h);
    const double e = __dsub_rn(__dsub_rn(1.0, s), h);     // exact: 1 - h = s + e
    return __dadd_rn(s, __dadd_rn(e, t));
}

// sin(x) for |x| < 0.25 with the -x^3/6 term in double-double (same idea as cos_small).
__device__ __forceinline__ double sin_small(double x) {
    const double ax = fabs(x);
    if (!(ax < 0.25)) return sin(x);
    const double w = __dmul_rn(ax, ax);
    const double wl = __fma_rn(ax, ax, -w);
    const double m6h = -0.16666666666666666;              // RN(-1/6)
    const double m6l = -9.25185853854297e-18;             // -1/6 - m6h
    const double ph = __dmul_rn(ax, w);
    const double pl = __fma_rn(ax, w, -ph);
    const double qh = __dmul_rn(ph, m6h);
    double ql = __fma_rn(ph, m6h, -qh);
    ql = __fma_rn(ph, m6l, ql);
    ql = __fma_rn(pl, m6h, ql);
    ql = __fma_rn(__dmul_rn(ax, wl), m6h, ql);
    double t = -1.0 / 1307674368000.0;
    t = __fma_rn(t, w, 1.0 / 6227020800.0);
    t = __fma_rn(t, w, -1.0 / 39916800.0);
    t = __fma_rn(t, w, 1.0 / 362880.0);
    t = __fma_rn(t, w, -1.0 / 5040.0);
    t = __fma_rn(t, w, 1.0 / 120.0);
    t = __dmul_rn(__dmul_rn(ax, __dmul_rn(w, w)), t);     // x^5/120 - x^7/5040 + ...
    const double s = __dadd_rn(ax, qh);
    const double e = __dadd_rn(__dsub_rn(ax, s), qh);     // exact: ax + qh = s + e
    const double r = __dadd_rn(s, __dadd_rn(e, __dadd_rn(ql, t)));
    return copysign(r, x);
}

// atan2 with the octant directions returned as the correctly rounded constants glibc gives
// there (a regular grid centred on the radar puts whole rows/columns/diagonals on them, and
// round-number azimuth tables then sit exactly on the comparison boundary).
__device__ __forceinline__ double atan2_exactdirs(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    if (ay == 0.0) return signbit(x) ? copysign(kPi, y) : y;
    if (ax == 0.0) return copysign(kHalfPi, y);
    if (ax == ay && ax <= 1.7976931348623157e308)
        return copysign(signbit(x) ? 2.356194490192345 : 0.7853981633974483, y);
    return atan2(y, x);
}

// RadarGridC.pyx:34-38
__device__ __forceinline__ double azimuth_from_xy(double x, double y) {
    double az = __dsub_rn(kHalfPi, atan2_exactdirs(y, x));
    if (az < 0.0) az = __dadd_rn(az, kTwoPi);
    return __dmul_rn(az, kRad2Deg);
}

// RadarGridC.pyx:236-252 — fill-aware 1-D interpolation, unfused like the reference.
__device__ __forceinline__ double interp_linear(double c, double c0, double c1, double d0,
                                                double d1, double fill) {
    if (c1 == c0) return fill;
    const bool f0 = (d0 == fill), f1 = (d1 == fill);
    if (f0 && f1) return fill;
    if (f0) return d1;
    if (f1) return d0;
    const double num = __dadd_rn(__dmul_rn(__dsub_rn(c1, c), d0), __dmul_rn(__dsub_rn(c, c0), d1));
    return __ddiv_rn(num, __dsub_rn(c1, c0));
}

// RadarGridC.pyx:41-51 — bisect_right.  `guess` comes from the (near-)regular spacing of the
// table; it is verified against the two bracketing entries (which the interpolation needs
// anyway) and repaired by +-1 steps, falling back to the reference's bisection.  For a sorted
// table the result is the unique g with v[g-1] <= t < v[g], i.e. identical to the bisection.
__device__ __forceinline__ int search_right(const double* __restrict__ v, int n, double t,
                                            int guess, bool sorted) {
    if (sorted) {
        int g = min(max(guess, 0), n);
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
            const bool lo_ok = (g == 0) || !(t < v[g - 1]);
            const bool hi_ok = (g == n) || (t < v[g]);
            if (lo_ok && hi_ok) return g;
            g += lo_ok ? 1 : -1;
        }
    }
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t < v[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

}  // namespace pcg
