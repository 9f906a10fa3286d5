// pcg_section.cuh — vertical sections / pseudo-RHI cuts through a float64 volume.
//
// One thread per (sweep, line point):
//   * geometry of PRD.extract_section's per-sweep loop (pycwr/core/NRadar.py:1048-1060):
//     cartesian_xy_elevation_to_range_z (pycwr/core/transforms.py:224-250) — ground distance, the
//     4/3-earth inversion to slant range and beam height at the sweep's fixed elevation, azimuth;
//   * the strip interpolation of PRD._interpolate_ppi_strip_arrays (NRadar.py:833-913): azimuth table
//     extended by one wrapped ray on each side, right bisections in azimuth and range, the four-corner
//     weighted mean, then the four two-corner fall-backs in the reference's order (along range on the
//     lower ray, on the upper ray, along azimuth at the near gate, at the far gate).  Missing = NaN.
//
// The interpolation arithmetic is spelled with explicit round-to-nearest operations in the reference's
// evaluation order, so explicit (azimuth, range) targets give bit-identical values; with the fused
// geometry the targets carry the last-ulp differences between CUDA's and NumPy's cos / sin / atan2.
#pragma once
#include "pcg_kernels.cuh"

namespace pcg {

struct SectionArgs {
    VolumeView vol;
    int field;
    int npts;
    int geometry;          // 1: targets from (x, y); 0: explicit targets
    int per_sweep;         // explicit targets: [ne][npts] when set, else one row shared by every sweep
    double h, Re;
    const double* x;
    const double* y;
    const double* t_az;
    const double* t_rng;
    double fill;           // the volume's missing marker (treated like NaN)
    double* out_val;       // [ne][npts]
    double* out_az;        // [ne][npts]   (geometry mode only)
    double* out_rng;
    double* out_z;
};

// numpy's remainder for a positive divisor (npy_divmod): fmod, then shifted to the divisor's sign
__device__ __forceinline__ double floor_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) {
        if (r < 0.0) r = add(r, b);
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

__device__ __forceinline__ bool present(double v, double fill) { return isfinite(v) && v != fill; }

__global__ void __launch_bounds__(128) section_kernel(SectionArgs a) {
    const int sweep = blockIdx.y;
    const SweepDesc sw = a.vol.sweeps[sweep];
    const double* __restrict__ A = a.vol.az + sw.az_off;
    const double* __restrict__ R = a.vol.rng + sw.rng_off;
    const double* __restrict__ V = (const double*)a.vol.val[a.field] + sw.val_off;
    const int naz = sw.naz, ng = sw.ng;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.npts; p += gridDim.x * blockDim.x) {
        const size_t o = (size_t)sweep * a.npts + p;
        double taz, trg;
        if (a.geometry) {
            const double x = a.x[p], y = a.y[p];
            const double s = hypot(x, y);
            const double phi = div(s, a.Re);
            const double cphi = cos(phi), sphi = sin(phi);
            const double denom = sub(cphi, mul(sphi, sw.tan_e));
            const double rr = add(a.Re, a.h);
            const double z = sub(div(rr, denom), a.Re);
            const double g = add(a.Re, z);
            const double rsq = sub(add(mul(rr, rr), mul(g, g)), mul(mul(mul(2.0, rr), g), cphi));
            trg = (rsq != rsq) ? rsq : sqrt(rsq > 0.0 ? rsq : 0.0);        // np.maximum keeps NaN
            taz = mul(floor_mod(sub(kHalfPi, atan2(y, x)), mul(2.0, kPi)), kRad2Deg);
            a.out_az[o] = taz;
            a.out_rng[o] = trg;
            a.out_z[o] = z;
        } else {
            const size_t ti = a.per_sweep ? o : (size_t)p;
            taz = a.t_az[ti];
            trg = a.t_rng[ti];
        }

        double out = nan;
        const double t = floor_mod(taz, 360.0);
        if (naz >= 2 && ng >= 2 && isfinite(t) && isfinite(trg) && trg >= R[0] && trg <= R[ng - 1]) {
            // extended azimuth table: [A[naz-1]-360, A[0..naz-1], A[0]+360]
            auto ext = [&](int j) -> double {
                return j == 0 ? sub(A[naz - 1], 360.0) : (j == naz + 1 ? add(A[0], 360.0) : A[j - 1]);
            };
            int lo = 0, hi = naz + 2;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ext(mid) <= t) lo = mid + 1; else hi = mid;
            }
            const int ah = min(max(lo, 1), naz + 1), al = ah - 1;
            lo = 0; hi = ng;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (R[mid] <= trg) lo = mid + 1; else hi = mid;
            }
            const int gh = min(max(lo, 1), ng - 1), gl = gh - 1;
            const double az0 = ext(al), az1 = ext(ah), r0 = R[gl], r1 = R[gh];
            if (az1 > az0 && r1 > r0) {
                const int row0 = al == 0 ? naz - 1 : al - 1;             // ext row -> stored ray
                const int row1 = ah == naz + 1 ? 0 : ah - 1;
                const double m00 = V[(size_t)row0 * ng + gl], m01 = V[(size_t)row0 * ng + gh];
                const double m10 = V[(size_t)row1 * ng + gl], m11 = V[(size_t)row1 * ng + gh];
                const bool f00 = present(m00, a.fill), f01 = present(m01, a.fill);
                const bool f10 = present(m10, a.fill), f11 = present(m11, a.fill);
                const double wa1 = sub(az1, t), wa0 = sub(t, az0);
                const double wr1 = sub(r1, trg), wr0 = sub(trg, r0);
                if (f00 && f01 && f10 && f11) {
                    double num = mul(mul(m00, wa1), wr1);
                    num = add(num, mul(mul(m10, wa0), wr1));
                    num = add(num, mul(mul(m01, wa1), wr0));
                    num = add(num, mul(mul(m11, wa0), wr0));
                    out = div(num, mul(sub(az1, az0), sub(r1, r0)));
                }
                if (out != out && f00 && f01) out = div(add(mul(m00, wr1), mul(m01, wr0)), sub(r1, r0));
                if (out != out && f10 && f11) out = div(add(mul(m10, wr1), mul(m11, wr0)), sub(r1, r0));
                if (out != out && f00 && f10) out = div(add(mul(m00, wa1), mul(m10, wa0)), sub(az1, az0));
                if (out != out && f01 && f11) out = div(add(mul(m01, wa1), mul(m11, wa0)), sub(az1, az0));
            }
        }
        a.out_val[o] = out;
    }
}

}  // namespace pcg
