// pcg_products.cuh — elementwise companions of the gridding kernels (all HBM-bound streaming):
//   compose_kernel   compose_network_volume on already gridded per-radar volumes
//                    (pycwr/interp/RadarInterp.py:1399-1487), restated operation by operation in
//                    FP64 with the radar axis accumulated in order like np.sum(axis=0)
//   column_kernel    derive_cr / derive_vil / derive_et (pycwr/core/RadarProduct.py:26-85)
#pragma once
#include "pcg_math.cuh"

namespace pcg {

constexpr int kComposeMaxRadars = 64;

struct ComposeArgs {
    const double* vol[kComposeMaxRadars];      // [nz*ncol] each
    const double* qual[kComposeMaxRadars];     // [nz*ncol] each, or all NULL
    const double* w2d;                         // [nradar][ncol]: exp(-4 d^2/R^2) or d^2 (nearest); NULL for max
    int nradar;
    int method;                                // 0 max, 1 nearest, 2 exp_weighted, 3 quality_weighted
    int has_quality;
    long long ncol, nvox;
    double fill;
    double* out;
    int16_t* count;
};

__device__ __forceinline__ bool finite_not_fill(double v, double fill) {
    return (v == v) && (fabs(v) != __longlong_as_double(0x7ff0000000000000LL)) && (v != fill);
}

__global__ void __launch_bounds__(256) compose_kernel(const __grid_constant__ ComposeArgs p) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.nvox;
         i += (long long)gridDim.x * blockDim.x) {
        const long long col = i % p.ncol;
        int cnt = 0;
        double num = 0.0, den = 0.0;           // weighted sums / best value, best distance
        double best = -inf, bestd = inf;
        bool first = true;
        for (int r = 0; r < p.nradar; ++r) {
            const double v = p.vol[r][i];
            const bool ok = finite_not_fill(v, p.fill);
            cnt += ok ? 1 : 0;
            if (p.method == 0) {
                if (ok && v > best) best = v;
            } else if (p.method == 1) {
                const double d = ok ? p.w2d[(long long)r * p.ncol + col] : inf;
                if (d < bestd) { bestd = d; best = v; }       // argmin keeps the first minimum
            } else {
                double w = p.w2d[(long long)r * p.ncol + col];
                if (p.method == 3 && p.has_quality) w = mul(w, p.qual[r][i]);
                const double tn = ok ? mul(v, w) : 0.0;
                const double td = ok ? w : 0.0;
                if (first) { num = tn; den = td; first = false; }
                else { num = add(num, tn); den = add(den, td); }
            }
        }
        double o = p.fill;
        if (p.method == 0) { if (cnt > 0) o = best; }
        else if (p.method == 1) { if (bestd != inf && bestd == bestd) o = best; }
        else if (den > 0.0) o = div(num, den);
        p.out[i] = o;
        if (p.count) p.count[i] = (int16_t)cnt;
    }
}

struct ColumnArgs {
    const double* vol;          // [nz][ncol]
    const double* levels;       // [nz] (device)
    int nz;
    long long ncol;
    double fill, min_dbz, max_dbz_cap, threshold;
    double* cr;                 // each optional
    double* vil;
    double* et;
    unsigned char* topped;
    int nan_as_fill;            // the volume already carries NaN in its empty cells (the network tile kernel stores the
                                // dataset's marker itself): read them back as `fill`, which is what derive_et sees
};

__global__ void __launch_bounds__(256) column_kernel(const __grid_constant__ ColumnArgs p) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < p.ncol;
         c += (long long)gridDim.x * blockDim.x) {
        double best = -inf;
        bool any = false, any_vil = false;
        double vil = 0.0, prev_liq = 0.0, prev_z = 0.0;
        int top = -1;
        for (int k = 0; k < p.nz; ++k) {
            double v = p.vol[(size_t)k * (size_t)p.ncol + (size_t)c];
            if (p.nan_as_fill && v != v) v = p.fill;
            const bool ok = finite_not_fill(v, p.fill);
            if (ok) { any = true; if (v > best) best = v; }
            if (p.vil) {
                // RadarProduct.py:33-46: 3.44e-6 * (10^(clip(v)/10))^(4/7) where valid & v >= min_dbz, trapezoid
                const bool okv = ok && (v >= p.min_dbz);
                any_vil |= okv;
                double liq = 0.0;
                if (okv) {
                    const double cl = v > p.max_dbz_cap ? p.max_dbz_cap : v;
                    liq = mul(3.44e-6, pow(pow(10.0, div(cl, 10.0)), 4.0 / 7.0));
                }
                const double z = p.levels[k];
                if (k > 0) vil = add(vil, div(mul(sub(z, prev_z), add(liq, prev_liq)), 2.0));
                prev_liq = liq;
                prev_z = z;
            }
            if (ok && v >= p.threshold) top = k;
        }
        if (p.cr) p.cr[c] = any ? best : nan;
        if (p.vil) p.vil[c] = (p.nz >= 2 && any_vil) ? vil : nan;
        if (p.et || p.topped) {
            // RadarProduct.py:49-85
            double et = nan;
            unsigned char tp = 0;
            if (top >= 0) {
                const double z0 = p.levels[top];
                if (top >= p.nz - 1) { et = z0; tp = 1; }
                else {
                    const double z1 = p.levels[top + 1];
                    const double v0 = p.vol[(size_t)top * (size_t)p.ncol + (size_t)c];
                    double v1 = p.vol[(size_t)(top + 1) * (size_t)p.ncol + (size_t)c];
                    if (p.nan_as_fill && v1 != v1) v1 = p.fill;
                    const bool fin1 = (v1 == v1) && (fabs(v1) != inf);
                    if (fin1 && v1 < p.threshold && v0 > p.threshold && v1 != v0)
                        et = add(z0, mul(div(sub(p.threshold, v0), sub(v1, v0)), sub(z1, z0)));
                    else
                        et = z0;
                }
            }
            if (p.et) p.et[c] = et;
            if (p.topped) p.topped[c] = tp;
        }
    }
}

}  // namespace pcg
