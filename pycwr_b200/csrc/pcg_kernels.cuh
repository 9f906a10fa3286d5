// pcg_kernels.cuh — sm_100a kernels of the polar-volume -> Cartesian gridding path.
//
// Data layout in HBM (one pcg_volume): a float64 azimuth arena and a float64 range arena (the
// per-sweep tables back to back), one value arena per moment (sweeps back to back, each sweep a
// contiguous radial x gate tile, float64 or float32), and a SweepDesc table.  Output planes are
// (nz, nx*ny) float64 with ny fastest — the reference's (nz, nx, ny) C order — so a warp that
// owns 32 consecutive columns writes 256 contiguous bytes per level.
//
// Work decomposition: one thread per grid COLUMN (ix, iy).  Everything that depends on (x, y)
// only — ground distance, phi = s/Re, cos(phi), azimuth, the azimuth bracket search inputs —
// is computed once and reused for every level (CAPPI/3-D) or sweep (CR) of the column.
#pragma once
#include <stdint.h>

#include "pcg_math.cuh"

namespace pcg {

constexpr int kMaxSweeps = 128;     // descriptor table staged in shared memory
constexpr int kMaxFields = 4;
constexpr int kIdxStride = 10;

enum State { kSkipBeam = 0, kSkipLow = 1, kSkipHigh = 2, kSingle = 3, kPair = 4, kEmpty = 5 };
enum Flag { kWrapped = 1, kRClamp = 2, kRHigh = 4, kRLow = 8, kDegenerate = 16 };

struct SweepDesc {
    int naz, ng;
    unsigned az_off, rng_off;        // element offsets into the azimuth / range arenas
    unsigned long long val_off;      // element offset into each value arena
    double elev;                     // fixed elevation (deg)
    double tan_e;                    // tan(elev * DEG2RAD), evaluated on the host (glibc)
    double r_first, r_last;          // range[0], range[ng-1]
    double az_first, az_inv_step;    // index guess: floor((az - az_first) * az_inv_step) + 1
    double rng_inv_step;             // index guess: floor((r - r_first) * rng_inv_step) + 1
    int sorted;                      // tables verified ascending on the host at pack time
    int pad;
    double cthr;                     // -sin(elev): the sweep's elevation as a threshold in cosine space
};

struct LevelConst {                  // per output level, evaluated on the host in IEEE double
    double z;
    double g2;                       // g*g,            g = Re + z
    double a2g2;                     // a*a + g*g,      a = Re + h
    double twoag;                    // (2.0*a)*g
};

struct VolumeView {
    const double* az;
    const double* rng;
    const void* val[kMaxFields];
    const SweepDesc* sweeps;
    int ne;
    int nfield;
};

struct GridArgs {
    const double* gx;
    const double* gy;
    long long ncol;
    double radar_x, radar_y;         // local = grid - radar (RadarGridC.pyx:571-572)
    int shift;                       // apply the subtraction (0: use the grid as is)
    int nx, ny;                      // grid shape (ny fastest); r2 kernel: 2-D warp footprints
    int py_log2;                     // a warp covers 2^py_log2 consecutive columns of (32 >> py_log2) rows
    int tiles_y;
};

constexpr int kMaxLevelsPerLaunch = 64;   // level constants travel in the kernel parameters

struct CappiArgs {
    VolumeView vol;
    GridArgs grid;
    int nz;                          // levels in this launch (<= kMaxLevelsPerLaunch)
    unsigned long long field_stride; // elements between the planes of consecutive moments
    double a, a2, twoa, Re, fill;
    double beam_half;                // ne == 1 gate (RadarGridC.pyx:400-404)
    int nearest_gate, lowest_sweep, highest_sweep;
    int merge_max;                   // fill-aware running max into out (RadarGridC.pyx:586-592)
    double* out;                     // [nfield][nz][ncol] (pre-offset to this launch's first level)
    int32_t* idx;                    // optional [nz][ncol][kIdxStride]
    LevelConst levels[kMaxLevelsPerLaunch];
};

struct CrArgs {
    VolumeView vol;
    GridArgs grid;
    double a, a2, twoa, Re, fill;
    int single_sweep;                // >= 0: ppi_to_grid of that sweep, elevation tan_override
    double tan_override;
    int nearest_gate;
    double* out;                     // [ncol]
    int32_t* idx;                    // optional [ncol][kIdxStride] (single_sweep only)
};

template <typename VT>
__device__ __forceinline__ double load_val(const VT* __restrict__ p, size_t i) {
    return (double)__ldg(p + i);
}

struct Bracket {                     // one sweep's interpolation stencil for a (az, r) target
    int iaz, ir, flags;
    int row_lo;
    double azv, az_lo, az_hi;
    double rv, r_lo, r_hi;
    bool ok;
};

// RadarGridC.pyx:264-295 — range gate test, azimuth bracket with wrap, gate bracket.
__device__ __forceinline__ Bracket bracket_sweep(const SweepDesc& d, const double* __restrict__ az_t,
                                                 const double* __restrict__ rg_t, double az,
                                                 double r, bool nearest_gate) {
    Bracket b;
    b.iaz = -1; b.ir = -1; b.flags = 0; b.ok = false; b.row_lo = 0;
    b.azv = az; b.az_lo = 0.0; b.az_hi = 0.0; b.rv = r; b.r_lo = 0.0; b.r_hi = 0.0;
    if (d.naz < 2 || d.ng < 2) { b.flags = kDegenerate; return b; }
    if (r > d.r_last) { b.flags = kRHigh; return b; }
    if (r < d.r_first) {
        if (!nearest_gate) { b.flags = kRLow; return b; }
        b.rv = d.r_first;
        b.flags |= kRClamp;
    }
    const double* A = az_t + d.az_off;
    const double* R = rg_t + d.rng_off;
    const int gaz = __double2int_rd(mul(sub(az, d.az_first), d.az_inv_step)) + 1;
    int iaz = search_right(A, d.naz, az, gaz, d.sorted != 0);
    if (iaz >= d.naz) { iaz = 0; b.azv = sub(az, 360.0); b.flags |= kWrapped; }
    b.az_hi = A[iaz];
    if (iaz == 0) { b.az_lo = sub(A[d.naz - 1], 360.0); b.row_lo = d.naz - 1; }
    else { b.az_lo = A[iaz - 1]; b.row_lo = iaz - 1; }
    const int gr = __double2int_rd(mul(sub(b.rv, d.r_first), d.rng_inv_step)) + 1;
    int ir = search_right(R, d.ng, b.rv, gr, d.sorted != 0);
    if (ir <= 0) ir = 1; else if (ir >= d.ng) ir = d.ng - 1;
    b.r_lo = R[ir - 1];
    b.r_hi = R[ir];
    b.iaz = iaz; b.ir = ir; b.ok = true;
    return b;
}

// RadarGridC.pyx:297-299 — azimuth lerp at both gates, then range lerp (A4 rules each time).
template <typename VT>
__device__ __forceinline__ double sample_field(const SweepDesc& d, const Bracket& b,
                                               const VT* __restrict__ val, double fill) {
    if (!b.ok) return fill;
    const VT* v = val + d.val_off;
    const size_t lo = (size_t)b.row_lo * d.ng + b.ir;
    const size_t hi = (size_t)b.iaz * d.ng + b.ir;
    const double v00 = load_val(v, lo - 1), v01 = load_val(v, lo);
    const double v10 = load_val(v, hi - 1), v11 = load_val(v, hi);
    const double er0 = interp_linear(b.azv, b.az_lo, b.az_hi, v00, v10, fill);
    const double er1 = interp_linear(b.azv, b.az_lo, b.az_hi, v01, v11, fill);
    return interp_linear(b.rv, b.r_lo, b.r_hi, er0, er1, fill);
}

__device__ __forceinline__ void stage_sweeps(SweepDesc* s_sw, const SweepDesc* g_sw, int ne) {
    const int words = ne * (int)(sizeof(SweepDesc) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(g_sw);
    int* dst = reinterpret_cast<int*>(s_sw);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// K_cappi3d — get_CAPPI_xy / get_CAPPI_3d / one radar of get_mosaic_CAPPI_3d
// (RadarGridC.pyx:356-477, 485-521, 568-592).
// ---------------------------------------------------------------------------------------------
template <typename VT, int NF, bool EMIT>
__global__ void __launch_bounds__(256) cappi3d_kernel(const __grid_constant__ CappiArgs p) {
    extern __shared__ unsigned char smem_raw[];
    SweepDesc* s_sw = reinterpret_cast<SweepDesc*>(smem_raw);
    stage_sweeps(s_sw, p.vol.sweeps, p.vol.ne);
    __syncthreads();

    const int ne = p.vol.ne;
    const double fill = p.fill;
    const long long ncol = p.grid.ncol;
    const size_t field_stride = (size_t)p.field_stride;

    for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < ncol;
         col += (long long)gridDim.x * blockDim.x) {
        double x = p.grid.gx[col], y = p.grid.gy[col];
        if (p.grid.shift) { x = sub(x, p.grid.radar_x); y = sub(y, p.grid.radar_y); }
        // column invariants (RadarGridC.pyx:142-146,158)
        const double s = sqrt_rn(add(mul(x, x), mul(y, y)));
        const double phi = div(s, p.Re);
        const double cphi = cos_small(phi);
        const double az = azimuth_from_xy(x, y);

        for (int iz = 0; iz < p.nz; ++iz) {
            const LevelConst& lc = p.levels[iz];
            const size_t o = (size_t)iz * (size_t)ncol + (size_t)col;
            double res[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) res[f] = fill;
            int state = kEmpty, lower = -1, upper = -1;
            Bracket b0, b1;
            b0.iaz = b0.ir = -1; b0.flags = 0; b0.ok = false;
            b1.iaz = b1.ir = -1; b1.flags = 0; b1.ok = false;
            bool have0 = false, have1 = false;

            if (ne > 0) {
                // RadarGridC.pyx:146-157
                double r2 = sub(lc.a2g2, mul(lc.twoag, cphi));
                if (r2 < 0.0) r2 = 0.0;
                const double r = sqrt_rn(r2);
                double el;
                if (r == 0.0) {
                    el = 0.0;
                } else {
                    double c = div(sub(add(p.a2, r2), lc.g2), mul(p.twoa, r));
                    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
                    el = mul(sub(acos(c), kHalfPi), kRad2Deg);
                }
                bool skip = false;
                if (ne == 1) {
                    // RadarGridC.pyx:408-423
                    if (fabs(sub(el, s_sw[0].elev)) > p.beam_half) { state = kSkipBeam; skip = true; }
                    else { lower = upper = 0; }
                } else if (el < s_sw[0].elev) {
                    if (!p.lowest_sweep) { state = kSkipLow; skip = true; } else { lower = upper = 0; }
                } else if (el > s_sw[ne - 1].elev) {
                    if (!p.highest_sweep) { state = kSkipHigh; skip = true; } else { lower = upper = ne - 1; }
                } else {
                    // bisect_right over the fixed elevations (RadarGridC.pyx:434-439)
                    int lo = 0, hi = ne;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (el < s_sw[mid].elev) hi = mid; else lo = mid + 1;
                    }
                    if (lo >= ne) lo = ne - 1;
                    lower = lo - 1; upper = lo;
                }
                if (!skip) {
                    state = (lower == upper) ? kSingle : kPair;
                    const SweepDesc& d0 = s_sw[lower];
                    b0 = bracket_sweep(d0, p.vol.az, p.vol.rng, az, r, p.nearest_gate != 0);
                    have0 = true;
                    if (lower == upper) {
#pragma unroll
                        for (int f = 0; f < NF; ++f)
                            res[f] = sample_field(d0, b0, static_cast<const VT*>(p.vol.val[f]), fill);
                    } else {
                        const SweepDesc& d1 = s_sw[upper];
                        b1 = bracket_sweep(d1, p.vol.az, p.vol.rng, az, r, p.nearest_gate != 0);
                        have1 = true;
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const double v0 = sample_field(d0, b0, static_cast<const VT*>(p.vol.val[f]), fill);
                            const double v1 = sample_field(d1, b1, static_cast<const VT*>(p.vol.val[f]), fill);
                            res[f] = interp_linear(el, d0.elev, d1.elev, v0, v1, fill);
                        }
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                double* dst = p.out + (size_t)f * field_stride + o;
                if (p.merge_max) {
                    const double cur = *dst;
                    if (res[f] == fill) continue;
                    if (cur == fill || res[f] > cur) *dst = res[f];
                } else {
                    *dst = res[f];
                }
            }
            if (EMIT) {
                int32_t* rec = p.idx + o * kIdxStride;
                rec[0] = state; rec[1] = lower; rec[2] = upper;
                rec[3] = have0 ? b0.iaz : -1; rec[4] = have0 ? b0.ir : -1; rec[5] = have0 ? b0.flags : -1;
                rec[6] = have1 ? b1.iaz : -1; rec[7] = have1 ? b1.ir : -1; rec[8] = have1 ? b1.flags : -1;
                rec[9] = -1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K_cr — get_CR_xy (RadarGridC.pyx:596-624) and ppi_to_grid (RadarGridC.pyx:303-352).
// One pass: every sweep gridded at its fixed elevation, fill/NaN-aware max kept in a register.
// ---------------------------------------------------------------------------------------------
template <typename VT, bool EMIT>
__global__ void __launch_bounds__(256) cr_kernel(const CrArgs p) {
    extern __shared__ unsigned char smem_raw[];
    SweepDesc* s_sw = reinterpret_cast<SweepDesc*>(smem_raw);
    stage_sweeps(s_sw, p.vol.sweeps, p.vol.ne);
    __syncthreads();

    const double fill = p.fill;
    const long long ncol = p.grid.ncol;
    const VT* val = static_cast<const VT*>(p.vol.val[0]);
    const int e_begin = p.single_sweep >= 0 ? p.single_sweep : 0;
    const int e_end = p.single_sweep >= 0 ? p.single_sweep + 1 : p.vol.ne;

    for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < ncol;
         col += (long long)gridDim.x * blockDim.x) {
        double x = p.grid.gx[col], y = p.grid.gy[col];
        if (p.grid.shift) { x = sub(x, p.grid.radar_x); y = sub(y, p.grid.radar_y); }
        // RadarGridC.pyx:108-118
        const double s = sqrt_rn(add(mul(x, x), mul(y, y)));
        const double phi = div(s, p.Re);
        const double sphi = sin_small(phi);
        const double cphi = cos_small(phi);
        const double az = azimuth_from_xy(x, y);

        double best = 0.0;
        bool any = false;
        double single = fill;
        for (int ie = e_begin; ie < e_end; ++ie) {
            const SweepDesc& d = s_sw[ie];
            if (d.naz < 2 || d.ng < 2) {
                if (EMIT) {
                    int32_t* rec = p.idx + (size_t)col * kIdxStride;
                    rec[0] = kSingle; rec[1] = 0; rec[2] = 0; rec[3] = -1; rec[4] = -1; rec[5] = kDegenerate;
                    rec[6] = rec[7] = rec[8] = rec[9] = -1;
                }
                continue;
            }
            const double tan_e = p.single_sweep >= 0 ? p.tan_override : d.tan_e;
            // RadarGridC.pyx:113-129
            const double denom = sub(cphi, mul(sphi, tan_e));
            double r;
            if (denom == 0.0) {
                r = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                const double g = div(p.a, denom);
                double r2 = sub(add(p.a2, mul(g, g)), mul(mul(p.twoa, g), cphi));
                if (r2 < 0.0) r2 = 0.0;
                r = sqrt_rn(r2);
            }
            const Bracket b = bracket_sweep(d, p.vol.az, p.vol.rng, az, r, p.nearest_gate != 0);
            const double v = sample_field(d, b, val, fill);
            if (EMIT) {
                int32_t* rec = p.idx + (size_t)col * kIdxStride;
                rec[0] = kSingle; rec[1] = 0; rec[2] = 0; rec[3] = b.iaz; rec[4] = b.ir; rec[5] = b.flags;
                rec[6] = rec[7] = rec[8] = rec[9] = -1;
            }
            single = v;
            // RadarGridC.pyx:620-623 — fill -> NaN, NaN dropped, max of the rest
            if (v == fill || v != v) continue;
            if (!any || v > best) best = v;
            any = true;
        }
        p.out[col] = (p.single_sweep >= 0) ? single : (any ? best : fill);
    }
}

// ---------------------------------------------------------------------------------------------
// Diagnostics: the device geometry for arbitrary points (tests measure agreement with glibc).
// ---------------------------------------------------------------------------------------------
__global__ void debug_c2a_kernel(const double* x, const double* y, size_t n, LevelConst lc,
                                 double a2, double twoa, double Re, double* az, double* r,
                                 double* el) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double s = sqrt_rn(add(mul(x[i], x[i]), mul(y[i], y[i])));
        const double cphi = cos_small(div(s, Re));
        double r2 = sub(lc.a2g2, mul(lc.twoag, cphi));
        if (r2 < 0.0) r2 = 0.0;
        const double rr = sqrt_rn(r2);
        double e = 0.0;
        if (rr != 0.0) {
            double c = div(sub(add(a2, r2), lc.g2), mul(twoa, rr));
            c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
            e = mul(sub(acos(c), kHalfPi), kRad2Deg);
        }
        az[i] = azimuth_from_xy(x[i], y[i]);
        r[i] = rr;
        el[i] = e;
    }
}

__global__ void debug_xye_kernel(const double* x, const double* y, size_t n, double tan_e,
                                 double a, double a2, double twoa, double Re, double* az,
                                 double* r) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double s = sqrt_rn(add(mul(x[i], x[i]), mul(y[i], y[i])));
        const double phi = div(s, Re);
        const double sphi = sin_small(phi), cphi = cos_small(phi);
        const double denom = sub(cphi, mul(sphi, tan_e));
        double rr;
        if (denom == 0.0) {
            rr = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const double g = div(a, denom);
            double r2 = sub(add(a2, mul(g, g)), mul(mul(twoa, g), cphi));
            if (r2 < 0.0) r2 = 0.0;
            rr = sqrt_rn(r2);
        }
        az[i] = azimuth_from_xy(x[i], y[i]);
        r[i] = rr;
    }
}

}  // namespace pcg
