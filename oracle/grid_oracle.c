/*
 * oracle/grid_oracle.c — CPU restatement of pycwr's polar-volume -> Cartesian
 * gridding kernels, with the interpolation indices made visible.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pycwr_b200/ may link, import or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py have something to compare the CUDA path with.
 *
 * It restates (does not copy) the algorithm of the reference file
 * pycwr/core/RadarGridC.pyx; every function cites the lines it follows.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile) so the
 * arithmetic is unfused IEEE double exactly like the reference's own build
 * (x86-64 baseline, no FMA), and links the same glibc libm.
 *
 * Pinning: tests/test_oracle_pinning.py checks this file bit-for-bit on
 * values against the compiled reference (oracle/_ref, built from the
 * reference sources where they lie) and against the committed fixtures in
 * tests/golden/ (generated from that compiled reference), and against the
 * known answers of the reference's own tests (15.0 / 35.0 / 10.0 / -999->10).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* RadarGridC.pyx:6-9 */
static const double OG_PI = 3.141592653589793;
#define OG_DEG2RAD (OG_PI / 180.0)
#define OG_RAD2DEG (180.0 / OG_PI)

/* one interpolation record per output voxel (int32 x OG_IDX_STRIDE) */
#define OG_IDX_STRIDE 10
enum { OG_SKIP_BEAM = 0, OG_SKIP_LOW = 1, OG_SKIP_HIGH = 2, OG_SINGLE = 3, OG_PAIR = 4,
       OG_EMPTY = 5 };
enum { OG_F_WRAPPED = 1, OG_F_RCLAMP = 2, OG_F_RHIGH = 4, OG_F_RLOW = 8, OG_F_DEGEN = 16 };

typedef struct {
    int ne;
    const int *naz;            /* [ne] rays per sweep   */
    const int *ng;             /* [ne] gates per sweep  */
    const double *const *az;   /* [ne] -> [naz] sorted azimuth (deg)           */
    const double *const *rng;  /* [ne] -> [ng]  sorted slant range (m)         */
    const double *const *val;  /* [ne] -> [naz*ng] row-major (radial, gate)    */
    const double *fix_elev;    /* [ne] ascending fixed elevation (deg)         */
} og_volume;

/* RadarGridC.pyx:22-27 */
static double og_clip_unit(double v) {
    if (v > 1.0) return 1.0;
    if (v < -1.0) return -1.0;
    return v;
}

/* RadarGridC.pyx:34-38 */
double og_xy_to_azimuth(double x, double y) {
    double az = OG_PI / 2.0 - atan2(y, x);
    if (az < 0.0) az += 2.0 * OG_PI;
    return az * OG_RAD2DEG;
}

/* RadarGridC.pyx:41-51 — bisect_right */
static int og_search_right(const double *v, int n, double target) {
    ptrdiff_t lo = 0, hi = n, mid;
    while (lo < hi) {
        mid = (lo + hi) / 2;
        if (target < v[mid]) hi = mid; else lo = mid + 1;
    }
    return (int)lo;
}

/* RadarGridC.pyx:72-95 */
void og_antenna_to_cartesian(double ranges, double azimuth, double elevation, double h, double Re,
                             double *x, double *y, double *z) {
    double theta_a = azimuth * OG_DEG2RAD;
    double theta_e = elevation * OG_DEG2RAD;
    double sin_e = sin(theta_e), cos_e = cos(theta_e);
    double a = Re + h;
    double arc, s;
    *z = sqrt(ranges * ranges + a * a + 2.0 * ranges * a * sin_e) - Re;
    arc = og_clip_unit(ranges * cos_e / (Re + *z));
    s = Re * asin(arc);
    *x = s * sin(theta_a);
    *y = s * cos(theta_a);
}

/* RadarGridC.pyx:98-129 */
void og_xye_to_antenna(double x, double y, double elevation, double h, double Re,
                       double *azimuth, double *ranges, double *z) {
    double s = sqrt(x * x + y * y);
    double phi = s / Re;
    double theta_e = elevation * OG_DEG2RAD;
    double sin_phi = sin(phi), cos_phi = cos(phi);
    double denom = cos_phi - sin_phi * tan(theta_e);
    double a = Re + h, g, r2;
    *azimuth = og_xy_to_azimuth(x, y);
    if (denom == 0.0) { *ranges = NAN; *z = NAN; return; }
    g = a / denom;
    *z = g - Re;
    r2 = a * a + g * g - 2.0 * a * g * cos_phi;
    if (r2 < 0.0) r2 = 0.0;
    *ranges = sqrt(r2);
}

/* RadarGridC.pyx:132-158 */
void og_cartesian_to_antenna(double x, double y, double z, double h, double Re,
                             double *azimuth, double *ranges, double *elevation) {
    double s = sqrt(x * x + y * y);
    double phi = s / Re;
    double a = Re + h, g = Re + z;
    double r2 = a * a + g * g - 2.0 * a * g * cos(phi);
    double c;
    if (r2 < 0.0) r2 = 0.0;
    *ranges = sqrt(r2);
    if (*ranges == 0.0) {
        *elevation = 0.0;
    } else {
        c = og_clip_unit((a * a + r2 - g * g) / (2.0 * a * *ranges));
        *elevation = (acos(c) - OG_PI / 2.0) * OG_RAD2DEG;
    }
    *azimuth = og_xy_to_azimuth(x, y);
}

/* RadarGridC.pyx:236-252 (public name interp_azimuth, :479-481) */
double og_interp_linear(double c, double c0, double c1, double d0, double d1, double fill) {
    if (c1 == c0) return fill;
    if (d0 == fill && d1 == fill) return fill;
    if (d0 == fill) return d1;
    if (d1 == fill) return d0;
    return ((c1 - c) * d0 + (c - c0) * d1) / (c1 - c0);
}

/* RadarGridC.pyx:200-233 */
double og_interp_ppi(double az, double r, double az_0, double az_1, double r_0, double r_1,
                     double m00, double m01, double m10, double m11, double fill) {
    double az_span = az_1 - az_0, r_span = r_1 - r_0;
    if (az_span == 0.0 || r_span == 0.0) return fill;
    if (m00 != fill && m01 != fill && m10 != fill && m11 != fill)
        return (m00 * (az_1 - az) * (r_1 - r) + m10 * (az - az_0) * (r_1 - r) +
                m01 * (az_1 - az) * (r - r_0) + m11 * (az - az_0) * (r - r_0)) / r_span / az_span;
    if (m00 != fill && m01 != fill) return (m00 * (r_1 - r) + m01 * (r - r_0)) / r_span;
    if (m10 != fill && m11 != fill) return (m10 * (r_1 - r) + m11 * (r - r_0)) / r_span;
    if (m00 != fill && m10 != fill) return (m00 * (az_1 - az) + m10 * (az - az_0)) / az_span;
    if (m01 != fill && m11 != fill) return (m01 * (az_1 - az) + m11 * (az - az_0)) / az_span;
    return fill;
}

/* RadarGridC.pyx:255-299.  idx (optional) receives {iaz, ir, flags}. */
static double og_sweep_sample(const double *az_t, int naz, const double *rg_t, int ng,
                              const double *val, double az, double r, double fill,
                              int use_nearest_gate, int32_t *idx) {
    int iaz, ir, flags = 0, row_lo;
    double azv = az, az_last, rv = r, er0, er1;
    if (idx) { idx[0] = -1; idx[1] = -1; idx[2] = 0; }
    if (naz < 2 || ng < 2) { if (idx) idx[2] = OG_F_DEGEN; return fill; }
    if (rv > rg_t[ng - 1]) { if (idx) idx[2] = OG_F_RHIGH; return fill; }
    if (rv < rg_t[0]) {
        if (!use_nearest_gate) { if (idx) idx[2] = OG_F_RLOW; return fill; }
        rv = rg_t[0];
        flags |= OG_F_RCLAMP;
    }
    iaz = og_search_right(az_t, naz, azv);
    if (iaz >= naz) { iaz = 0; azv -= 360.0; flags |= OG_F_WRAPPED; }
    if (iaz == 0) az_last = az_t[naz - 1] - 360.0; else az_last = az_t[iaz - 1];
    ir = og_search_right(rg_t, ng, rv);
    if (ir <= 0) ir = 1; else if (ir >= ng) ir = ng - 1;
    /* value_view[iaz - 1] with iaz == 0 wraps to the last radial (:297-298) */
    row_lo = (iaz == 0) ? naz - 1 : iaz - 1;
    er0 = og_interp_linear(azv, az_last, az_t[iaz], val[(size_t)row_lo * ng + ir - 1],
                           val[(size_t)iaz * ng + ir - 1], fill);
    er1 = og_interp_linear(azv, az_last, az_t[iaz], val[(size_t)row_lo * ng + ir],
                           val[(size_t)iaz * ng + ir], fill);
    if (idx) { idx[0] = iaz; idx[1] = ir; idx[2] = flags; }
    return og_interp_linear(rv, rg_t[ir - 1], rg_t[ir], er0, er1, fill);
}

/* RadarGridC.pyx:303-352 — one sweep on a plane.  out[ncol]; idx optional [ncol][OG_IDX_STRIDE]. */
void og_ppi_to_grid(const double *az_t, int naz, const double *rg_t, int ng, double elevation,
                    const double *val, double h, double Re, const double *gx, const double *gy,
                    long ncol, double fill, int blind_code, double *out, int32_t *idx) {
    int use_nearest_gate = (blind_code == 1 || blind_code == 3 || blind_code == 4);
    long i;
    for (i = 0; i < ncol; ++i) {
        int32_t *rec = idx ? idx + (size_t)i * OG_IDX_STRIDE : NULL;
        double az, r, z;
        if (rec) { int k; for (k = 0; k < OG_IDX_STRIDE; ++k) rec[k] = -1; rec[0] = OG_SINGLE; rec[1] = rec[2] = 0; }
        if (naz < 2 || ng < 2) { out[i] = fill; if (rec) { rec[5] = OG_F_DEGEN; } continue; }
        og_xye_to_antenna(gx[i], gy[i], elevation, h, Re, &az, &r, &z);
        out[i] = og_sweep_sample(az_t, naz, rg_t, ng, val, az, r, fill, use_nearest_gate,
                                 rec ? rec + 3 : NULL);
    }
}

/* RadarGridC.pyx:596-624 — CR = max over sweeps of the fixed-elevation planes, fill ignored. */
void og_get_cr(const og_volume *v, double h, double Re, const double *gx, const double *gy,
               long ncol, double fill, double *out) {
    long i; int ie;
    for (i = 0; i < ncol; ++i) {
        double best = -INFINITY; int any = 0;
        for (ie = 0; ie < v->ne; ++ie) {
            double az, r, z, s;
            if (v->naz[ie] < 2 || v->ng[ie] < 2) continue;
            og_xye_to_antenna(gx[i], gy[i], v->fix_elev[ie], h, Re, &az, &r, &z);
            s = og_sweep_sample(v->az[ie], v->naz[ie], v->rng[ie], v->ng[ie], v->val[ie], az, r,
                                fill, 0, NULL);
            /* :620-623  fill -> NaN, NaN excluded, max over the rest (NaN samples are dropped
               by ~isnan as well) */
            if (s == fill || s != s) continue;
            if (!any || s > best) best = s;
            any = 1;
        }
        out[i] = any ? best : fill;
    }
}

/* RadarGridC.pyx:356-477 (one level) + :485-521 (levels loop; beam width is NOT forwarded
 * by get_CAPPI_3d, the caller passes 1.0 for that entry point). out[nz][ncol]. */
void og_get_cappi3d(const og_volume *v, double h, double Re, const double *gx, const double *gy,
                    long ncol, const double *levels, int nz, double fill, int blind_code,
                    double beam_width_deg, double *out, int32_t *idx) {
    int use_nearest_gate = (blind_code == 1 || blind_code == 3 || blind_code == 4);
    int use_lowest = (blind_code == 2 || blind_code == 3 || blind_code == 4);
    int use_highest = (blind_code == 4);
    int ne = v->ne, iz;
    long i;
    double beam_half;
    if (beam_width_deg <= 0.0) beam_width_deg = 1.0;
    beam_half = beam_width_deg * 0.5;
    if (beam_half < 0.1) beam_half = 0.1;
    for (iz = 0; iz < nz; ++iz) {
        for (i = 0; i < ncol; ++i) {
            size_t o = (size_t)iz * ncol + i;
            int32_t *rec = idx ? idx + o * OG_IDX_STRIDE : NULL;
            double az, r, el, v0, v1;
            int lo, hi, ie;
            out[o] = fill;
            if (rec) { int k; for (k = 0; k < OG_IDX_STRIDE; ++k) rec[k] = -1; rec[0] = OG_EMPTY; }
            if (ne == 0) continue;
            og_cartesian_to_antenna(gx[i], gy[i], levels[iz], h, Re, &az, &r, &el);
            if (ne == 1) {
                if (fabs(el - v->fix_elev[0]) > beam_half) { if (rec) rec[0] = OG_SKIP_BEAM; continue; }
                if (rec) { rec[0] = OG_SINGLE; rec[1] = rec[2] = 0; }
                out[o] = og_sweep_sample(v->az[0], v->naz[0], v->rng[0], v->ng[0], v->val[0], az, r,
                                         fill, use_nearest_gate, rec ? rec + 3 : NULL);
                continue;
            }
            if (el < v->fix_elev[0]) {
                if (!use_lowest) { if (rec) rec[0] = OG_SKIP_LOW; continue; }
                lo = hi = 0;
            } else if (el > v->fix_elev[ne - 1]) {
                if (!use_highest) { if (rec) rec[0] = OG_SKIP_HIGH; continue; }
                lo = hi = ne - 1;
            } else {
                ie = og_search_right(v->fix_elev, ne, el);
                if (ie >= ne) ie = ne - 1;
                lo = ie - 1; hi = ie;
            }
            if (rec) { rec[0] = (lo == hi) ? OG_SINGLE : OG_PAIR; rec[1] = lo; rec[2] = hi; }
            v0 = og_sweep_sample(v->az[lo], v->naz[lo], v->rng[lo], v->ng[lo], v->val[lo], az, r,
                                 fill, use_nearest_gate, rec ? rec + 3 : NULL);
            if (lo == hi) { out[o] = v0; continue; }
            v1 = og_sweep_sample(v->az[hi], v->naz[hi], v->rng[hi], v->ng[hi], v->val[hi], az, r,
                                 fill, use_nearest_gate, rec ? rec + 6 : NULL);
            out[o] = og_interp_linear(el, v->fix_elev[lo], v->fix_elev[hi], v0, v1, fill);
        }
    }
}

/* RadarGridC.pyx:525-593 — per radar: shift grid, CAPPI_3d (blind "mask", beam width 1.0),
 * fill-aware running max.  scratch = 2*ncol + nz*ncol doubles.  Re[nradar]. */
void og_get_mosaic_cappi3d(int nradar, const og_volume *vols, const double *radar_x,
                           const double *radar_y, const double *radar_h, const double *Re,
                           const double *gx, const double *gy, long ncol, const double *levels,
                           int nz, double fill, double *out, double *scratch) {
    double *lx = scratch, *ly = scratch + ncol, *rv = scratch + 2 * ncol;
    size_t n = (size_t)nz * ncol, k;
    int ir; long i;
    for (k = 0; k < n; ++k) out[k] = fill;
    for (ir = 0; ir < nradar; ++ir) {
        for (i = 0; i < ncol; ++i) { lx[i] = gx[i] - radar_x[ir]; ly[i] = gy[i] - radar_y[ir]; }
        og_get_cappi3d(&vols[ir], radar_h[ir], Re[ir], lx, ly, ncol, levels, nz, fill, 0, 1.0, rv, NULL);
        for (k = 0; k < n; ++k) {
            if (rv[k] == fill) continue;
            if (out[k] == fill || rv[k] > out[k]) out[k] = rv[k];
        }
    }
}

/* Bulk geometry (used to pin device libm against glibc): per point az, r, el for one level. */
void og_bulk_cartesian_to_antenna(const double *x, const double *y, long n, double z, double h,
                                  double Re, double *az, double *r, double *el) {
    long i;
    for (i = 0; i < n; ++i) og_cartesian_to_antenna(x[i], y[i], z, h, Re, az + i, r + i, el + i);
}

void og_bulk_xye_to_antenna(const double *x, const double *y, long n, double elev, double h,
                            double Re, double *az, double *r, double *z) {
    long i;
    for (i = 0; i < n; ++i) og_xye_to_antenna(x[i], y[i], elev, h, Re, az + i, r + i, z + i);
}
