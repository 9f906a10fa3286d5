// pcg_wind.cuh — the gridding step of retrieve_wind_volume_xy (pycwr/retrieve/WindField.py):
//   wind_gather_kernel    _grid_vvp_retrieval_xy (:629-689) + _aggregate_neighbors (:597-626): every VVP
//                         sample within horizontal_radius_m of a target (once expanded to
//                         max_horizontal_radius_m when fewer than min_neighbors are found), weights
//                         max(count,1) * clip(fraction,0,1) / max(rmse,0.5) / max(d,1)^power
//   wind_vertical_kernel  _vertical_profile_from_sweeps (:692-810) + _compute_voxel_quality (:813-851)
//                         for every column of the target grid (_reconstruct_wind_volume_chunk, :1183-1229)
// The reference walks a cKDTree per target and a Python loop per column; here one thread owns one
// target / column and the samples stream through shared memory in fixed index order (deterministic
// sums).  Both radii are accumulated in the same pass, so the expansion costs no second search.
// Everything is FP64 like the reference: these fields are velocities (1e-5 relative bar).
#pragma once
#include "pcg_math.cuh"

namespace pcg {

constexpr int kWindTile = 256;      // samples staged per shared-memory tile
constexpr int kWindFields = 10;     // x, y, z, u, v, fit_rmse, valid_count, valid_fraction, condition_number, azimuth_sector_count
constexpr int kWindMaxSweeps = 64;

struct WindGatherArgs {
    const double* samples;          // [kWindFields][ns] structure of arrays, already filtered to finite samples
    long long ns;
    const double* tx;               // [nt]
    const double* ty;
    long long nt;
    double r1sq, r2sq;              // squared search radii (r2sq <= r1sq: no expansion)
    int min_neighbors;
    double power;
    double* u; double* v; double* z; double* rmse; double* cond;      // [nt] (NaN where undefined)
    int* valid_count; int* neighbor_count; int* azsec;                // [nt]
    unsigned char* expanded;                                          // [nt]
};

struct WindAcc {
    double w, u, v, z, rmse, vc, cond, az;
    int n;
    bool any;
};

__device__ __forceinline__ void wind_acc_init(WindAcc& a) {
    a.w = a.u = a.v = a.z = a.rmse = a.vc = a.cond = a.az = 0.0;
    a.n = 0;
    a.any = false;
}

__global__ void __launch_bounds__(kWindTile) wind_gather_kernel(const __grid_constant__ WindGatherArgs p) {
    __shared__ double s[kWindFields][kWindTile];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < p.nt;
    const double x0 = active ? p.tx[t] : 0.0, y0 = active ? p.ty[t] : 0.0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const bool expand = p.r2sq > p.r1sq;
    WindAcc a1, a2;
    wind_acc_init(a1);
    wind_acc_init(a2);
    for (long long base = 0; base < p.ns; base += kWindTile) {
        const int cnt = (int)min((long long)kWindTile, p.ns - base);
        __syncthreads();
        if ((int)threadIdx.x < cnt)
            for (int f = 0; f < kWindFields; ++f) s[f][threadIdx.x] = p.samples[(size_t)f * p.ns + base + threadIdx.x];
        __syncthreads();
        if (!active) continue;
        for (int j = 0; j < cnt; ++j) {
            const double dx = sub(s[0][j], x0), dy = sub(s[1][j], y0);
            const double d2 = add(mul(dx, dx), mul(dy, dy));
            const bool in1 = d2 <= p.r1sq;
            const bool in2 = expand && (d2 <= p.r2sq);
            if (!(in1 | in2)) continue;
            // WindField.py:605-614
            const double vc = s[6][j], vf = s[7][j], rm = s[5][j];
            const double q = div(mul(fmax(vc, 1.0), fmin(fmax(vf, 0.0), 1.0)), fmax(rm, 0.5));
            const double d = fmax(sqrt_rn(d2), 1.0);
            double w = div(q, (p.power == 2.0) ? mul(d, d) : pow(d, p.power));
            if (!((w == w) && (fabs(w) != inf) && (w > 0.0))) w = 0.0;
#pragma unroll
            for (int set = 0; set < 2; ++set) {
                if (!(set == 0 ? in1 : in2)) continue;
                WindAcc& a = set == 0 ? a1 : a2;
                a.n += 1;
                a.any |= w > 0.0;
                a.w = add(a.w, w);
                a.u = add(a.u, mul(s[3][j], w));
                a.v = add(a.v, mul(s[4][j], w));
                a.z = add(a.z, mul(s[2][j], w));
                a.rmse = add(a.rmse, mul(rm, w));
                a.vc = add(a.vc, mul(vc, w));
                a.cond = add(a.cond, mul(s[8][j], w));
                a.az = add(a.az, mul(s[9][j], w));
            }
        }
    }
    if (!active) return;
    // WindField.py:666-688
    bool expanded = false;
    const WindAcc* a = &a1;
    if (a1.n < p.min_neighbors && expand) { a = &a2; expanded = true; }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double u = nan, v = nan, z = nan, rm = nan, cd = nan;
    int vcount = 0, ncount = 0, az = 0;
    unsigned char ex = 0;
    if (a->n >= p.min_neighbors && a->any) {
        u = div(a->u, a->w); v = div(a->v, a->w); z = div(a->z, a->w);
        rm = div(a->rmse, a->w); cd = div(a->cond, a->w);
        vcount = (int)rint(div(a->vc, a->w));            // np.round: half to even
        az = (int)rint(div(a->az, a->w));
        ncount = a->n;
        ex = expanded ? 1 : 0;
    }
    p.u[t] = u; p.v[t] = v; p.z[t] = z; p.rmse[t] = rm; p.cond[t] = cd;
    p.valid_count[t] = vcount; p.neighbor_count[t] = ncount; p.azsec[t] = az;
    p.expanded[t] = ex;
}

struct WindVerticalArgs {
    const double* sz; const double* su; const double* sv; const double* srmse;   // [nsweep][ncol]
    const double* scount; const double* sneigh; const double* sexp;
    int nsweep;
    long long ncol;
    const double* levels;           // [nz] (device), strictly increasing
    int nz;
    double tol, max_gap;
    double* u; double* v; double* rmse; double* spread; double* score;            // [nz][ncol]
    int* valid_count; int* source_count; int* neighbor_count; int* support;
    unsigned char* flag;
};

// WindField.py:813-851
__device__ __forceinline__ void voxel_quality(int source, int valid, double rmse, int neighbor, double spread,
                                              double max_gap, bool expanded, double& score, unsigned char& flag) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    score = 0.0; flag = 0;
    if (source <= 0 || valid <= 0 || !((rmse == rmse) && fabs(rmse) != inf)) return;
    const double support_c = fmin(1.0, div((double)source, 3.0));
    const double sample_c = fmin(1.0, div((double)valid, 60.0));
    const double neigh_c = fmin(1.0, div((double)max(neighbor, 0), 6.0));
    const double rmse_c = exp(div(-fmax(rmse, 0.0), 6.0));
    double gap_c = 1.0;
    if ((spread == spread) && fabs(spread) != inf && max_gap > 0.0) gap_c = fmax(0.0, sub(1.0, div(spread, max_gap)));
    double sc = mul(0.30, support_c);
    sc = add(sc, mul(0.20, sample_c));
    sc = add(sc, mul(0.20, neigh_c));
    sc = add(sc, mul(0.20, rmse_c));
    sc = add(sc, mul(0.10, gap_c));
    sc = mul(100.0, sc);
    if (expanded) sc = mul(sc, 0.9);
    sc = fmin(fmax(sc, 0.0), 100.0);
    score = sc;
    flag = sc <= 0.0 ? 0 : (sc < 40.0 ? 1 : (sc < 70.0 ? 2 : 3));
}

__global__ void __launch_bounds__(128) wind_vertical_kernel(const __grid_constant__ WindVerticalArgs p) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.ncol) return;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    // valid sweeps of this column, sorted by height (WindField.py:722-747; insertion sort is stable,
    // equal heights keep sweep order)
    double h[kWindMaxSweeps];
    unsigned char id[kWindMaxSweeps];
    int n = 0;
    for (int s = 0; s < p.nsweep; ++s) {
        const size_t o = (size_t)s * (size_t)p.ncol + (size_t)c;
        const double z = p.sz[o], cu = p.su[o], cv = p.sv[o], cr = p.srmse[o], cc = p.scount[o], cn = p.sneigh[o],
                     ce = p.sexp[o];
        auto fin = [inf](double t) { return (t == t) && fabs(t) != inf; };
        if (!(fin(z) && fin(cu) && fin(cv) && fin(cr) && fin(cc) && fin(cn) && fin(ce) && cc > 0.0)) continue;
        int j = n++;
        while (j > 0 && h[j - 1] > z) { h[j] = h[j - 1]; id[j] = id[j - 1]; --j; }
        h[j] = z; id[j] = (unsigned char)s;
    }
    for (int iz = 0; iz < p.nz; ++iz) {
        const double lev = p.levels[iz];
        const size_t o = (size_t)iz * (size_t)p.ncol + (size_t)c;
        double ou = nan, ov = nan, orm = nan, ospread = nan, oscore = 0.0;
        int ovc = 0, osrc = 0, onb = 0, osup = 0;
        unsigned char oflag = 0;
        // ---- samples within the vertical tolerance (WindField.py:751-781)
        double sw = 0.0, suu = 0.0, svv = 0.0, srm = 0.0, scnt = 0.0, snb = 0.0, hmin = inf, hmax = -inf;
        int near = 0;
        bool anyw = false, expanded = false;
        for (int j = 0; j < n; ++j) {
            const double dh = fabs(sub(h[j], lev));
            if (!(dh <= p.tol)) continue;
            const size_t so = (size_t)id[j] * (size_t)p.ncol + (size_t)c;
            const double cnt = p.scount[so], rm = p.srmse[so];
            double w = div(cnt, mul(fmax(rm, 0.5), fmax(dh, 1.0)));
            if (!((w == w) && fabs(w) != inf && w > 0.0)) w = 0.0;
            anyw |= w > 0.0;
            ++near;
            hmin = fmin(hmin, h[j]); hmax = fmax(hmax, h[j]);
            expanded |= p.sexp[so] > 0.5;
            sw = add(sw, w);
            suu = add(suu, mul(p.su[so], w));
            svv = add(svv, mul(p.sv[so], w));
            srm = add(srm, mul(rm, w));
            scnt = add(scnt, mul(cnt, w));
            snb = add(snb, mul(p.sneigh[so], w));
        }
        bool done = false;
        if (near > 0 && anyw) {
            ospread = near > 1 ? sub(hmax, hmin) : 0.0;
            ou = div(suu, sw); ov = div(svv, sw); orm = div(srm, sw);
            ovc = (int)rint(div(scnt, sw));
            onb = (int)rint(div(snb, sw));
            osrc = near; osup = near;
            voxel_quality(osrc, ovc, orm, onb, ospread, p.max_gap, expanded, oscore, oflag);
            done = true;
        }
        if (!done) {
            // ---- two-point interpolation across the gap (WindField.py:782-808)
            int lower = -1, upper = -1;
            for (int j = 0; j < n; ++j) {
                if (h[j] < lev) lower = j;
                if (upper < 0 && h[j] > lev) upper = j;
            }
            if (lower >= 0 && upper >= 0) {
                const double span = sub(h[upper], h[lower]);
                if ((span == span) && fabs(span) != inf && span > 0.0 && !(span > p.max_gap)) {
                    const size_t lo = (size_t)id[lower] * (size_t)p.ncol + (size_t)c;
                    const size_t up = (size_t)id[upper] * (size_t)p.ncol + (size_t)c;
                    const double f = div(sub(lev, h[lower]), span), g = sub(1.0, f);
                    ou = add(mul(g, p.su[lo]), mul(f, p.su[up]));
                    ov = add(mul(g, p.sv[lo]), mul(f, p.sv[up]));
                    orm = add(mul(g, p.srmse[lo]), mul(f, p.srmse[up]));
                    ovc = (int)rint(add(mul(g, p.scount[lo]), mul(f, p.scount[up])));
                    onb = (int)rint(add(mul(g, p.sneigh[lo]), mul(f, p.sneigh[up])));
                    osrc = 2; osup = 2;
                    ospread = span;
                    voxel_quality(2, ovc, orm, onb, span, p.max_gap, (p.sexp[lo] > 0.5) || (p.sexp[up] > 0.5), oscore, oflag);
                }
            }
        }
        p.u[o] = ou; p.v[o] = ov; p.rmse[o] = orm; p.spread[o] = ospread; p.score[o] = oscore;
        p.valid_count[o] = ovc; p.source_count[o] = osrc; p.neighbor_count[o] = onb; p.support[o] = osup;
        p.flag[o] = oflag;
    }
}

}  // namespace pcg
