// pcg_network.cuh — K_net: the multi-radar merge of run_radar_network_3d as ONE pass per field.
//
// Replaces, per output voxel and fused over all radars (pycwr/interp/RadarInterp.py):
//   grid_single_radar_to_latlon_3d  :963-1023   get_CAPPI_3d on the radar-local grid + range-sanity filter
//   _compute_quality_weight_volume  :603-696    observable gating x range / beam / vertical terms
//   _compute_observability_mask     :540-600    the same gating -> coverage_count, blind_mask (:1938-1945)
//   _build_composite_weight_cache   :1379-1396  exp(-4 d^2 / R^2) (or d^2 for "nearest")
//   compose_network_volume          :1399-1487  max / nearest / exp_weighted / quality_weighted, valid_count
//   grid_single_radar_cr_to_latlon + CR max-merge :1026-1068, 2047-2063
// The reference materialises an (nradar, nz, nlat, nlon) float64 stack per quantity; here the
// per-radar values live in registers and only the merged outputs touch HBM.
//
// Decomposition: radar-major passes over per-radar tile lists (see below): the per-radar values live
// only in registers, the merged accumulators (10 bytes per voxel) in HBM.
//
// Exactness: values, valid_count come from the same voxel body as get_CAPPI_3d (indices bit-exact).
// The observability gates compare the reference's NumPy-twin geometry (np.hypot / ranges**2,
// pycwr/core/transforms.py:201-222) with table entries; the twin agrees with the gridding geometry
// to a few ulps, so a voxel is decided from the gridding's r / elevation bracket unless it lies
// within a guard distance of a gate, in which case the twin chain itself is evaluated.
#pragma once
#include "pcg_r2.cuh"

namespace pcg {

#ifndef NET_MINB
#define NET_MINB 8
#endif
constexpr int kNetThreads = 128;
constexpr int kNetMaxSweeps = 32;    // per radar, bounded by the shared-memory azimuth table

enum NetMethod { kNetMax = 0, kNetNearest = 1, kNetExp = 2, kNetQuality = 3 };
enum NetTerm { kTermRange = 1, kTermBeam = 2, kTermVertical = 4 };

struct NetPair {                     // sweep pair (k, k+1) of one radar
    double obs_first, obs_last;      // max(first gates), min(last gates) (RadarInterp.py:671-672)
    // the same two gates in r2 space with a relative guard of 1e-9 (the NumPy twin's r agrees with the
    // gridding's to a few ulps): r2 inside (of2_hi, ol2_lo) is certainly observable, outside
    // [of2_lo, ol2_hi] certainly not, anything else is decided by the twin chain itself
    double of2_lo, of2_hi, ol2_lo, ol2_hi;
    float beam_k;                    // tan(deg2rad(0.5*(bw_k + bw_k+1)) * 0.5) / beam_radius_ref (0: term off)
    float vert_w;                    // 1 / (1 + (gap_half / vertical_gap_ref)^2)             (1: term off)
};

struct NetPairQ {                    // per (radar, sweep k): what a sampled voxel of pair (k, k+1) reads besides its
    float q1, q2, q3, q4, q5;        // level row — the elevation-weight polynomial (FastSweep) and the quality terms
    float beam_k, vert_w;            // of NetPair, side by side: two 128-bit loads (tile-persistent kernel)
    float pad;
};

struct NetLevel {                    // per (radar, level) constants, host-evaluated
    LevelConst cy;                   // the gridding chain (RadarGridC.pyx:145-146)
    double a2g2_np, twoag_np, g2_np; // the NumPy twin: a**2 + g**2, 2*a*g, g**2 (transforms.py:213-218)
};

struct NetRadar {
    VolumeView vol;
    const PairDesc* pairs;           // thresholds pre-scaled by 2a (this radar's height)
    const NetPair* npairs;
    const NetLevel* levels;          // [nz]
    R2Ctx r2;                        // r2-space tables of this radar (lp: [nz] rows of 2 ne + 3 records)
    double x, y;                     // site on the common grid (m)
    double Re, a, a2, twoa, guard, inv_twoa;
    double a2_np;                    // a**2 as Python evaluates it
    double cull2;                    // columns with d^2 above this are out of reach of every gate
    double src_lo, src_hi;           // range-sanity window (src_lo > src_hi: no valid source value -> all fill)
    float inv_decay;                 // 1 / range_decay_m   (0: term off)
    float inv_r2inf4;                // 4 / influence_radius^2
    int ne;
    // the range-sanity window for FLOAT values: for a float v, (double)v < src_lo <=> v < src_lo_f (the smallest
    // float >= src_lo) and (double)v > src_hi <=> v > src_hi_f (the largest float <= src_hi); no valid source
    // value: +inf / -inf, which no finite v passes
    float src_lo_f, src_hi_f;
    int pad;
    const NetPairQ* pq;              // [ne]
};

constexpr int kNetMaxPeers = 16;

struct NetArgs {
    const NetRadar* radars;
    int nradar;
    int nz;
    const double* gx;
    const double* gy;
    long long ncol;
    int nx, ny;                      // grid shape (ny fastest)
    int py_log2, tiles_y;            // warp footprint: 2^py_log2 columns x (32 >> py_log2) rows; blocks per grid row band
    double fill;
    float fill32;
    int quality;                     // multiply the 2-D weight by the geometric quality weight
    int sanity;                      // apply the range-sanity filter
    int nearest_gate, lowest_sweep, highest_sweep;
    double* composite;               // [nz][ncol]
    int16_t* valid_count;            // [nz][ncol] or NULL
    int16_t* coverage;               // [nz][ncol] or NULL
    int8_t* blind;                   // [nz][ncol] or NULL
    double* cr;                      // [ncol] or NULL (NaN where no radar has an echo)
    double* quality_out;             // [nz][ncol] or NULL: quality weight of the LAST radar (per-radar diagnostics)
    unsigned long long* tie_count;   // voxels whose observability was decided by the twin chain
    // tile-persistent kernel (pcg_net2.cuh)
    double out_missing;              // value stored in empty composite cells (fill, or NaN for the dataset layout)
    int ne_max;                      // most sweeps of any radar: rows of the shared-memory azimuth table
    int z0;                          // first level of this launch (level chunks when the accumulators exceed shared memory)
    // multi-GPU: this rank's slab of the composite is also stored into the full grid of every peer (NVLink P2P
    // stores from the finalize kernel: the all-gather of the finished slabs, fused with producing them)
    int npeer;
    long long peer_plane;            // columns of one level of the FULL grid
    long long peer_off;              // first column of this slab inside a level of the full grid
    double* peer[kNetMaxPeers];      // composite [nz][peer_plane] on every GPU of the job (own buffer included)
};

// hypot with a single rounding (what glibc >= 2.35 returns): sqrt of the double-double x^2 + y^2
// followed by one Newton correction on the residual.
__device__ __forceinline__ double hypot_cr(double x, double y) {
    x = fabs(x); y = fabs(y);
    if (y > x) { const double t = x; x = y; y = t; }
    if (x == 0.0) return 0.0;
    const double xx = mul(x, x), xl = __fma_rn(x, x, -xx);
    const double yy = mul(y, y), yl = __fma_rn(y, y, -yy);
    const double sh = add(xx, yy);
    const double sl = add(add(sub(xx, sh), yy), add(xl, yl));   // x^2 + y^2 = sh + sl (xx >= yy)
    const double h = sqrt_rn(sh);
    // residual of h^2 against sh + sl
    const double hh = mul(h, h), hl = __fma_rn(h, h, -hh);
    const double res = add(sub(sub(sh, hh), hl), sl);
    return add(h, div(res, mul(2.0, h)));
}

// The NumPy-twin observability of one voxel (RadarInterp.py:566-599 with transforms.py:201-222).
__device__ __noinline__ bool twin_observable(double x, double y, const NetRadar& R, const NetLevel& L,
                                             const SweepDesc* s_sw, const NetPair* s_np, int ne) {
    const double s = hypot_cr(x, y);
    const double cphi = cos_small(div(s, R.Re));
    double r2 = sub(L.a2g2_np, mul(L.twoag_np, cphi));
    if (!(r2 > 0.0)) r2 = (r2 != r2) ? r2 : 0.0;
    const double r = sqrt_rn(r2);
    // cos_arg = (a**2 + ranges**2 - g**2) / (2*a*ranges); 0/0 -> NaN -> not finite -> unobservable
    const double c0 = div(sub(add(R.a2_np, mul(r, r)), L.g2_np), mul(R.twoa, r));
    if (c0 != c0) return false;
    const double c = c0 > 1.0 ? 1.0 : (c0 < -1.0 ? -1.0 : c0);
    const double el = div(mul(sub(acos(c), kHalfPi), 180.0), kPi);
    if (!(el >= s_sw[0].elev) || !(el <= s_sw[ne - 1].elev)) return false;
    int lo = 0, hi = ne;                              // np.searchsorted(side="right")
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (el < s_sw[mid].elev) hi = mid; else lo = mid + 1; }
    int up = lo < 1 ? 1 : (lo > ne - 1 ? ne - 1 : lo);
    const NetPair& np = s_np[up - 1];
    return (r >= np.obs_first) && (r <= np.obs_last);
}

// ---------------------------------------------------------------------------------------------
// Radar-major passes.  One launch per radar over the grid tiles its range disc reaches:
//   net_cull_kernel      per tile (one block = 128 columns = 4 warps x (8 x 4) footprint): bounding box of
//                        the tile's columns, tested against every radar's reach -> per-radar tile lists
//   net_radar_kernel     one radar, its tile list: the column's azimuth table, then every level through the
//                        r2 voxel path, read-modify-write of the per-voxel accumulators in HBM
//   net_finalize_kernel  accumulators -> composite / valid_count / coverage_count / blind_mask / CR
// Within a pass every block gathers from the SAME 97 MB volume, which stays L2-resident (126 MB), the
// descriptor tables of the radar sit in shared memory, the column prologue runs once per (column, radar)
// and warps never idle on radars that miss them.  Passes run in radar order on one stream, so the sums
// accumulate in the reference's order (np.sum over axis 0) without atomics.  Weights use the hardware
// exp2 / reciprocal (relative error < 1e-6 for these arguments, against a 1e-5 bar on the weights).
// ---------------------------------------------------------------------------------------------
struct NetCell {               // one voxel's accumulators: a single 128-bit load / store per (voxel, radar)
    float wv;                  // weighted: sum w*v;  max / nearest: the chosen value
    float w;                   // weighted: sum w;    max / nearest: 1 when a value was chosen
    unsigned cnt;              // valid_count
    unsigned cov;              // coverage_count
};

struct NetAcc {
    NetCell* cell;             // [nz][ncol]
    double* best_d2;           // nearest only
    float* cr;                 // [ncol], NaN = no echo yet
    int* tiles;                // [nradar][ntile] tile lists
    int* tile_count;           // [nradar]
    int ntile;
};

__device__ __forceinline__ void net_tile_column(const NetArgs& p, int tile, bool& active, long long& col) {
    const int pyl = p.py_log2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int ly = lane & ((1 << pyl) - 1), lx = lane >> pyl;
    const int by = tile % p.tiles_y, bx = tile / p.tiles_y;
    int iy = ((by * nwarp + warp) << pyl) + ly;
    int ix = bx * (32 >> pyl) + lx;
    active = (ix < p.nx) & (iy < p.ny);
    ix = min(ix, p.nx - 1); iy = min(iy, p.ny - 1);
    col = (long long)ix * p.ny + iy;
}

__global__ void __launch_bounds__(kNetThreads) net_cull_kernel(const __grid_constant__ NetArgs p, const NetAcc acc) {
    __shared__ double s_red[4 * (kNetThreads / 32)];
    bool active;
    long long col;
    net_tile_column(p, blockIdx.x, active, col);
    const double gx = p.gx[col], gy = p.gy[col];            // inactive lanes repeat a valid column
    double x0 = gx, x1 = gx, y0 = gy, y1 = gy;
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        s_red[4 * w] = x0; s_red[4 * w + 1] = x1; s_red[4 * w + 2] = y0; s_red[4 * w + 3] = y1;
    }
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        x0 = fmin(x0, s_red[4 * w]); x1 = fmax(x1, s_red[4 * w + 1]);
        y0 = fmin(y0, s_red[4 * w + 2]); y1 = fmax(y1, s_red[4 * w + 3]);
    }
    // NaN coordinates poison fmin / fmax silently: a tile holding one is handed to every radar
    const bool odd = __syncthreads_or(!(gx == gx) || !(gy == gy));
    for (int ir = threadIdx.x; ir < p.nradar; ir += blockDim.x) {
        const NetRadar& R = p.radars[ir];
        const double ddx = fmax(fmax(x0 - R.x, R.x - x1), 0.0);
        const double ddy = fmax(fmax(y0 - R.y, R.y - y1), 0.0);
        if (odd || !(ddx * ddx + ddy * ddy > R.cull2)) {
            const int slot = atomicAdd(acc.tile_count + ir, 1);
            acc.tiles[(size_t)ir * acc.ntile + slot] = blockIdx.x;
        }
    }
}

template <int METHOD, bool UNI>
__global__ void __launch_bounds__(kNetThreads, NET_MINB) net_radar_kernel(const __grid_constant__ NetArgs p,
                                                                           const NetAcc acc, int ir) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NetRadar& R = p.radars[ir];
    const int ne = R.ne;
    const unsigned T = blockDim.x;
    SweepDesc* s_sw = reinterpret_cast<SweepDesc*>(smem_raw);
    PairDesc* s_pr = reinterpret_cast<PairDesc*>(s_sw + ne);
    FastSweep* s_fs = reinterpret_cast<FastSweep*>(s_pr + (ne - 1));
    NetPair* s_np = reinterpret_cast<NetPair*>(s_fs + ne);
    AzEntry* my_az = reinterpret_cast<AzEntry*>(s_np + (ne - 1)) + (size_t)threadIdx.x * ne;    // [thread][sweep]
    stage_sweeps(s_sw, R.vol.sweeps, ne);
    {
        const int w1 = (ne - 1) * (int)(sizeof(PairDesc) / sizeof(int)), w2 = ne * (int)(sizeof(FastSweep) / sizeof(int)),
                  w3 = (ne - 1) * (int)(sizeof(NetPair) / sizeof(int));
        const int* a = reinterpret_cast<const int*>(R.pairs);       // already scaled by 2a on the host
        const int* b = reinterpret_cast<const int*>(R.r2.fs);
        const int* c = reinterpret_cast<const int*>(R.npairs);
        for (int i = threadIdx.x; i < w1; i += T) reinterpret_cast<int*>(s_pr)[i] = a[i];
        for (int i = threadIdx.x; i < w2; i += T) reinterpret_cast<int*>(s_fs)[i] = b[i];
        for (int i = threadIdx.x; i < w3; i += T) reinterpret_cast<int*>(s_np)[i] = c[i];
    }
    __syncthreads();

    bool active;
    long long col;
    net_tile_column(p, acc.tiles[(size_t)ir * acc.ntile + blockIdx.x], active, col);
    if (!active) return;
    // radar-local coordinates (RadarInterp.py:990-991) and the 2-D distance of the weight cache
    const double x = sub(p.gx[col], R.x), y = sub(p.gy[col], R.y);
    const double d2 = add(mul(x, x), mul(y, y));
    if (!(d2 <= R.cull2)) return;                     // beyond every gate: nothing to add (NaN: reference chain below)
    const float fill32 = p.fill32;
    // column invariants of this radar (RadarGridC.pyx:142-146,158)
    const double s = sqrt_rn(d2);
    const double phi = div(s, R.Re);
    const double cphi = cos_small(phi);
    const double az = azimuth_from_xy(x, y);
    column_azimuth_table_r2<false>(s_sw, ne, R.r2.az_t, az, my_az, nullptr, 1);
    ColumnCtx c;
    c.s_sw = s_sw; c.s_pr = s_pr;
    c.my_az = my_az; c.my_iaz = nullptr; c.az_stride = 1;
    c.ne = ne;
    c.rg_t = reinterpret_cast<const double2*>(R.vol.rng);
    c.a2 = R.a2; c.twoa = R.twoa; c.guard = R.guard; c.inv_twoa = R.inv_twoa;
    c.fill32 = fill32;
    c.nearest_gate = p.nearest_gate != 0; c.lowest_sweep = p.lowest_sweep != 0;
    c.highest_sweep = p.highest_sweep != 0;
    c.my_tag = nullptr; c.azt = nullptr; c.az = 0.0;
    R2Ctx q = R.r2;
    q.fs = s_fs;
    // exp(-4 d^2 / R_inf^2) (RadarInterp.py:1389)
    const float w2d = (METHOD == kNetExp || METHOD == kNetQuality) ? __expf(-(float)d2 * R.inv_r2inf4) : 1.f;
    unsigned long long ties = 0;

    int k = -1, k6 = 0;
    const LevelPair* lp = R.r2.lp;
    const NetLevel* lv = R.levels;
    size_t o = (size_t)col;
    const double src_lo = R.src_lo, src_hi = R.src_hi;
    const float inv_decay = R.inv_decay;
#pragma unroll 1
    for (int i = 0; i < p.nz; ++i, o += (size_t)p.ncol, lp += 2 * ne + 3) {
        // the voxel's accumulators are requested before the (long) voxel body so that their latency hides
        // behind it (22 ms -> 15 ms on C4)
        NetCell cell = acc.cell[o];
        bool dirty = false;
        float res[1];
        VoxelInfo vi;
        r2_voxel<1, false, UNI>(c, q, R.vol.val, lv[i].cy, lp, cphi, k, k6, res, vi);
        float v = res[0];
        // range-sanity filter (RadarInterp.py:1003-1014), compared in double like the reference
        if (p.sanity && v != fill32) {
            const double vd = (double)v;
            if (!(src_lo <= src_hi) || vd < src_lo || vd > src_hi) v = fill32;
        }
        const bool valid = (v != fill32) && (v == v) && (fabsf(v) != __int_as_float(0x7f800000));
        // observability (RadarInterp.py:566-599 / 636-684)
        const int kk = vi.lower < ne - 1 ? vi.lower : ne - 2;
        const bool is_pair = vi.bracket == kPair;
        const NetPair& np = s_np[is_pair ? kk : 0];
        const double r2 = vi.r2;
        bool obs = is_pair & (r2 > np.of2_hi) & (r2 < np.ol2_lo);
        const bool sure_out = (r2 < np.of2_lo) | (r2 > np.ol2_hi);
        if (vi.tie || (is_pair & !obs & !sure_out) || !(r2 > 0.0)) {
            obs = twin_observable(sub(p.gx[col], R.x), sub(p.gy[col], R.y), R, lv[i], s_sw, s_np, ne);
            ++ties;
        }
        float qw = 0.f;
        if (obs) {
            // RadarInterp.py:687-695
            const float rf = (float)vi.r;
            const float t = rf * inv_decay;
            const float b = rf * np.beam_k;
            qw = __expf(-t * t) * rcp_approx(fmaf(b, b, 1.f)) * np.vert_w;
            cell.cov += 1;
            dirty = true;
        }
        if (p.quality_out) p.quality_out[o] = (double)qw;
        if (valid) {
            cell.cnt += 1;
            dirty = true;
            if (METHOD == kNetMax) {
                if (cell.w == 0.f || v > cell.wv) cell.wv = v;
                cell.w = 1.f;
            } else if (METHOD == kNetNearest) {
                if (d2 < acc.best_d2[o]) { acc.best_d2[o] = d2; cell.wv = v; cell.w = 1.f; }
            } else {
                const float w = (METHOD == kNetQuality && p.quality) ? w2d * qw : w2d;
                cell.wv = fmaf(w, v, cell.wv);
                cell.w += w;
            }
        }
        if (dirty) acc.cell[o] = cell;
    }

    if (p.cr != nullptr) {
        // get_CR_xy on the local grid: every sweep at its fixed elevation, blind "mask"
        // (RadarGridC.pyx:596-624), then the NaN-aware max over radars (RadarInterp.py:2047-2063)
        const double sphi = sin_small(phi);
        const double2* rg_t = reinterpret_cast<const double2*>(R.vol.rng);
        const float4* val = static_cast<const float4*>(R.vol.val[0]);
        float best = acc.cr[col];
        for (int ie = 0; ie < ne; ++ie) {
            const SweepDesc& d = s_sw[ie];
            if (d.naz < 2 || d.ng < 2) continue;
            const double denom = sub(cphi, mul(sphi, d.tan_e));
            if (denom == 0.0) continue;
            const double g = div(R.a, denom);
            double r2 = sub(add(R.a2, mul(g, g)), mul(mul(R.twoa, g), cphi));
            if (r2 < 0.0) r2 = 0.0;
            const double r = sqrt_rn(r2);
            if (r != r) continue;
            const RangeBracket rb = range_bracket(d, rg_t, r, false);
            const float v = sample_pairs(val, my_az[ie], rb, fill32);
            if (v == fill32 || v != v) continue;
            if (!(best >= v)) best = v;               // NaN (nothing yet) or smaller
        }
        acc.cr[col] = best;
    }
    if (ties && p.tie_count) atomicAdd(p.tie_count, ties);
}

template <int METHOD>
__global__ void __launch_bounds__(256) net_finalize_kernel(const __grid_constant__ NetArgs p, const NetAcc acc) {
    const size_t nvox = (size_t)p.nz * (size_t)p.ncol;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (size_t)gridDim.x * blockDim.x) {
        const NetCell cell = acc.cell[i];
        const float awv = cell.wv, aw = cell.w;
        double out = p.fill;
        if (METHOD == kNetMax || METHOD == kNetNearest) {
            if (aw != 0.f) out = (double)awv;
        } else {
            if (aw > 0.f) out = (double)awv / (double)aw;            // has_data = total_weight > 0
        }
        const unsigned c8 = cell.cov;
        if (p.composite) p.composite[i] = out;
        if (p.npeer) {
            // every warp stores its 256 contiguous bytes to all peers at once: the ranks' traffic is spread over
            // all NVLink destinations all the time.  (Measured at N = 8, 1.9 GB of egress per GPU: 5.3 ms for the
            // whole step = NCCL's all_gather after the kernels; one block row per peer took 10.3 ms with every rank
            // storing to the same peer first, 7.2 ms with the peer order rotated by rank.)
            const size_t iz = i / (size_t)p.ncol;
            const size_t at = iz * (size_t)p.peer_plane + (size_t)p.peer_off + (i - iz * (size_t)p.ncol);
            for (int q = 0; q < p.npeer; ++q) p.peer[q][at] = out;
        }
        if (p.valid_count) p.valid_count[i] = (int16_t)cell.cnt;
        if (p.coverage) p.coverage[i] = (int16_t)c8;
        if (p.blind) p.blind[i] = (int8_t)(c8 == 0u);
        if (p.cr && i < (size_t)p.ncol) p.cr[i] = (double)acc.cr[i];   // NaN where no radar has an echo
    }
}

// ---------------------------------------------------------------------------------------------
// gates_kernel — _compute_observability_mask (RadarInterp.py:540-600) and
// _compute_quality_weight_volume (:603-696) as stand-alone functions: the NumPy-twin chain
// (transforms.py:201-222) restated in FP64 for every voxel, any sweep count (including the
// single-sweep beam-width gate).  Needs only the range-table ends and the fixed elevations.
// ---------------------------------------------------------------------------------------------
struct GateSweep { double elev, r_first, r_last; };

struct GatesArgs {
    const GateSweep* sweeps;      // [ne]
    const NetPair* pairs;         // [ne-1]: obs_first/last; beam_k, vert_w unused (double versions below)
    const double* beam_tan;       // [max(ne-1,1)]: tan(deg2rad(representative_beam) * 0.5)
    const double* gap_half;       // [max(ne-1,1)]
    const NetLevel* levels;       // [nz]
    int ne, nz;
    const double* gx;
    const double* gy;
    long long ncol;
    double radar_x, radar_y, Re, a2_np, twoa;
    double beam_half;             // single sweep: max(bw/2, 0.1)
    int terms;                    // kTerm* bits
    double range_decay, beam_ref, gap_ref;
    unsigned char* mask;          // [nz][ncol] or NULL
    double* quality;              // [nz][ncol] or NULL
};

__global__ void __launch_bounds__(256) gates_kernel(const __grid_constant__ GatesArgs p) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.ncol) return;
    const double x = sub(p.gx[col], p.radar_x), y = sub(p.gy[col], p.radar_y);
    const double cphi = cos_small(div(hypot_cr(x, y), p.Re));
    const int ne = p.ne;
    for (int iz = 0; iz < p.nz; ++iz) {
        const NetLevel& L = p.levels[iz];
        double r2 = sub(L.a2g2_np, mul(L.twoag_np, cphi));
        if (!(r2 > 0.0)) r2 = (r2 != r2) ? r2 : 0.0;
        const double r = sqrt_rn(r2);
        const double c0 = div(sub(add(p.a2_np, mul(r, r)), L.g2_np), mul(p.twoa, r));
        bool obs = (c0 == c0) && (r == r);
        int pair = 0;
        if (obs) {
            const double c = c0 > 1.0 ? 1.0 : (c0 < -1.0 ? -1.0 : c0);
            const double el = div(mul(sub(acos(c), kHalfPi), 180.0), kPi);
            if (ne == 1) {
                obs = (fabs(sub(el, p.sweeps[0].elev)) <= p.beam_half) && (r >= p.sweeps[0].r_first) &&
                      (r <= p.sweeps[0].r_last);
            } else {
                obs = (el >= p.sweeps[0].elev) && (el <= p.sweeps[ne - 1].elev);
                if (obs) {
                    int lo = 0, hi = ne;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (el < p.sweeps[mid].elev) hi = mid; else lo = mid + 1; }
                    pair = (lo < 1 ? 1 : (lo > ne - 1 ? ne - 1 : lo)) - 1;
                    obs = (r >= p.pairs[pair].obs_first) && (r <= p.pairs[pair].obs_last);
                }
            }
        }
        const size_t o = (size_t)iz * (size_t)p.ncol + (size_t)col;
        if (p.mask) p.mask[o] = obs ? 1 : 0;
        if (p.quality) {
            double w = 0.0;
            if (obs) {
                w = 1.0;
                if (p.terms & kTermRange) { const double t = div(r, p.range_decay); w = mul(w, exp(-mul(t, t))); }
                if (p.terms & kTermBeam) {
                    const double b = div(mul(r, p.beam_tan[pair]), p.beam_ref);
                    w = mul(w, div(1.0, add(1.0, mul(b, b))));
                }
                if ((p.terms & kTermVertical) && ne > 1) {
                    const double gq = div(p.gap_half[pair], p.gap_ref);
                    w = mul(w, div(1.0, add(1.0, mul(gq, gq))));
                }
            }
            p.quality[o] = w;
        }
    }
}

// min / max / count of the valid (finite, != fill) source values of one sweep tile: the window of
// the range-sanity filter (RadarInterp.py:1003-1007).  Doubles are ordered through their bit
// patterns (sign-flipped) so unsigned atomics implement min / max.
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void minmax_kernel(const double* __restrict__ src, size_t n, double fill,
                              unsigned long long* __restrict__ out /* [min_key, max_key, count] */) {
    unsigned long long lo = ~0ull, hi = 0ull, cnt = 0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = src[i];
        if (v == fill || v != v || fabs(v) == __longlong_as_double(0x7ff0000000000000LL)) continue;
        const unsigned long long kx = order_key(v);
        lo = kx < lo ? kx : lo;
        hi = kx > hi ? kx : hi;
        ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_down_sync(0xffffffffu, lo, o);
        const unsigned long long h2 = __shfl_down_sync(0xffffffffu, hi, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    // one set of atomics per BLOCK: thousands of warps hitting three addresses serialise (38 us -> 5 us)
    __shared__ unsigned long long s_lo[32], s_hi[32], s_cnt[32];
    const int warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[warp] = lo; s_hi[warp] = hi; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarp; ++w) {
            lo = s_lo[w] < lo ? s_lo[w] : lo;
            hi = s_hi[w] > hi ? s_hi[w] : hi;
            cnt += s_cnt[w];
        }
        if (cnt) {
            atomicMin(out, lo);
            atomicMax(out + 1, hi);
            atomicAdd(out + 2, cnt);
        }
    }
}

}  // namespace pcg
