// pcg_r2.cuh — the voxel path in r^2 space (kernel v7).
//
// The reference's per-voxel chain is  r2 = a^2 + g^2 - 2ag cos(phi)  ->  r = sqrt(r2)  ->
// el = acos((a^2 + r2 - g^2) / (2 a r)) - pi/2  ->  bisect(el), bisect(r)  (RadarGridC.pyx:146-157,
// 424-439, 275-295).  Everything after r2 is a MONOTONE function of r2 at a fixed level, so every
// comparison the reference makes can be made on r2 itself against host-precomputed thresholds:
//
//  * gate bracket:   r >= R[j]  <=>  r2 >= tg[j],  tg[j] = the smallest double whose IEEE square
//    root reaches R[j] (exact: sqrt is monotone and correctly rounded on both sides).
//  * sweep bracket:  el >= e_k  <=>  r <= rho_k(z) = -a sin e_k + sqrt(g^2 - a^2 cos^2 e_k), the slant
//    range at which beam k crosses the level.  The reference decides on its ROUNDED el, so the
//    threshold carries a guard band [t_lo, t_hi] = rho_k^2 -+ margin that bounds every rounding of the
//    reference chain (a^2, g^2 and the two additions of `num`: 2 ulp(max(a^2, g^2)), i.e. <= 0.06 in r2
//    for Re + h ~ 8.5e6; the division / acos / scaling: <= 3e-8 * rho); voxels inside a band (~1e-5 of a COLUMN per
//    threshold) take the reference's own chain.  Levels at or below the antenna with sweeps that are
//    not clearly above the horizon are flagged and take that chain entirely.
//
// What remains per voxel in FP64: r2 (one multiply, one add — bit-identical to the reference), four
// threshold compares, two gate compares and two subtractions whose results feed FP32 weights:
//   w_r  = (r2 - R[j]^2) / (r + R[j]) / (R[j+1] - R[j])
//   u    = c - C_k = (r2 - rho_k^2) / (r + rho_k) * (1/2a + D / (r rho_k)),   D = (g^2 - a^2) / 2a
//   w_el = u (q1 + q2 u + .. + q5 u^4)          (per sweep pair, host-fitted and validated to 1e-7)
// No FP64 square root, division or polynomial is left on the common path.
#pragma once
#include "pcg_fast.cuh"

namespace pcg {

struct LevelSweep {               // per (level, slot); slot k + 1 <-> sweep k, slots 0 and ne + 1 are sentinels
    double t_lo, t_hi;            // r2 < t_lo: certainly el >= e_k;   r2 > t_hi: certainly el < e_k
    double rho2;                  // rho_k^2
    float rho, b1;                // rho_k,  D / rho_k
};

struct LevelPair {                // per (level, k + 1), k = -1 .. ne-1: what confirms and samples pair k
    double t_lo, t_hi;            // t_lo[k], t_hi[k+1]: r2 strictly between them <=> certainly e_k <= el < e_{k+1}
    double rho2;                  // of sweep k
    float rho, b1;
};

enum FastSweepFlag { kFsPolyOk = 1, kFsShared = 2, kFsUsable = 4 };

struct FastSweep {                // per sweep k (and the pair (k, k+1)), constant per volume
    double tg_first, tg_gt_last;  // r >= R[0] <=> r2 >= tg_first;   r <= R[ng-1] <=> r2 < tg_gt_last
    unsigned rng_off;
    int ngm1;
    float c0, inv_step;           // gate guess: floor(r * inv_step + c0),  c0 = 1 - R[0] * inv_step
    float q1, q2, q3, q4;
    float q5;
    int flags;                    // kFsPolyOk (pair poly validated), kFsShared (k and k+1 share one range table),
    int pad0, pad1;               // kFsUsable (sweep non-degenerate, tables strictly increasing)
};

struct GateAux { double rsq; float r, inv_dr; };   // R[j]^2, R[j], 1 / (R[j+1] - R[j])

struct R2Ctx {
    const LevelPair* lp;          // [nz] rows of [LevelPair x (ne + 1)][LevelSweep x (ne + 2)] (32-byte records)
    const FastSweep* fs;          // [ne]
    const double2* g_t;           // {tg[j], tg[j+1]} per gate (tg_gt_last at the end)
    const GateAux* h_t;
    const double2* az_t;          // {A[i], 1 / (A[i+1] - A[i])} per ray, same offsets as the azimuth arena
    float inv2a;
    // UNI kernels: the one range table every sweep shares
    unsigned u_rng_off;
    int u_ngm1;
    float u_inv_step, u_c0, u_step, u_r_first;
};

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// the reference chain for one voxel (v6 body), out of line: rare on the r2 path.  Everything crosses
// the call in a private frame so that the caller's state never has its address taken.
struct SlowFrame {
    ColumnCtx c;
    LevelConst lc;
    VoxelInfo vi;
    int k6;
};

template <int NF, bool EMIT, bool LAZY>
__device__ __noinline__ void slow_voxel(SlowFrame* fr, const void* const* val, double cphi, float* res) {
    float tmp[NF];
    fast_voxel<NF, EMIT, LAZY>(fr->c, val, fr->lc, cphi, fr->k6, tmp, fr->vi);
#pragma unroll
    for (int f = 0; f < NF; ++f) res[f] = tmp[f];
}

// azimuth bracket of every sweep for one column, eagerly ([sweep] table per column)
template <bool EMIT>
__device__ __forceinline__ void column_azimuth_table_r2(const SweepDesc* s_sw, int ne, const double2* __restrict__ azt,
                                                        double az, AzEntry* my_az, int* my_iaz, unsigned stride) {
#pragma unroll 1
    for (int sw = 0; sw < ne; ++sw) {
        AzEntry e;
        int rec;
        az_bracket_one(s_sw[sw], azt, az, e, rec);
        my_az[(size_t)sw * stride] = e;
        if (EMIT) my_iaz[(size_t)sw * stride] = rec;
    }
}

// UNI: every sweep shares ONE range table with constant gate spacing and every sweep / pair is usable
// and poly-validated (the common case: PRD.get_vol_data with aligned ranges) — the range-table
// constants are then kernel uniforms, R[j] is an FMA and no per-sweep flag is consulted.  Otherwise
// they are read per sweep from `q.fs` and per gate from `q.h_t`.
template <int NF, bool EMIT, bool UNI, bool LAZY = false>
__device__ __forceinline__ void r2_voxel(const ColumnCtx& c, const R2Ctx& q, const void* const* val,
                                         const LevelConst& lc, const LevelPair* __restrict__ lp, double cphi,
                                         int& k, int& k6, float (&res)[NF], VoxelInfo& vi) {
    const int ne = c.ne;
    const float fill32 = c.fill32;
    // RadarGridC.pyx:146-149 — the reference's own operations: r2 is bit-identical
    const double r2 = sub(lc.a2g2, mul(lc.twoag, cphi));
    // One level row is [LevelPair x (ne + 1)][LevelSweep x (ne + 2)].  r2 <= 0 / NaN and levels flagged
    // for the reference chain need no test of their own: every threshold is positive (the top sentinel is
    // 0) and flagged rows hold -inf / +inf, so the confirmation fails and the walk ends in `slow`.
    const LevelSweep* ls = reinterpret_cast<const LevelSweep*>(lp + (ne + 1));
    bool slow = false;
    {
        // settle k in [-1, ne-1]: e_k <= el < e_{k+1} (k = -1 below the lowest sweep, ne-1 above the highest).
        // One 16-byte entry {t_lo[k], t_hi[k+1]} confirms the current k; the per-sweep bands are only
        // read when the column has to move.
        const double2 cf = __ldg(reinterpret_cast<const double2*>(lp + (k + 1)));
        if (!((r2 < cf.x) & (r2 > cf.y))) {
#pragma unroll 1
            for (;;) {
                const double2* row = reinterpret_cast<const double2*>(ls + (k + 1));
                const double2 cur = __ldg(row);
                const double2 nxt = __ldg(row + 2);
                if ((r2 < cur.x) & (r2 > nxt.y)) break;
                if (r2 > cur.y) { --k; continue; }
                if (r2 < nxt.x) { ++k; continue; }
                slow = true;                                  // inside a guard band
                break;
            }
        }
    }
    if (!slow) {
        const bool pair = (unsigned)k < (unsigned)(ne - 1);
        // blind-zone rules (RadarGridC.pyx:424-433): below the lowest / above the highest sweep
        const bool skip = ((k < 0) & !c.lowest_sweep) | ((k >= ne - 1) & !c.highest_sweep);
        const int lower = max(k, 0);
        vi.bracket = pair ? kPair : (k < 0 ? kSkipLow : kSkipHigh);
        vi.tie = false;
        vi.state = skip ? vi.bracket : (pair ? kPair : kSingle);
        vi.lower = lower; vi.upper = lower + (pair ? 1 : 0); vi.r2 = r2; vi.have_r = false; vi.r = 0.0;
        if (skip) {
#pragma unroll
            for (int f = 0; f < NF; ++f) res[f] = fill32;
            if (EMIT) { vi.iaz0 = vi.ir0 = vi.fl0 = vi.iaz1 = vi.ir1 = vi.fl1 = -1; }
            return;
        }
        // ---- sample sweep `lower` (and `lower + 1` for a pair) ----
        const FastSweep* f = q.fs + lower;
        unsigned rng_off = q.u_rng_off;
        int ngm1 = q.u_ngm1;
        float inv_step = q.u_inv_step, c0 = q.u_c0;
        bool usable = true;
        if (!UNI) {
            const int4 tab = *reinterpret_cast<const int4*>(&f->rng_off);      // rng_off, ngm1, c0, inv_step
            rng_off = (unsigned)tab.x; ngm1 = tab.y; c0 = __int_as_float(tab.z); inv_step = __int_as_float(tab.w);
            const int need = pair ? (kFsPolyOk | kFsShared | kFsUsable) : kFsUsable;
            usable = (f->flags & need) == need;
        }
        const float r2f = (float)r2;
        const float rinv = rsqrt_approx(r2f);
        const float rf = r2f * rinv;
        // (a degenerate sweep has no gate interval to index: nothing is loaded for it)
        const int gi = usable ? min(max(__float2int_rd(fmaf(rf, inv_step, c0)), 1), ngm1) : 1;
        const unsigned gidx = usable ? rng_off + (unsigned)(gi - 1) : q.u_rng_off;
        const double2 tg = __ldg(q.g_t + gidx);          // r in [R[gi-1], R[gi])  <=>  tg.x <= r2 < tg.y
        if (usable & (r2 >= tg.x) & (r2 < tg.y)) {
            float w_r;
            if (UNI) {
                // regular table: R[gi-1] = R[0] + (gi-1) * step, and tg.x is R[gi-1]^2 to within two ulps
                const float rj = fmaf((float)(gi - 1), q.u_step, q.u_r_first);
                w_r = (float)sub(r2, tg.x) * q.u_inv_step * rcp_approx(rf + rj);
            } else {
                const int4 hv = __ldg(reinterpret_cast<const int4*>(q.h_t + gidx));     // {R^2 (lo, hi), R, 1/dR}
                w_r = (float)sub(r2, __hiloint2double(hv.y, hv.x)) * __int_as_float(hv.w) *
                      rcp_approx(rf + __int_as_float(hv.z));
            }
            const AzEntry ea = az_entry<LAZY>(c, lower);
            vi.r = (double)rf;
            if (pair) {
                const int4 lk = __ldg(reinterpret_cast<const int4*>(lp + (k + 1)) + 1);  // {rho2 (lo, hi), rho, b1}
                const float u = (float)sub(r2, __hiloint2double(lk.y, lk.x)) * rcp_approx(rf + __int_as_float(lk.z)) *
                                fmaf(__int_as_float(lk.w), rinv, q.inv2a);
                const float4 qa = *reinterpret_cast<const float4*>(&f->q1);
                float w_el = fmaf(f->q5, u, qa.w);
                w_el = fmaf(w_el, u, qa.z);
                w_el = fmaf(w_el, u, qa.y);
                w_el = fmaf(w_el, u, qa.x);
                w_el *= u;
                const AzEntry eb = az_entry<LAZY>(c, lower + 1);
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const float4* v = static_cast<const float4*>(val[fi]);
                    const float4 qa4 = __ldg(v + (ea.row + (unsigned)gi)), qb4 = __ldg(v + (eb.row + (unsigned)gi));
                    const float ea0 = fmaf(ea.w, qa4.y - qa4.x, qa4.x), ea1 = fmaf(ea.w, qa4.w - qa4.z, qa4.z);
                    const float eb0 = fmaf(eb.w, qb4.y - qb4.x, qb4.x), eb1 = fmaf(eb.w, qb4.w - qb4.z, qb4.z);
                    const float v0 = lerp_fill(w_r, ea0, ea1, fill32);
                    const float v1 = lerp_fill(w_r, eb0, eb1, fill32);
                    res[fi] = lerp_fill(w_el, v0, v1, fill32);
                }
            } else {
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const float4* v = static_cast<const float4*>(val[fi]);
                    const float4 qa4 = __ldg(v + (ea.row + (unsigned)gi));
                    const float ea0 = fmaf(ea.w, qa4.y - qa4.x, qa4.x), ea1 = fmaf(ea.w, qa4.w - qa4.z, qa4.z);
                    res[fi] = lerp_fill(w_r, ea0, ea1, fill32);
                }
            }
            if (EMIT) {
                const int ra = az_record<LAZY>(c, lower);
                vi.ir0 = gi; vi.iaz0 = ra & 0x3fffffff; vi.fl0 = (ra >> 30) & 1;
                vi.iaz1 = vi.ir1 = vi.fl1 = -1;
                if (pair) {
                    const int rb2 = az_record<LAZY>(c, lower + 1);
                    vi.ir1 = gi; vi.iaz1 = rb2 & 0x3fffffff; vi.fl1 = (rb2 >> 30) & 1;
                }
            }
            return;
        }
    }
    // ---- anything unusual: the reference's own chain (guard band, flagged level, r2 <= 0, ragged or
    //      degenerate tables, range gating / clamping, gate guess off) ----
    SlowFrame fr;
    fr.c = c; fr.lc = lc; fr.k6 = k6;
    float tmp[NF];
    slow_voxel<NF, EMIT, LAZY>(&fr, val, cphi, tmp);
#pragma unroll
    for (int f = 0; f < NF; ++f) res[f] = tmp[f];
    k6 = fr.k6;
    vi = fr.vi;
    vi.r2 = r2; vi.have_r = true;
}

// ---------------------------------------------------------------------------------------------
// K_cappi3d (v7): get_CAPPI_3d / get_CAPPI_xy (ne >= 2) / one radar of the mosaic
// ---------------------------------------------------------------------------------------------
struct R2Args {
    FastArgs base;
    R2Ctx r2;
    int lazy_az;                        // many-sweep volume: azimuth brackets on demand (2-entry cache per column)
    int* nonfinite;                     // optional: set to 1 when a target coordinate is NaN / inf (host entry points)
};

#ifndef R2_MINB
#define R2_MINB 5
#endif
template <int NF, bool EMIT, bool UNI, bool LAZY>
__global__ void __launch_bounds__(256, R2_MINB) cappi3d_r2_kernel(const __grid_constant__ R2Args P) {
    const FastArgs& p = P.base;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ne = p.vol.ne;
    SweepDesc* s_sw = reinterpret_cast<SweepDesc*>(smem_raw);
    PairDesc* s_pr = reinterpret_cast<PairDesc*>(s_sw + ne);
    FastSweep* s_fs = reinterpret_cast<FastSweep*>(s_pr + (ne - 1));
    // azimuth entries per thread: [thread][sweep] (eager), or a tagged 2-entry cache [thread][2] (lazy)
    constexpr bool lazy = LAZY;
    const int nslot = lazy ? 2 : ne;
    AzEntry* s_az = reinterpret_cast<AzEntry*>(s_fs + ne);
    int* s_tag = reinterpret_cast<int*>(s_az + (size_t)nslot * blockDim.x);          // lazy only
    int* s_iaz = s_tag + (lazy ? (size_t)nslot * blockDim.x : 0);                   // EMIT only
    stage_sweeps(s_sw, p.vol.sweeps, ne);
    stage_pairs(s_pr, p.pairs, ne - 1, p.twoa);
    {
        const int words = ne * (int)(sizeof(FastSweep) / sizeof(int));
        const int* src = reinterpret_cast<const int*>(P.r2.fs);
        int* dst = reinterpret_cast<int*>(s_fs);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const float fill32 = p.fill32;
    const long long ncol = p.grid.ncol;
    // 2-D warp footprint: lanes cover PY consecutive columns (the fastest output axis) of 32 / PY rows, so a
    // warp's gathers stay inside a compact patch of the polar volume (fewer distinct lines per request)
    // while every row segment still stores PY * 8 contiguous bytes; 4 x (warps / 4) warps tile the block.
    long long col;
    {
        const int pyl = p.grid.py_log2;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        const int wy = nwarp >= 4 ? 4 : nwarp;
        const int ly = lane & ((1 << pyl) - 1), lx = lane >> pyl;
        const int by = blockIdx.x % p.grid.tiles_y, bx = blockIdx.x / p.grid.tiles_y;
        const int iy = ((by * wy + (warp % wy)) << pyl) + ly;
        const int ix = (bx * (nwarp / wy) + (warp / wy)) * (32 >> pyl) + lx;
        if (ix >= p.grid.nx || iy >= p.grid.ny) return;
        col = (long long)ix * p.grid.ny + iy;
    }

    double x = p.grid.gx[col], y = p.grid.gy[col];
    if (p.grid.shift) { x = sub(x, p.grid.radar_x); y = sub(y, p.grid.radar_y); }
    // column invariants (RadarGridC.pyx:142-146,158)
    const double s = sqrt_rn(add(mul(x, x), mul(y, y)));
    // NaN / inf targets: the reference's result hinges on NaN semantics operation by operation, which only the
    // float64 layout restates; tell the host so that it can redo the call there
    if (P.nonfinite && !(s < __longlong_as_double(0x7ff0000000000000LL))) *P.nonfinite = 1;
    const double cphi = cos_small(div(s, p.Re));
    const double az = azimuth_from_xy(x, y);

    ColumnCtx c;
    c.s_sw = s_sw; c.s_pr = s_pr;
    c.my_az = s_az + (size_t)threadIdx.x * nslot; c.my_iaz = EMIT ? s_iaz + (size_t)threadIdx.x * nslot : nullptr;
    c.az_stride = 1;
    // (only the lazy variant reads these: constants otherwise, so they cost no registers)
    c.my_tag = LAZY ? s_tag + (size_t)threadIdx.x * nslot : nullptr;
    c.azt = LAZY ? P.r2.az_t : nullptr;
    c.az = LAZY ? az : 0.0;
    c.ne = ne;
    c.rg_t = reinterpret_cast<const double2*>(p.vol.rng);
    c.a2 = p.a2; c.twoa = p.twoa; c.guard = p.guard; c.inv_twoa = p.inv_twoa;
    c.fill32 = fill32;
    c.nearest_gate = p.nearest_gate != 0; c.lowest_sweep = p.lowest_sweep != 0; c.highest_sweep = p.highest_sweep != 0;
    if (lazy) { c.my_tag[0] = -1; c.my_tag[1] = -1; }
    else column_azimuth_table_r2<EMIT>(s_sw, ne, P.r2.az_t, az, s_az + (size_t)threadIdx.x * ne,
                                       s_iaz + (size_t)threadIdx.x * ne, 1);
    R2Ctx q = P.r2;
    q.fs = s_fs;

    int k = -1, k6 = 0;
    double* out_col = p.out + col;
    const LevelPair* lp = P.r2.lp;
    const bool merge_max = p.merge_max != 0;
#pragma unroll 1
    for (int iz = 0; iz < p.nz; ++iz) {
        float res[NF];
        VoxelInfo vi;
        r2_voxel<NF, EMIT, UNI, LAZY>(c, q, p.vol.val, p.levels[iz], lp, cphi, k, k6, res, vi);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double* dst = out_col + (size_t)f * (size_t)p.field_stride;
            // the fill value is float-representable (checked on the host), so widening res restores it
            const double v = (double)res[f];
            if (merge_max) {
                if (res[f] != fill32) {
                    const double cur = *dst;
                    if (cur == p.fill || v > cur) *dst = v;
                }
            } else {
                *dst = v;
            }
        }
        if (EMIT) {
            int32_t* rec = p.idx + ((size_t)iz * (size_t)ncol + (size_t)col) * kIdxStride;
            const bool have0 = (vi.state == kSingle || vi.state == kPair), have1 = (vi.state == kPair);
            rec[0] = vi.state;
            rec[1] = have0 ? vi.lower : -1; rec[2] = have0 ? vi.upper : -1;
            rec[3] = have0 ? vi.iaz0 : -1; rec[4] = have0 ? vi.ir0 : -1; rec[5] = have0 ? vi.fl0 : -1;
            rec[6] = have1 ? vi.iaz1 : -1; rec[7] = have1 ? vi.ir1 : -1; rec[8] = have1 ? vi.fl1 : -1;
            rec[9] = -1;
        }
        out_col += ncol;
        lp += 2 * ne + 3;
    }
}

}  // namespace pcg
