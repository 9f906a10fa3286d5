// pcg_r3.cuh — kernel v8 of get_CAPPI_3d / get_CAPPI_xy / one radar of the mosaic (RadarGridC.pyx:356-593).
//
// Same decisions as v7 (pcg_r2.cuh: every comparison of the reference made on r^2 against host-built thresholds, so
// the selected sweeps / rays / gates are the reference's by construction), rebuilt around what the ncu profiles of v7
// showed (profiles/r01_final_cappi3d_r2.md, profiles/r02_*):
//   * v7 was bound by the L1 data pipe (75-85 % busy), not by instruction issue: a warp-wide 128-bit load writes
//     512 bytes back to the register file whether its 32 addresses differ or not, and v7 issued five per voxel.
//     Two of them (the sweep-pair confirmation entry and the pair record) are the same for all lanes that share a
//     sweep pair: here the threshold rows of a launch's levels and the pair polynomials ride in the KERNEL
//     PARAMETERS (constant bank, up to 32 764 bytes) and are read through the constant cache (CT kernels).
//   * the remaining loads of a level (gate band, two azimuth entries, two quads per moment) are requested TOGETHER,
//     before anything is compared, for the sweep pair the column is in (uniform-table volumes: every speculative
//     address is valid) — one memory round trip per level instead of three dependent ones; a column that leaves
//     its pair walks to the new one and asks again.
//   * the reference chain for the rare voxels that need it (guard bands, range gating ...) runs AFTER the level
//     loop on a bit mask of marked levels: the hot loop contains no call and keeps its state in registers;
//     loop-invariant parameters are pinned in registers (ptxas otherwise re-reads them every level).
//   * azimuth table [sweep][thread] (conflict-free), interior azimuth brackets verified in line with one repair
//     step, quads resolved against missing data in RANGE as well at pack time (pack_quads_kernel): a quad is either
//     all-fill or fill-free, so only the elevation lerp tests for fill.
//   * per-column prologue (22 % of the samples of the first v8 captures): the grid coordinates are requested before
//     the descriptors are staged, sqrt / cos / atan2 run ahead of the block barrier, and the azimuth brackets are
//     resolved three sweeps at a time (the guesses of a group and the table entries they point at are requested
//     together: three L2 round trips per column instead of nine).
//   * volumes with more than 16 sweeps (LAZY): the azimuth brackets of the column's CURRENT sweep pair live in
//     registers and are computed when the column moves; confirm / walk form one block with a warp barrier behind
//     it; voxels beyond the last gate are filled in line.  Config C5: 1.08 ms (pcg_r2.cuh) -> 0.60 ms.
// Measured and rejected on B200: carrying the gate band and the azimuth-lerped samples from level
// to level (63 % of the voxels keep both) — fewer L1 wavefronts, but nearly every warp has a lane that left its
// band, so the warp walks both paths and ends up slower; prefetching the next level's quads (LAZY, C5: 0.747 ms
// against 0.715 ms); five or nine azimuth brackets per group instead of three (C3: 0.3295 / 0.3646 ms against 0.3277).
#pragma once
#include "pcg_r2.cuh"

namespace pcg {

struct AzDesc {                   // what the azimuth prologue needs of a sweep: 32 bytes, two 128-bit loads
    double az_first, az_inv_step;
    unsigned az_off;
    int naz;
    unsigned val_off;
    int ng;
};

// general azimuth bracket (ends of the table, wrap, repairs, irregular tables): out of line
__device__ __noinline__ void az_bracket_general(const SweepDesc* d, const double2* azt, double az, AzEntry* e, int* rec) {
    AzEntry ee;
    int rr;
    az_bracket_one(*d, azt, az, ee, rr);
    *e = ee;
    *rec = rr;
}

// the same, returned in registers {row, w, rec} (nothing of the caller has its address taken)
__device__ __noinline__ int4 az_bracket_general_v(const SweepDesc* d, const double2* azt, double az) {
    AzEntry ee;
    int rr;
    az_bracket_one(*d, azt, az, ee, rr);
    return make_int4((int)ee.row, __float_as_int(ee.w), rr, 0);
}

// the reference chain for one voxel, everything by value (rare: guard bands, flagged levels, ragged or degenerate
// tables, range gating / clamping, wrong gate guesses)
struct SlowArgs {
    const SweepDesc* s_sw;
    const PairDesc* s_pr;
    const AzEntry* my_az;
    const int* my_iaz;
    const double2* rg_t;
    const void* const* val;
    unsigned az_stride;
    int ne;
    double a2, twoa, guard, inv_twoa;
    float fill32;
    int blind;                    // bit 0 nearest_gate, 1 lowest_sweep, 2 highest_sweep
    // LAZY kernels: my_az / my_iaz are the tagged 2-entry caches of pcg_fast.cuh (az_entry<true>)
    int* my_tag;
    const double2* azt;
    double az;
};

template <int NF, bool EMIT, bool LAZY = false>
__device__ __noinline__ void r3_slow_voxel(const SlowArgs* a, const LevelConst* lc, double cphi, int* k6, float* res,
                                           VoxelInfo* vi) {
    ColumnCtx c;
    c.s_sw = a->s_sw; c.s_pr = a->s_pr; c.my_az = a->my_az; c.my_iaz = a->my_iaz; c.az_stride = a->az_stride;
    c.ne = a->ne; c.rg_t = a->rg_t; c.a2 = a->a2; c.twoa = a->twoa; c.guard = a->guard; c.inv_twoa = a->inv_twoa;
    c.fill32 = a->fill32;
    c.nearest_gate = (a->blind & 1) != 0; c.lowest_sweep = (a->blind & 2) != 0; c.highest_sweep = (a->blind & 4) != 0;
    c.my_tag = LAZY ? a->my_tag : nullptr; c.azt = LAZY ? a->azt : nullptr; c.az = LAZY ? a->az : 0.0;
    float tmp[NF];
    int k = *k6;
    VoxelInfo v;
    fast_voxel<NF, EMIT, LAZY>(c, a->val, *lc, cphi, k, tmp, v);
    *k6 = k;
    *vi = v;
#pragma unroll
    for (int f = 0; f < NF; ++f) res[f] = tmp[f];
}

// The r2-space threshold rows of a launch's levels and the pair polynomials travel in the KERNEL PARAMETERS
// (constant bank, up to 32 764 bytes since CUDA 12.1) when they fit: a warp mostly shares its sweep pair, so the
// 32-byte confirmation / pair record of a voxel is one constant-cache broadcast instead of two 128-bit loads per
// LANE through the L1 data pipe — which, at 512 bytes of register write-back per warp-wide 128-bit load, was the
// limiter of v7 (l1tex data-pipe 75-85 % busy, measured; profiles/r02_*).  Rows that do not fit (many sweeps) stay
// in global memory (CT = false).
constexpr int kR3TabBytes = 27 * 1024;
constexpr int kR3PolySweeps = 16;
struct R3Args {
    R2Args a;
    float poly[kR3PolySweeps][8];                   // q1 .. q5 of pair (k, k+1), padded to 32 bytes
    double tab[kR3TabBytes / 8];                    // rows of this launch's levels (CT kernels), as 8-byte words
};
static_assert(sizeof(R3Args) <= 32764, "kernel parameters are limited to 32764 bytes");

#ifndef R3_MINB
#define R3_MINB 4
#endif
// Loop-invariant kernel parameters the level loop reads: held in registers (an opaque asm keeps the compiler from
// re-materialising them from the constant bank on every level, one issue slot each).  -DR3_NOPIN: leave it to ptxas.
#ifndef R3_NOPIN
// (adding a run-time zero ptxas cannot prove to be zero: an empty asm, a mov or a self-shuffle are all seen through)
__device__ __forceinline__ float r3_pin(float v, int z) { return __int_as_float(__float_as_int(v) + z); }
__device__ __forceinline__ int r3_pin(int v, int z) { return v + z; }
__device__ __forceinline__ unsigned r3_pin(unsigned v, int z) { return v + (unsigned)z; }
template <typename T>
__device__ __forceinline__ const T* r3_pin(const T* v, int z) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(v) + z);
}
#define R3_PIN(x) x = r3_pin(x, pin_zero)
#else
#define R3_PIN(x)
#endif

// Gathers from the packed volume can ask L2 to fetch the whole 128-byte line (PCG_LTC = 128) or two lines (256) on a
// miss instead of the one 32-byte sector.  Measured on B200 and left OFF: no change for the network kernel
// (11.73 / 11.73 / 11.70 ms for 0 / 128 / 256), 128 slower on config C3 (0.366 vs 0.343 ms), 256 the same as 0.
#ifndef PCG_LTC
#define PCG_LTC 0
#endif
__device__ __forceinline__ float4 ldg_vol(const float4* p) {
#if PCG_LTC == 128
    float4 v;
    asm volatile("ld.global.nc.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#elif PCG_LTC == 256
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ double2 ldg_vol(const double2* p) {
#if PCG_LTC == 128
    double2 v;
    asm volatile("ld.global.nc.L2::128B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#elif PCG_LTC == 256
    double2 v;
    asm volatile("ld.global.nc.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// shared-memory reads of the level loop through 32-bit shared addresses held in registers (the compiler would
// otherwise rebuild the carve-up offsets from `ne` at every use)
__device__ __forceinline__ AzEntry lds_az(unsigned addr) {
    AzEntry e;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.row), "=f"(e.w) : "r"(addr) : "memory");
    return e;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// One column's azimuth bracket of one sweep (RadarGridC.pyx:282-289): interior guesses verified in line against the
// two entries they bracket (entries g-1 and g must both exist: 1 <= g <= naz-1; jittered tables put the arithmetic
// guess off by one for a few lanes of almost every warp: one repair step in line), anything else through the
// general routine.
__device__ __forceinline__ void r3_az_bracket_loaded(const AzDesc& d, int g, double2 lo, double hi, const SweepDesc* sw_desc,
                                                     const double2* __restrict__ azt, double az, AzEntry& e, int& rec) {
    // (g, lo = T[g-1], hi = T[g].x: requested by the caller when 1 <= g <= naz-1)
    bool done = false;
    if (g >= 1 && g < d.naz) {
        const double2* Tt = azt + d.az_off + (unsigned)(g - 1);
        if (az < lo.x) {
            if (g >= 2) { --g; hi = lo.x; lo = __ldg(Tt - 1); }
        } else if (!(az < hi)) {
            if (g + 1 < d.naz) { ++g; lo = __ldg(Tt + 1); hi = __ldg(&Tt[2].x); }
        }
        if (!(az < lo.x) && (az < hi)) {
            e.row = d.val_off + (unsigned)(g - 1) * (unsigned)d.ng;
            e.w = (float)sub(az, lo.x) * (float)lo.y;
            rec = g;
            done = true;
        }
    }
    if (!done) {
        const int4 r = az_bracket_general_v(sw_desc, azt, az);
        e.row = (unsigned)r.x; e.w = __int_as_float(r.y); rec = r.z;
    }
}

__device__ __forceinline__ void r3_az_bracket(const AzDesc& d, const SweepDesc* sw_desc, const double2* __restrict__ azt,
                                              double az, AzEntry& e, int& rec) {
    const int g = __double2int_rd(mul(sub(az, d.az_first), d.az_inv_step)) + 1;
    double2 lo = make_double2(0.0, 0.0);
    double hi = 0.0;
    if (g >= 1 && g < d.naz) {
        const double2* Tt = azt + d.az_off + (unsigned)(g - 1);
        lo = __ldg(Tt);
        hi = __ldg(&Tt[1].x);
    }
    r3_az_bracket_loaded(d, g, lo, hi, sw_desc, azt, az, e, rec);
}

// LAZY (volumes with more than 16 sweeps, e.g. the 64-sweep phased-array config C5): a [sweep] azimuth table per
// column would take the block's shared memory (64 x 256 x 8 bytes) and most of it would never be read — a column
// crosses a fraction of the sweeps, climbing one pair at a time.  The brackets of the CURRENT pair live in registers
// instead: when the column moves up one pair the upper entry becomes the lower one and one new bracket is computed;
// the tagged 2-entry shared-memory cache of pcg_fast.cuh only serves the deferred reference chain.
#ifndef R3_LAZY_MINB
#define R3_LAZY_MINB 5
#endif
template <int NF, bool EMIT, bool UNI, bool MERGE, bool CT, bool LAZY = false>
__global__ void __launch_bounds__(256, LAZY ? R3_LAZY_MINB : R3_MINB) cappi3d_r3_kernel(const __grid_constant__ R3Args Q) {
    const R2Args& P = Q.a;
    const FastArgs& p = P.base;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ne = p.vol.ne;
    const unsigned T = blockDim.x;
    SweepDesc* s_sw = reinterpret_cast<SweepDesc*>(smem_raw);
    PairDesc* s_pr = reinterpret_cast<PairDesc*>(s_sw + ne);
    FastSweep* s_fs = reinterpret_cast<FastSweep*>(s_pr + (ne - 1));
    AzDesc* s_ad = reinterpret_cast<AzDesc*>(s_fs + ne);
    const int nslot = LAZY ? 2 : ne;
    AzEntry* s_az = reinterpret_cast<AzEntry*>(s_ad + ne);                 // [sweep][thread] (LAZY: [slot][thread])
    int* s_tag = reinterpret_cast<int*>(s_az + (size_t)nslot * T);         // LAZY only, [slot][thread]
    int* s_iaz = s_tag + (LAZY ? (size_t)nslot * T : 0);                   // EMIT only, [sweep or slot][thread]
    // the column and its coordinates first: the two loads are in flight while the descriptors are staged
    const long long ncol = p.grid.ncol;
    long long col;
    bool in_grid;
    {
        const int pyl = p.grid.py_log2;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = T >> 5;
        const int wy = nwarp >= 4 ? 4 : nwarp;
        const int ly = lane & ((1 << pyl) - 1), lx = lane >> pyl;
        const int by = blockIdx.x % p.grid.tiles_y, bx = blockIdx.x / p.grid.tiles_y;
        const int iy = ((by * wy + (warp % wy)) << pyl) + ly;
        const int ix = (bx * (nwarp / wy) + (warp / wy)) * (32 >> pyl) + lx;
        in_grid = ix < p.grid.nx && iy < p.grid.ny;
        col = in_grid ? (long long)ix * p.grid.ny + iy : 0;
    }
    double x = p.grid.gx[col], y = p.grid.gy[col];

    stage_sweeps(s_sw, p.vol.sweeps, ne);
    stage_pairs(s_pr, p.pairs, ne - 1, p.twoa);
    {
        const int words = ne * (int)(sizeof(FastSweep) / sizeof(int));
        const int* src = reinterpret_cast<const int*>(P.r2.fs);
        int* dst = reinterpret_cast<int*>(s_fs);
        for (int i = threadIdx.x; i < words; i += T) dst[i] = src[i];
        for (int i = threadIdx.x; i < ne; i += T) {
            const SweepDesc d = p.vol.sweeps[i];
            AzDesc a;
            a.az_first = d.az_first; a.az_inv_step = d.az_inv_step; a.az_off = d.az_off;
            a.naz = (d.naz >= 2 && d.ng >= 2) ? d.naz : 0;                 // 0: degenerate, general path
            a.val_off = (unsigned)d.val_off; a.ng = d.ng;
            s_ad[i] = a;
        }
    }
    // column invariants (RadarGridC.pyx:142-146,158): nothing here reads the staged tables, so it runs ahead of
    // the barrier
    if (p.grid.shift) { x = sub(x, p.grid.radar_x); y = sub(y, p.grid.radar_y); }
    const double s = sqrt_rn(add(mul(x, x), mul(y, y)));
    const double cphi = cos_small(div(s, p.Re));
    const double az = azimuth_from_xy(x, y);
    __syncthreads();
    if (!in_grid) return;
    if (P.nonfinite && !(s < __longlong_as_double(0x7ff0000000000000LL))) *P.nonfinite = 1;

    AzEntry* my_az = s_az + threadIdx.x;
    int* my_iaz = s_iaz + threadIdx.x;
    int* my_tag = s_tag + threadIdx.x;
    const double2* __restrict__ azt = P.r2.az_t;
    if (LAZY) {
        my_tag[0] = -1; my_tag[T] = -1;
    } else {
        // the column's azimuth bracket of every sweep, three sweeps at a time: the guesses of a group and the table
        // entries they point at are requested together (one L2 round trip per group instead of one per sweep)
#pragma unroll 1
        for (int s0 = 0; s0 < ne; s0 += 3) {
            AzDesc d[3];
            int g[3];
            double2 lo[3];
            double hi[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                d[j] = s_ad[min(s0 + j, ne - 1)];
                g[j] = __double2int_rd(mul(sub(az, d[j].az_first), d[j].az_inv_step)) + 1;
                lo[j] = make_double2(0.0, 0.0);
                hi[j] = 0.0;
                if (g[j] >= 1 && g[j] < d[j].naz) {
                    const double2* Tt = azt + d[j].az_off + (unsigned)(g[j] - 1);
                    lo[j] = __ldg(Tt);
                    hi[j] = __ldg(&Tt[1].x);
                }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int sw = s0 + j;
                if (sw < ne) {
                    AzEntry e;
                    int rec;
                    r3_az_bracket_loaded(d[j], g[j], lo[j], hi[j], s_sw + sw, azt, az, e, rec);
                    my_az[(size_t)sw * T] = e;
                    if (EMIT) my_iaz[(size_t)sw * T] = rec;
                }
            }
        }
    }
    // LAZY: the brackets of sweeps az_k and min(az_k + 1, ne - 1), az_k = -2: none yet
    AzEntry za, zb;
    int rec_a = -1, rec_b = -1, az_k = -2;
    za.row = 0u; za.w = 0.f; zb = za;

    float fill32 = p.fill32, inv2a = P.r2.inv2a;
    float u_inv_step = P.r2.u_inv_step, u_c0 = P.r2.u_c0, u_step = P.r2.u_step, u_r_first = P.r2.u_r_first;
    int u_ngm1 = P.r2.u_ngm1;
    unsigned u_rng_off = P.r2.u_rng_off;
    const double2* g_t = P.r2.g_t;
    const float4* val0 = static_cast<const float4*>(p.vol.val[0]);
    const bool lowest_sweep = p.lowest_sweep != 0, highest_sweep = p.highest_sweep != 0;
    // blind-zone rules (RadarGridC.pyx:424-433) as one window test on k: sampled iff k_first <= k <= k_last
    int k_first = lowest_sweep ? -1 : 0;
    unsigned k_span = (unsigned)((highest_sweep ? ne - 1 : ne - 2) - k_first);
    unsigned row_bytes = (unsigned)(2 * ne + 3) * 32u;     // one level row: [LevelPair x (ne+1)][LevelSweep x (ne+2)]
    unsigned ls_off = (unsigned)(ne + 1) * 32u;            // the LevelSweep half of a row
    unsigned pair_top = (unsigned)(ne - 1);                // k < pair_top (unsigned): between two sweeps
    unsigned az_addr = (unsigned)__cvta_generic_to_shared(my_az), az_pitch = T * (unsigned)sizeof(AzEntry);
    unsigned fs_addr = (unsigned)__cvta_generic_to_shared(&s_fs[0].q1);
    unsigned ad_addr = (unsigned)__cvta_generic_to_shared(s_ad);
    {
        const int pin_zero = (int)((unsigned long long)clock64() >> 63);     // 0, unknown to the compiler
        (void)pin_zero;
        R3_PIN(fill32); R3_PIN(inv2a); R3_PIN(g_t); R3_PIN(val0);
        R3_PIN(k_first); R3_PIN(k_span); R3_PIN(row_bytes); R3_PIN(ls_off); R3_PIN(pair_top);
        R3_PIN(az_addr); R3_PIN(az_pitch); R3_PIN(fs_addr);
        if (LAZY) R3_PIN(ad_addr);
        if (UNI) { R3_PIN(u_inv_step); R3_PIN(u_c0); R3_PIN(u_step); R3_PIN(u_r_first); R3_PIN(u_ngm1); R3_PIN(u_rng_off); }
    }
    const char* lrow = reinterpret_cast<const char*>(P.r2.lp);   // this level's row (block-uniform), global copy
    unsigned tab_off = 0;                                          // the same row inside Q.tab (CT)
    // one 16-byte half of a 32-byte record of this level's row
    auto row16 = [&](unsigned byte_off) -> int4 {
        if (CT) {                                                  // two 64-bit constant-bank loads
            const unsigned w = (tab_off + byte_off) >> 3;
            const double a = Q.tab[w], b = Q.tab[w + 1];
            return make_int4(__double2loint(a), __double2hiint(a), __double2loint(b), __double2hiint(b));
        }
        return __ldg(reinterpret_cast<const int4*>(lrow + byte_off));
    };
    auto as_d2 = [](const int4& v) -> double2 {
        return make_double2(__hiloint2double(v.y, v.x), __hiloint2double(v.w, v.z));
    };
    int k = -1, k6 = 0;
    char* out_b = reinterpret_cast<char*>(p.out + col);
    const size_t lvl_bytes = (size_t)ncol * sizeof(double);
    // levels whose voxel needs the reference chain are only MARKED here and resolved after the loop: the hot
    // loop then contains no call, so its state stays in registers (kMaxLevelsPerLaunch <= 64)
    unsigned long long slow_mask = 0ull;
#ifdef R3_SPEC
    constexpr bool SPEC = UNI && !LAZY;   // uniform-table volumes: every speculative address below is valid
#else
    constexpr bool SPEC = false;    // measured on B200, config C3: 0.386 ms with, 0.344 ms without
#endif

#ifndef R3_NOUNROLL2
#pragma unroll(LAZY ? 1 : 2)        // (0.3334 -> 0.3319 ms at config C3)
#else
#pragma unroll 1
#endif
    for (int iz = 0; iz < p.nz; ++iz, lrow += row_bytes, tab_off += row_bytes, out_b += lvl_bytes) {
        float res[NF];
        int st_state = kEmpty, st_lower = -1, st_upper = -1, r_iaz0 = -1, r_ir0 = -1, r_fl0 = -1, r_iaz1 = -1, r_ir1 = -1,
            r_fl1 = -1;
        // RadarGridC.pyx:146-149 — the reference's own operations: r2 is bit-identical
        const double r2 = sub(p.levels[iz].a2g2, mul(p.levels[iz].twoag, cphi));
        const float r2f = (float)r2;
        const float rinv = rsqrt_approx(r2f);
        const float rf = r2f * rinv;
        bool slow = false;
        // SPEC: everything a voxel of the CURRENT sweep pair needs is requested before anything is compared —
        // confirmation entry, pair record, the gate band of the guessed gate, both azimuth entries, the quads —
        // so a level costs one memory round trip instead of three dependent ones; a column that leaves its pair
        // (19 % of the voxels at config C3) walks to the new pair and asks again.
        int4 lk;
        double2 tg = make_double2(0.0, 0.0);
        int gi = 1;
        AzEntry ea, eb;
        float4 qa4[NF], qb4[NF];
        ea.row = 0u; ea.w = 0.f; eb = ea;
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) { qa4[fi] = make_float4(0.f, 0.f, 0.f, 0.f); qb4[fi] = qa4[fi]; }
        if (!SPEC && !CT) {
            // rows in global memory: ONE block with a single exit and a warp barrier behind it, so that the lanes that
            // walked and the lanes whose pair was confirmed at once run the rest of the level together (the nested
            // loops below left them apart until the end of the level: 1.39 executions of the sampling code per
            // voxel-warp at config C5).  The band
            // test that ends the walk IS the confirmation of pair k ({t_lo[k], t_hi[k+1]}); its {rho2, rho, b1} is the
            // second half of sweep k's own record.
            const unsigned pk = (unsigned)(k + 1) * 32u;
            const double2 cf = as_d2(row16(pk));
            lk = row16(pk + 16u);
            if (!((r2 < cf.x) & (r2 > cf.y))) {
#pragma unroll 1
                for (;;) {
                    const unsigned ro = ls_off + (unsigned)(k + 1) * 32u;
                    const double2 cur = as_d2(row16(ro));
                    const double2 nxt = as_d2(row16(ro + 32u));
                    if ((r2 < cur.x) & (r2 > nxt.y)) { lk = row16(ro + 16u); break; }
                    if (r2 > cur.y) { --k; continue; }
                    if (r2 < nxt.x) { ++k; continue; }
                    slow = true;
                    break;
                }
            }
            // (ptxas puts no convergence barrier behind this block on its own; every lane that has not left the
            // kernel runs every level, so the full mask is exact)
            __syncwarp();
        } else {
#pragma unroll 1
            for (;;) {
                const unsigned pk = (unsigned)(k + 1) * 32u;
                const double2 cf = as_d2(row16(pk));
                lk = row16(pk + 16u);                                 // {rho2, rho, b1} of pair k: used if k is confirmed
                if (SPEC) {
                    const int lo_s = max(k, 0), hi_s = min(lo_s + 1, ne - 1);
                    gi = min(max(__float2int_rd(fmaf(rf, u_inv_step, u_c0)), 1), u_ngm1);
                    tg = ldg_vol(g_t + (u_rng_off + (unsigned)(gi - 1)));
                    ea = lds_az(az_addr + (unsigned)lo_s * az_pitch);
                    eb = lds_az(az_addr + (unsigned)hi_s * az_pitch);
#pragma unroll
                    for (int fi = 0; fi < NF; ++fi) {
                        const float4* v = fi == 0 ? val0 : static_cast<const float4*>(p.vol.val[fi]);
                        qa4[fi] = ldg_vol(v + (ea.row + (unsigned)gi));
                        qb4[fi] = ldg_vol(v + (eb.row + (unsigned)gi));
                    }
                }
                if ((r2 < cf.x) & (r2 > cf.y)) break;
                // the column left its sweep pair: walk the per-sweep bands (pcg_r2.cuh)
#pragma unroll 1
                for (;;) {
                    const unsigned ro = ls_off + (unsigned)(k + 1) * 32u;
                    const double2 cur = as_d2(row16(ro));
                    const double2 nxt = as_d2(row16(ro + 32u));
                    if ((r2 < cur.x) & (r2 > nxt.y)) break;
                    if (r2 > cur.y) { --k; continue; }
                    if (r2 < nxt.x) { ++k; continue; }
                    slow = true;
                    break;
                }
                if (slow) break;
                // (rows in the constant bank: the second look at the pair entry is two broadcasts and keeps the pair
                // record's loads inside the pair branch — config C3: 0.332 ms with it, 0.339 ms without)
            }
        }
        if (slow) { slow_mask |= 1ull << iz; continue; }
        const bool pair = (unsigned)k < pair_top;
        const bool skip = (unsigned)(k - k_first) > k_span;
        const int lower = max(k, 0);
        if (EMIT) {
            st_state = skip ? (pair ? kPair : (k < 0 ? kSkipLow : kSkipHigh)) : (pair ? kPair : kSingle);
            st_lower = lower; st_upper = lower + (pair ? 1 : 0);
        }
        if (skip) {
#pragma unroll
            for (int f = 0; f < NF; ++f) res[f] = fill32;
        } else {
            float rj, idr;
            if (LAZY && lower != az_k) {
                // the column moved to another sweep pair: one step up keeps the upper bracket as the new lower one
                const int top = min(lower + 1, (int)pair_top);
                int sw = lower;
                if (lower == az_k + 1) { za = zb; rec_a = rec_b; sw = top; }
                az_k = lower;
#pragma unroll 1
                for (;;) {
                    AzEntry e;
                    int rec;
                    AzDesc d;
                    {
                        const unsigned da = ad_addr + (unsigned)sw * (unsigned)sizeof(AzDesc);
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(d.az_first), "=d"(d.az_inv_step) : "r"(da) : "memory");
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(d.az_off), "=r"(d.naz), "=r"(d.val_off), "=r"(d.ng) : "r"(da + 16u) : "memory");
                    }
                    r3_az_bracket(d, s_sw + sw, azt, az, e, rec);
                    if (sw == lower) { za = e; rec_a = rec; }
                    if (sw == top) { zb = e; rec_b = rec; break; }
                    sw = top;
                }
            }
            if (SPEC) {
                if (!((r2 >= tg.x) & (r2 < tg.y))) { slow_mask |= 1ull << iz; continue; }
                rj = fmaf((float)(gi - 1), u_step, u_r_first); idr = u_inv_step;
            } else {
                // ---- bracket, then gather (volumes with ragged / irregular / degenerate tables) ----
                const FastSweep* f = s_fs + lower;
                unsigned rng_off = u_rng_off;
                int ngm1 = u_ngm1;
                float inv_step = u_inv_step, c0 = u_c0;
                bool usable = true;
                if (!UNI) {
                    const int4 tab = *reinterpret_cast<const int4*>(&f->rng_off);
                    rng_off = (unsigned)tab.x; ngm1 = tab.y; c0 = __int_as_float(tab.z); inv_step = __int_as_float(tab.w);
                    const int need = pair ? (kFsPolyOk | kFsShared | kFsUsable) : kFsUsable;
                    usable = (f->flags & need) == need;
                }
                gi = usable ? min(max(__float2int_rd(fmaf(rf, inv_step, c0)), 1), ngm1) : 1;
                const unsigned gidx = usable ? rng_off + (unsigned)(gi - 1) : u_rng_off;
                tg = ldg_vol(g_t + gidx);                          // r in [R[gi-1], R[gi])  <=>  tg.x <= r2 < tg.y
#ifndef R3_NOGSPEC
                if (UNI) {
                    // uniform tables: the guessed gate's quads are asked for together with its band, before the band
                    // is checked (the guess is right for all but ~1e-4 of the voxels; every address is valid): one
                    // dependent memory round trip less per level, 0.342 -> 0.333 ms at config C3
                    ea = LAZY ? za : lds_az(az_addr + (unsigned)lower * az_pitch);
                    if (pair) eb = LAZY ? zb : lds_az(az_addr + (unsigned)(lower + 1) * az_pitch);
#pragma unroll
                    for (int fi = 0; fi < NF; ++fi) {
                        const float4* v = fi == 0 ? val0 : static_cast<const float4*>(p.vol.val[fi]);
                        qa4[fi] = ldg_vol(v + (ea.row + (unsigned)gi));
                        if (pair) qb4[fi] = ldg_vol(v + (eb.row + (unsigned)gi));
                    }
                }
#endif
                if (!(usable & (r2 >= tg.x) & (r2 < tg.y))) {
                    // beyond the last gate of the (shared) range table: every sample is missing, and so is the voxel
                    // (RadarGridC.pyx:275-276,246-251).  Index records of such voxels come from the reference chain.
                    if (!EMIT && usable && r2 >= s_fs[lower].tg_gt_last) {
                        if (!MERGE) {
#pragma unroll
                            for (int f2 = 0; f2 < NF; ++f2)
                                reinterpret_cast<double*>(out_b)[(size_t)f2 * (size_t)p.field_stride] = (double)fill32;
                        }
                    } else {
                        slow_mask |= 1ull << iz;
                    }
                    continue;
                }
                if (UNI) {
                    rj = fmaf((float)(gi - 1), u_step, u_r_first); idr = u_inv_step;
                } else {
                    const int4 hv = __ldg(reinterpret_cast<const int4*>(P.r2.h_t + gidx));   // {R^2, R, 1/dR}
                    rj = __int_as_float(hv.z); idr = __int_as_float(hv.w);
                }
#ifndef R3_NOGSPEC
                if (!UNI)
#endif
                {
                    ea = LAZY ? za : lds_az(az_addr + (unsigned)lower * az_pitch);
                    if (pair) eb = LAZY ? zb : lds_az(az_addr + (unsigned)(lower + 1) * az_pitch);
#pragma unroll
                    for (int fi = 0; fi < NF; ++fi) {
                        const float4* v = fi == 0 ? val0 : static_cast<const float4*>(p.vol.val[fi]);
                        qa4[fi] = ldg_vol(v + (ea.row + (unsigned)gi));
                        if (pair) qb4[fi] = ldg_vol(v + (eb.row + (unsigned)gi));
                    }
                }
            }
            // ---- the two weights and the lerps (RadarGridC.pyx:297-299,469-476) ----
            // tg[j] is R[j]^2 to within two ulps: w_r = (r^2 - R[j]^2) / ((r + R[j]) dR)
            const float w_r = (float)sub(r2, tg.x) * idr * rcp_approx(rf + rj);
            if (pair) {
                const float u = (float)sub(r2, __hiloint2double(lk.y, lk.x)) * rcp_approx(rf + __int_as_float(lk.z)) *
                                fmaf(__int_as_float(lk.w), rinv, inv2a);
                float4 qa;                                    // q1 .. q4
                float q5;
                if (CT) {
                    qa = *reinterpret_cast<const float4*>(&Q.poly[lower][0]);
                    q5 = Q.poly[lower][4];
                } else {
                    const unsigned fa = fs_addr + (unsigned)lower * (unsigned)sizeof(FastSweep);
                    qa = lds_f4(fa);
                    q5 = lds_f(fa + 16u);
                }
                float w_el = fmaf(q5, u, qa.w);
                w_el = fmaf(w_el, u, qa.z);
                w_el = fmaf(w_el, u, qa.y);
                w_el = fmaf(w_el, u, qa.x);
                w_el *= u;
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const float ea0 = fmaf(ea.w, qa4[fi].y - qa4[fi].x, qa4[fi].x), ea1 = fmaf(ea.w, qa4[fi].w - qa4[fi].z, qa4[fi].z);
                    const float eb0 = fmaf(eb.w, qb4[fi].y - qb4[fi].x, qb4[fi].x), eb1 = fmaf(eb.w, qb4[fi].w - qb4[fi].z, qb4[fi].z);
                    const float v0 = fmaf(w_r, ea1 - ea0, ea0);
                    const float v1 = fmaf(w_r, eb1 - eb0, eb0);
                    res[fi] = lerp_fill(w_el, v0, v1, fill32);
                }
            } else {
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const float ea0 = fmaf(ea.w, qa4[fi].y - qa4[fi].x, qa4[fi].x), ea1 = fmaf(ea.w, qa4[fi].w - qa4[fi].z, qa4[fi].z);
                    res[fi] = fmaf(w_r, ea1 - ea0, ea0);
                }
            }
            if (EMIT) {
                const int ra = LAZY ? rec_a : my_iaz[(size_t)lower * T];
                r_ir0 = gi; r_iaz0 = ra & 0x3fffffff; r_fl0 = (ra >> 30) & 1;
                if (pair) {
                    const int rb2 = LAZY ? rec_b : my_iaz[(size_t)(lower + 1) * T];
                    r_ir1 = gi; r_iaz1 = rb2 & 0x3fffffff; r_fl1 = (rb2 >> 30) & 1;
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double* dst = reinterpret_cast<double*>(out_b) + (size_t)f * (size_t)p.field_stride;
            // the fill value is float-representable (checked on the host), so widening res restores it
            const double v = (double)res[f];
            if (MERGE) {
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
            const bool have0 = (st_state == kSingle || st_state == kPair), have1 = (st_state == kPair);
            rec[0] = st_state;
            rec[1] = have0 ? st_lower : -1; rec[2] = have0 ? st_upper : -1;
            rec[3] = have0 ? r_iaz0 : -1; rec[4] = have0 ? r_ir0 : -1; rec[5] = have0 ? r_fl0 : -1;
            rec[6] = have1 ? r_iaz1 : -1; rec[7] = have1 ? r_ir1 : -1; rec[8] = have1 ? r_fl1 : -1;
            rec[9] = -1;
        }
    }

    // ---- the marked voxels: the reference's own chain (guard band, flagged level, r2 <= 0, ragged or degenerate
    //      tables, range gating / clamping, gate guess off), pcg_fast.cuh fast_voxel ----
    if (slow_mask) {
        SlowArgs sa;
        sa.s_sw = s_sw; sa.s_pr = s_pr; sa.my_az = my_az; sa.my_iaz = EMIT ? my_iaz : nullptr;
        sa.rg_t = reinterpret_cast<const double2*>(p.vol.rng); sa.val = p.vol.val; sa.az_stride = T; sa.ne = ne;
        sa.a2 = p.a2; sa.twoa = p.twoa; sa.guard = p.guard; sa.inv_twoa = p.inv_twoa; sa.fill32 = fill32;
        sa.blind = (p.nearest_gate != 0 ? 1 : 0) | (lowest_sweep ? 2 : 0) | (highest_sweep ? 4 : 0);
        sa.my_tag = LAZY ? my_tag : nullptr; sa.azt = azt; sa.az = az;
#pragma unroll 1
        while (slow_mask) {
            const int iz = __ffsll((long long)slow_mask) - 1;
            slow_mask &= slow_mask - 1;
            VoxelInfo vi;
            float tmp[NF];
            r3_slow_voxel<NF, EMIT, LAZY>(&sa, &p.levels[iz], cphi, &k6, tmp, &vi);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                double* dst = p.out + (size_t)f * (size_t)p.field_stride + (size_t)iz * (size_t)ncol + (size_t)col;
                const double v = (double)tmp[f];
                if (MERGE) {
                    if (tmp[f] != fill32) {
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
        }
    }
}

}  // namespace pcg
