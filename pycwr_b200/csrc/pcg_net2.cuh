// pcg_net2.cuh — the multi-radar merge of run_radar_network_3d as ONE tile-persistent kernel.
//
// Replaces, fused per output tile (pycwr/interp/RadarInterp.py):
//   grid_single_radar_to_latlon_3d :963-1023, _compute_quality_weight_volume :603-696, _compute_observability_mask
//   :540-600, _build_composite_weight_cache :1379-1396, compose_network_volume :1399-1487, the CR max-merge
//   :1026-1068,2047-2063, coverage_count / blind_mask :1938-1945 and build_network_dataset's fill -> NaN :1539.
//
// Round 1 ran one pass per radar over that radar's tile list with the per-voxel accumulators in HBM (pcg_network.cuh):
// a memset, a read-modify-write of a 16-byte cell per (voxel, reaching radar) and a finalize pass — about 32 GB of
// DRAM traffic for 4.7 GB of algorithmic bytes at config C4 — plus a cull kernel and a host sync for the tile counts.
// Here a block OWNS a tile of 128 grid columns (4 warps x (4 rows x 8 columns)) for the whole call:
//   1. it tests every radar's reach disc against the tile's bounding box (what net_cull_kernel did, no lists in HBM),
//   2. loops over the reaching radars IN RADAR ORDER (the sums accumulate in the reference's order, np.sum over
//      axis 0) — per (column, radar): column geometry, azimuth table in shared memory, every level through the
//      r2-space voxel path of pcg_r3.cuh,
//   3. keeps the tile's accumulators (sum w v, sum w as float32, two uint8 counters: 10 bytes per voxel) in shared
//      memory, each thread touching only its own column's cells (no barriers in the radar loop),
//   4. writes composite (with the caller's missing-value marker) / valid_count / coverage_count / blind_mask / CR
//      once.  No memset, no accumulator traffic, no finalize pass, no relabel pass, no host synchronisation.
// DRAM traffic is then the radar volumes (each streamed about once: tiles run in panels, so the resident blocks share
// radars), the grids and the outputs: 6.3 GB for 4.7 GB of algorithmic bytes at config C4.
//
// What the kernel waits for is not DRAM but dependent round trips to L2: with 4 x 53 KB of an SM's 256 KB configured as
// shared memory, ~28 KB of L1 are left and every small per-radar table comes from L2 (~300 cycles).  Hence:
//   * each warp keeps its own copy of the current radar's sweep descriptors, pair table, level constants and the
//     confirmation band of "below the lowest sweep" in shared memory, fetched by one lane per entry in one round trip;
//   * the azimuth brackets and the CR sweeps of a (column, radar) are taken three at a time, their table entries /
//     gate bands / quads requested together; the CR gate bracket is made in r2 space like the level loop's;
//   * a sample that can reach no requested output is not taken (quality-weighted composite without valid_count:
//     the one-sweep fill-in below the lowest beam has weight zero), and a column still below the lowest beam is
//     recognised from the warp's copy without any global access.
// 11.9 ms (round 1) -> 6.97 ms on one B200 at config C4; measured steps and rejected variants: DESIGN.md section 4.2.
#pragma once
#include "pcg_network.cuh"
#include "pcg_r3.cuh"

namespace pcg {

constexpr int kNet2Threads = 128;
#ifndef NET2_GRP
#define NET2_GRP 3                 // sweeps per group in the azimuth-bracket and CR loops (their loads go out together)
#endif
#ifndef NET2_PANEL
#define NET2_PANEL 8
#endif
// Tiles are numbered in PANELS of NET2_PANEL tiles across y, x fastest inside a panel, so that the ~600 blocks
// resident at any time cover a compact ~300 x 256-column patch of the grid instead of a 25-row band across all of it:
// they then gather from the same three or four radar volumes (L2-resident together) rather than from every radar the
// band crosses.  NET2_PANEL = 0: row-band order.
__device__ __forceinline__ void net2_tile_column(const NetArgs& p, int tile, bool& active, long long& col) {
    const int pyl = p.py_log2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int ly = lane & ((1 << pyl) - 1), lx = lane >> pyl;
    int by, bx;
    if (NET2_PANEL > 0) {
        const int tiles_x = (p.nx + (32 >> pyl) - 1) / (32 >> pyl);
        const int per_panel = NET2_PANEL * tiles_x;
        const int panel = tile / per_panel, r = tile - panel * per_panel;
        const int width = min(NET2_PANEL, p.tiles_y - panel * NET2_PANEL);
        bx = r / width;
        by = panel * NET2_PANEL + (r - bx * width);
    } else {
        by = tile % p.tiles_y; bx = tile / p.tiles_y;
    }
    int iy = ((by * nwarp + warp) << pyl) + ly;
    int ix = bx * (32 >> pyl) + lx;
    active = (ix < p.nx) & (iy < p.ny);
    ix = min(ix, p.nx - 1); iy = min(iy, p.ny - 1);
    col = (long long)ix * p.ny + iy;
}
#ifndef NET2_MINB
#define NET2_MINB 4
#endif
#ifndef NET2_SMEM_KB
#define NET2_SMEM_KB 55
#endif
constexpr int kNet2SmemKB = NET2_SMEM_KB;   // shared-memory budget of a block: levels per launch follow from it

// one warp's copy of a radar's per-sweep descriptors: AzDesc, pair table entry, tan(elevation); 16-byte granular
__host__ __device__ inline size_t net2_desc_bytes(int ne_max) {
    return ((size_t)ne_max * (sizeof(AzDesc) + sizeof(NetPairQ) + sizeof(double)) + 15) & ~(size_t)15;
}

// dynamic shared memory of net_tile_kernel
__host__ __device__ inline size_t net2_smem_bytes(int ne_max, int nz, bool nearest) {
    size_t b = 512;                                                    // bbox partials, hit flags, radar list
    b += (size_t)ne_max * kNet2Threads * sizeof(AzEntry);              // azimuth table [sweep][thread]
    b += (size_t)nz * kNet2Threads * (2 * sizeof(float) + sizeof(unsigned short));
    b = (b + 15) & ~(size_t)15;
    if (nearest) b += (size_t)nz * kNet2Threads * sizeof(double);
    b += (size_t)(kNet2Threads / 32) * net2_desc_bytes(ne_max);        // per-warp sweep descriptors
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(kNet2Threads / 32) * nz * 2 * sizeof(double2);       // per-warp level constants + lowest-sweep band
    return (b + 15) & ~(size_t)15;
}

// observability of a voxel in the first / last gate interval of its sweep pair, or flagged by the voxel path:
// the guard bands of pcg_network.cuh, then the NumPy-twin chain itself (out of line, rare)
__device__ __noinline__ bool net2_obs_edge(double r2, const NetPair* np, double x, double y, const NetRadar* R,
                                           const NetLevel* L, bool tie, unsigned long long* ties) {
    const bool in = (r2 > np->of2_hi) & (r2 < np->ol2_lo);
    const bool sure_out = (r2 < np->of2_lo) | (r2 > np->ol2_hi);
    if (!tie && (in | sure_out) && (r2 > 0.0)) return in;
    ++*ties;
    return twin_observable(x, y, *R, *L, R->vol.sweeps, R->npairs, R->ne);
}

__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f2(unsigned addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

// one sweep's CR sample by the reference chain (RadarGridC.pyx:596-624 through range_bracket / sample_pairs): the
// exact square root, the bracket verified against the range table itself (rare: the r2-space guess was not confirmed;
// volumes without one shared regular range table)
__device__ __noinline__ float net2_cr_exact(const SweepDesc* d, const double2* rg_t, const float4* val0,
                                            const AzEntry* az, double r2, float fill32) {
    const double r = sqrt_rn(r2);
    if (r != r) return fill32;
    const RangeBracket rb = range_bracket(*d, rg_t, r, false);
    return sample_pairs(val0, *az, rb, fill32);
}

template <int METHOD, bool UNI>
__global__ void __launch_bounds__(kNet2Threads, NET2_MINB) net_tile_kernel(const __grid_constant__ NetArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned T = kNet2Threads;
    double* s_red = reinterpret_cast<double*>(smem_raw);                         // [4 x 4 warps] = 128 B
    int* s_nlist = reinterpret_cast<int*>(smem_raw + 128);
    unsigned char* s_list = smem_raw + 256;                                        // [256] hit flags, then the list
    AzEntry* s_az = reinterpret_cast<AzEntry*>(smem_raw + 512);                    // [ne_max][T]
    float2* a_acc = reinterpret_cast<float2*>(s_az + (size_t)p.ne_max * T);        // [nz][T]: {sum w v, sum w}
    unsigned short* a_cc = reinterpret_cast<unsigned short*>(a_acc + (size_t)p.nz * T);   // [nz][T]: cnt | cov << 8
    double* a_d2 = reinterpret_cast<double*>(smem_raw + ((512 + (size_t)p.ne_max * T * sizeof(AzEntry) +
                                             (size_t)p.nz * T * 10 + 15) & ~(size_t)15));   // nearest only
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per-warp copies of the current radar's sweep descriptors (what the azimuth brackets and the CR pass read):
    // fetched by one lane per sweep in ONE round trip instead of one dependent L2 access per sweep and loop
    AzDesc* w_ad;
    float4* w_pq;                   // [ne][2]: the pair table of the radar (polynomial + quality terms)
    double* w_tan;
    double2* w_lev;                 // [nz][2]: {a2g2, twoag}, {t_lo, t_hi} of the "below the lowest sweep" record
    {
        const size_t acc_end = ((512 + (size_t)p.ne_max * T * sizeof(AzEntry) + (size_t)p.nz * T * 10 + 15) & ~(size_t)15) +
                               (METHOD == kNetNearest ? (size_t)p.nz * T * sizeof(double) : 0);
        const size_t desc = net2_desc_bytes(p.ne_max);
        unsigned char* base = smem_raw + acc_end + (size_t)warp * desc;
        w_ad = reinterpret_cast<AzDesc*>(base);
        w_pq = reinterpret_cast<float4*>(base + (size_t)p.ne_max * sizeof(AzDesc));
        w_tan = reinterpret_cast<double*>(base + (size_t)p.ne_max * (sizeof(AzDesc) + sizeof(NetPairQ)));
        const size_t lev0 = (acc_end + (T >> 5) * desc + 15) & ~(size_t)15;
        w_lev = reinterpret_cast<double2*>(smem_raw + lev0) + (size_t)warp * p.nz * 2;
    }

    bool active;
    long long col;
    net2_tile_column(p, blockIdx.x, active, col);
    const double gx = p.gx[col], gy = p.gy[col];            // inactive lanes repeat a valid column
    // ---- 1. which radars reach this tile (bounding box against reach disc, as net_cull_kernel) ----
    {
        double x0 = gx, x1 = gx, y0 = gy, y1 = gy;
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
            x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0) { s_red[4 * warp] = x0; s_red[4 * warp + 1] = x1; s_red[4 * warp + 2] = y0; s_red[4 * warp + 3] = y1; }
        // NaN coordinates poison fmin / fmax silently: a tile holding one is handed to every radar
        const bool odd = __syncthreads_or(!(gx == gx) || !(gy == gy));
        for (int w = 0; w < (int)(T >> 5); ++w) {
            x0 = fmin(x0, s_red[4 * w]); x1 = fmax(x1, s_red[4 * w + 1]);
            y0 = fmin(y0, s_red[4 * w + 2]); y1 = fmax(y1, s_red[4 * w + 3]);
        }
        for (int ir = tid; ir < p.nradar; ir += T) {
            const NetRadar& R = p.radars[ir];
            const double ddx = fmax(fmax(x0 - R.x, R.x - x1), 0.0);
            const double ddy = fmax(fmax(y0 - R.y, R.y - y1), 0.0);
            s_list[ir] = (odd || !(ddx * ddx + ddy * ddy > R.cull2)) ? 1 : 0;
        }
        __syncthreads();
        if (warp == 0) {                                      // ordered compaction: the list is in radar order
            int base = 0;
            for (int c = 0; c < p.nradar; c += 32) {
                const int ir = c + lane;
                const bool hit = ir < p.nradar && s_list[ir] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                __syncwarp();
                if (hit) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned char)ir;   // base + rank <= ir: in place
                base += __popc(m);
                __syncwarp();
            }
            if (lane == 0) *s_nlist = base;
        }
    }
    // ---- 3. this column's accumulators ----
    const int nz = p.nz;
    for (int iz = 0; iz < nz; ++iz) {
        a_acc[(size_t)iz * T + tid] = make_float2(0.f, 0.f);
        a_cc[(size_t)iz * T + tid] = 0;
        if (METHOD == kNetNearest) a_d2[(size_t)iz * T + tid] = __longlong_as_double(0x7ff0000000000000LL);
    }
    __syncthreads();
    const int nlist = *s_nlist;
    float cr_best = __int_as_float(0x7fc00000);              // NaN: no echo yet
    unsigned long long ties = 0;
    const float fill32 = p.fill32;
    AzEntry* my_az = s_az + tid;
    const unsigned az_addr = (unsigned)__cvta_generic_to_shared(my_az);
    constexpr unsigned az_pitch = T * (unsigned)sizeof(AzEntry);
    const unsigned acc_addr = (unsigned)__cvta_generic_to_shared(a_acc + tid);
    const unsigned cc_addr = (unsigned)__cvta_generic_to_shared(a_cc + tid);
#ifndef NET2_NOSPEC
    constexpr bool SPEC = UNI;      // uniform-table volumes: every speculative address of the level loop is valid
#else
    constexpr bool SPEC = false;
#endif

    // The 2-D weights exp(-4 d^2 / R_inf^2) (RadarInterp.py:1389) of a column are taken RELATIVE to its nearest
    // reaching radar: exp(-4 (d^2 - d_min^2) / R_inf^2).  Numerator and denominator of the weighted mean scale by
    // the same factor, so the quotient is the reference's; but the float32 sums no longer underflow where every
    // reaching radar is far compared with the influence radius (the reference accumulates in float64: a voxel seen
    // only from 4.7 influence radii away has a value there and had none here).
    double d2_min = 0.0;
    if (METHOD == kNetExp || METHOD == kNetQuality) {
        d2_min = __longlong_as_double(0x7ff0000000000000LL);
        for (int li = 0; li < nlist; ++li) {
            const NetRadar& R = p.radars[s_list[li]];
            const double x = sub(gx, R.x), y = sub(gy, R.y);
            const double d2 = add(mul(x, x), mul(y, y));
            if ((d2 <= R.cull2) & (d2 < d2_min)) d2_min = d2;
        }
    }

    // Dead work: in the quality-weighted composite a radar that does not OBSERVE a voxel has weight zero there, so its
    // sample can only show in valid_count.  When the caller does not ask for valid_count (run_radar_network_3d
    // never does: `composite, _ = compose_network_volume(...)`, RadarInterp.py:1957,1990) the voxels a radar merely
    // fills in from its lowest sweep ('lowest_sweep' / 'hybrid' blind methods: 62 % of the sampled voxels at config
    // C4) are not sampled at all — composite, coverage_count, blind_mask and CR are bit for bit what they would be.
    const bool dead_unobserved = (METHOD == kNetQuality) && p.quality != 0 && p.valid_count == nullptr;

    // ---- 2. the reaching radars, in radar order ----
#pragma unroll 1
    for (int li = 0; li < nlist; ++li) {
        const int ir = s_list[li];
        const NetRadar& R = p.radars[ir];
        const int ne = R.ne;
        {
            // this radar's sweep descriptors for the warp (every lane takes part, whether its column is in reach or not)
            __syncwarp();
            const SweepDesc* __restrict__ sw_t = R.vol.sweeps;
            for (int sw = lane; sw < ne; sw += 32) {
                const SweepDesc* d = sw_t + sw;
                AzDesc a;
                a.az_first = d->az_first; a.az_inv_step = d->az_inv_step; a.az_off = d->az_off;
                a.naz = (d->naz >= 2 && d->ng >= 2) ? d->naz : 0;                  // 0: degenerate, general path
                a.val_off = (unsigned)d->val_off; a.ng = d->ng;
                w_ad[sw] = a;
                w_tan[sw] = d->tan_e;
                const float4* pq = reinterpret_cast<const float4*>(R.pq + sw);
                w_pq[2 * sw] = __ldg(pq);
                w_pq[2 * sw + 1] = __ldg(pq + 1);
            }
            // ... and its level constants with the confirmation band of "below the lowest sweep" (record 0 of a
            // level's row): a column climbing towards the lowest beam is then recognised without a global access
            const unsigned rb = (unsigned)(2 * ne + 3) * 32u;
            const char* rows = reinterpret_cast<const char*>(R.r2.lp) + (size_t)p.z0 * rb;
            const NetLevel* lvl = R.levels + p.z0;
            for (int iz = lane; iz < p.nz; iz += 32) {
                w_lev[2 * iz] = make_double2(__ldg(&lvl[iz].cy.a2g2), __ldg(&lvl[iz].cy.twoag));
                w_lev[2 * iz + 1] = __ldg(reinterpret_cast<const double2*>(rows + (size_t)iz * rb));
            }
            __syncwarp();
        }
        // radar-local coordinates (RadarInterp.py:990-991) and the 2-D distance of the weight cache
        const double x = sub(gx, R.x), y = sub(gy, R.y);
        const double d2 = add(mul(x, x), mul(y, y));
        if (!active || !(d2 <= R.cull2)) continue;            // beyond every gate: nothing to add (NaN: chain below)
        // column invariants of this radar (RadarGridC.pyx:142-146,158)
        const double s = sqrt_rn(d2);
        const double phi = div(s, R.Re);
        const double cphi = cos_small(phi);
        {
            // the column's azimuth bracket of every sweep (RadarGridC.pyx:282-289), three sweeps at a time: the table
            // entries the three guesses point at are requested together
            const double az = azimuth_from_xy(x, y);
            const double2* __restrict__ azt = R.r2.az_t;
#pragma unroll 1
            for (int sw0 = 0; sw0 < ne; sw0 += NET2_GRP) {
                AzDesc d[NET2_GRP];
                int g[NET2_GRP];
                double2 lo[NET2_GRP];
                double hi[NET2_GRP];
#pragma unroll
                for (int j = 0; j < NET2_GRP; ++j) {
                    const int sw = min(sw0 + j, ne - 1);
                    d[j] = w_ad[sw];
                    g[j] = __double2int_rd(mul(sub(az, d[j].az_first), d[j].az_inv_step)) + 1;
                    lo[j] = make_double2(0.0, 0.0); hi[j] = 0.0;
                    if (g[j] >= 1 && g[j] < d[j].naz) {
                        const double2* Tt = azt + d[j].az_off + (unsigned)(g[j] - 1);
                        lo[j] = __ldg(Tt);
                        hi[j] = __ldg(&Tt[1].x);
                    }
                }
#pragma unroll
                for (int j = 0; j < NET2_GRP; ++j) {
                    const int sw = sw0 + j;
                    if (sw >= ne) break;
                    AzEntry e;
                    int rec;
                    bool done = false;
                    int gg = g[j];
                    if (gg >= 1 && gg < d[j].naz) {
                        const double2* Tt = azt + d[j].az_off + (unsigned)(gg - 1);
                        double2 l = lo[j];
                        double h = hi[j];
                        if (az < l.x) {
                            if (gg >= 2) { --gg; h = l.x; l = __ldg(Tt - 1); }
                        } else if (!(az < h)) {
                            if (gg + 1 < d[j].naz) { ++gg; l = __ldg(Tt + 1); h = __ldg(&Tt[2].x); }
                        }
                        if (!(az < l.x) && (az < h)) {
                            e.row = d[j].val_off + (unsigned)(gg - 1) * (unsigned)d[j].ng;
                            e.w = (float)sub(az, l.x) * (float)l.y;
                            done = true;
                        }
                    }
                    if (!done) az_bracket_general(R.vol.sweeps + sw, azt, az, &e, &rec);
                    my_az[(size_t)sw * T] = e;
                }
            }
        }
        // exp(-4 d^2 / R_inf^2) (RadarInterp.py:1389)
        const float w2d = (METHOD == kNetExp || METHOD == kNetQuality) ? __expf(-(float)sub(d2, d2_min) * R.inv_r2inf4) : 1.f;
        const bool lowest_sweep = p.lowest_sweep != 0, highest_sweep = p.highest_sweep != 0;
        const int k_first = lowest_sweep ? -1 : 0;
        const unsigned k_span = (unsigned)((highest_sweep ? ne - 1 : ne - 2) - k_first);
        const unsigned row_bytes = (unsigned)(2 * ne + 3) * 32u;
        const unsigned ls_off = (unsigned)(ne + 1) * 32u;
        const unsigned pair_top = (unsigned)(ne - 1);
        const char* lrow = reinterpret_cast<const char*>(R.r2.lp) + (size_t)p.z0 * row_bytes;
        const NetLevel* lv = R.levels + p.z0;
        const float4* val0 = static_cast<const float4*>(R.vol.val[0]);
        const double2* g_t = R.r2.g_t;
        const FastSweep* fs_t = R.r2.fs;
        const NetPairQ* pq_t = R.pq;
        const NetPair* np_t = R.npairs;
        const float inv2a = R.r2.inv2a;
        const float u_inv_step = R.r2.u_inv_step, u_c0 = R.r2.u_c0, u_step = R.r2.u_step, u_r_first = R.r2.u_r_first;
        const int u_ngm1 = R.r2.u_ngm1;
        const unsigned u_rng_off = R.r2.u_rng_off;
        const float src_lo_f = R.src_lo_f, src_hi_f = R.src_hi_f;
        const float inv_decay = R.inv_decay;
        const bool sanity = p.sanity != 0;
        int k = -1;
        unsigned long long slow_mask = 0ull;

        // one voxel's contribution: range-sanity filter (RadarInterp.py:1003-1014, on the float value against the
        // float-rounded window: the same decisions as the reference's double comparison), observability weight
        // (:566-599 / 636-695), then the merge rule of compose_network_volume (:1448-1487)
        auto merge = [&](int iz, float v, bool obs, float rf, float beam_k, float vert_w) {
            bool valid = (v != fill32) & (fabsf(v) < __int_as_float(0x7f800000));        // finite and not fill
            if (sanity) valid = valid & (v >= src_lo_f) & (v <= src_hi_f);
            if (!(obs | valid)) return;
            const unsigned cc_a = cc_addr + (unsigned)iz * (T * 2u);
            unsigned cc = lds_u16(cc_a);
            if (obs) cc += 0x100u;
            if (valid) {
                cc += 1u;
                const unsigned ac = acc_addr + (unsigned)iz * (T * 8u);
                float2 a = lds_f2(ac);
                if (METHOD == kNetMax) {
                    if (a.y == 0.f || v > a.x) a.x = v;
                    a.y = 1.f;
                } else if (METHOD == kNetNearest) {
                    const size_t c = (size_t)iz * T + tid;
                    if (d2 < a_d2[c]) { a_d2[c] = d2; a.x = v; a.y = 1.f; }
                } else {
                    float w = w2d;
                    if (METHOD == kNetQuality && p.quality) {
                        float qw = 0.f;
                        if (obs) {
                            const float t = rf * inv_decay;
                            const float b = rf * beam_k;
                            qw = __expf(-t * t) * rcp_approx(fmaf(b, b, 1.f)) * vert_w;
                        }
                        w *= qw;
                    }
                    a.x = fmaf(w, v, a.x);
                    a.y += w;
                }
                sts_f2(ac, a);
            }
            sts_u16(cc_a, cc);
        };

#pragma unroll 1
        for (int iz = 0; iz < nz; ++iz, lrow += row_bytes) {
            // RadarGridC.pyx:146-149 — the reference's own operations: r2 is bit-identical
            // (NetLevel is 56 bytes: its members are 8-byte, not 16-byte, aligned)
            const double2 lc = w_lev[2 * iz];
            const double r2 = sub(lc.x, mul(lc.y, cphi));
            if (dead_unobserved && k == -1) {
                // still certainly below the lowest sweep (the confirmation test of pair -1, from the warp's copy):
                // a one-sweep fill-in with weight zero — nothing to load, nothing to add
                const double2 c0 = w_lev[2 * iz + 1];
                if ((r2 < c0.x) & (r2 > c0.y)) continue;
            }
            const float r2f = (float)r2;
            const float rinv = rsqrt_approx(r2f);
            const float rf = r2f * rinv;
            bool slow = false;
            // Everything a voxel of the CURRENT sweep pair needs is requested before anything is compared
            // (uniform-table volumes: every sweep usable, one shared range table, so every speculative address
            // is valid): the confirmation entry, the pair record, the gate band of the guessed gate, both
            // azimuth entries, both quads and the pair polynomial are in flight together — one memory round
            // trip per level instead of three dependent ones.  A column that leaves its pair (19 % of the
            // voxels at config C3) walks to the new pair and asks again.
            int4 lk = make_int4(0, 0, 0, 0);
            double2 tg = make_double2(0.0, 0.0);
            int gi = 1;
            AzEntry ea, eb;
            float4 qa4, qb4, q14, q58;           // q14 = {q1 .. q4}, q58 = {q5, beam_k, vert_w, -} of the pair
            ea.row = 0u; ea.w = 0.f; eb = ea;
            qa4 = make_float4(0.f, 0.f, 0.f, 0.f); qb4 = qa4; q14 = qa4; q58 = qa4;
#pragma unroll 1
            for (;;) {
                const int4* pk = reinterpret_cast<const int4*>(lrow + (unsigned)(k + 1) * 32u);
                const double2 cf = __ldg(reinterpret_cast<const double2*>(pk));
                if (SPEC) {
                    // what the GUESSED pair needs: a column below the lowest / above the highest sweep samples one
                    // sweep, so the second quad, the pair record and the pair polynomial are not even asked for
                    // (62 % of the sampled voxels at config C4: a third of the L1 wavefronts of the level)
                    const int lo_s = max(k, 0);
                    gi = min(max(__float2int_rd(fmaf(rf, u_inv_step, u_c0)), 1), u_ngm1);
                    if (!dead_unobserved || (unsigned)k < pair_top) {
                        tg = ldg_vol(g_t + (u_rng_off + (unsigned)(gi - 1)));
                        ea = lds_az(az_addr + (unsigned)lo_s * az_pitch);
                        qa4 = ldg_vol(val0 + (ea.row + (unsigned)gi));
                    }
                    if ((unsigned)k < pair_top) {
                        lk = __ldg(pk + 1);
                        eb = lds_az(az_addr + (unsigned)(lo_s + 1) * az_pitch);
                        qb4 = ldg_vol(val0 + (eb.row + (unsigned)gi));
                        q14 = w_pq[2 * lo_s];
                        q58 = w_pq[2 * lo_s + 1];
                    }
                } else {
                    lk = __ldg(pk + 1);
                }
                if ((r2 < cf.x) & (r2 > cf.y)) break;
                const char* ls = lrow + ls_off;
#pragma unroll 1
                for (;;) {
                    const double2* row = reinterpret_cast<const double2*>(ls + (unsigned)(k + 1) * 32u);
                    const double2 cur = __ldg(row);
                    const double2 nxt = __ldg(row + 2);
                    if ((r2 < cur.x) & (r2 > nxt.y)) break;
                    if (r2 > cur.y) { --k; continue; }
                    if (r2 < nxt.x) { ++k; continue; }
                    slow = true;
                    break;
                }
                if (slow) break;
            }
            if (slow) { slow_mask |= 1ull << iz; continue; }
            const bool pair = (unsigned)k < pair_top;
            const bool skip = (unsigned)(k - k_first) > k_span;
            const int lower = max(k, 0);
            if (skip) {                                        // blind-zone rules: nothing sampled, nothing observable
                continue;                                      // (nothing valid, nothing observable: no contribution)
            }
            if (dead_unobserved && !pair) continue;            // one-sweep fill-in with weight zero: nothing to add
            int ngm1 = u_ngm1;
            float rj, idr;
            if (SPEC) {
                if (!((r2 >= tg.x) & (r2 < tg.y))) { slow_mask |= 1ull << iz; continue; }
                rj = fmaf((float)(gi - 1), u_step, u_r_first); idr = u_inv_step;
            } else {
                // ---- bracket, then gather ----
                unsigned rng_off = u_rng_off;
                float inv_step = u_inv_step, c0 = u_c0;
                bool usable = true;
                if (!UNI) {
                    const int4 tab = __ldg(reinterpret_cast<const int4*>(&fs_t[lower].rng_off));
                    rng_off = (unsigned)tab.x;
                    ngm1 = tab.y;
                    c0 = __int_as_float(tab.z); inv_step = __int_as_float(tab.w);
                    const int need = pair ? (kFsPolyOk | kFsShared | kFsUsable) : kFsUsable;
                    usable = (__ldg(&fs_t[lower].flags) & need) == need;
                }
                gi = usable ? min(max(__float2int_rd(fmaf(rf, inv_step, c0)), 1), ngm1) : 1;
                const unsigned gidx = usable ? rng_off + (unsigned)(gi - 1) : u_rng_off;
                tg = ldg_vol(g_t + gidx);                        // r in [R[gi-1], R[gi])  <=>  tg.x <= r2 < tg.y
                if (!(usable & (r2 >= tg.x) & (r2 < tg.y))) { slow_mask |= 1ull << iz; continue; }
                if (UNI) {
                    rj = fmaf((float)(gi - 1), u_step, u_r_first); idr = u_inv_step;
                } else {
                    const int4 hv = __ldg(reinterpret_cast<const int4*>(R.r2.h_t + gidx));   // {R^2, R, 1/dR}
                    rj = __int_as_float(hv.z); idr = __int_as_float(hv.w);
                }
                ea = lds_az(az_addr + (unsigned)lower * az_pitch);
                qa4 = ldg_vol(val0 + (ea.row + (unsigned)gi));
                if (pair) {
                    eb = lds_az(az_addr + (unsigned)(lower + 1) * az_pitch);
                    qb4 = ldg_vol(val0 + (eb.row + (unsigned)gi));
                    q14 = w_pq[2 * lower];
                    q58 = w_pq[2 * lower + 1];
                }
            }
            const float ea0 = fmaf(ea.w, qa4.y - qa4.x, qa4.x), ea1 = fmaf(ea.w, qa4.w - qa4.z, qa4.z);
            const float w_r = (float)sub(r2, tg.x) * idr * rcp_approx(rf + rj);
            float v = fmaf(w_r, ea1 - ea0, ea0);
            bool obs = false;
            if (pair) {
                const float u = (float)sub(r2, __hiloint2double(lk.y, lk.x)) * rcp_approx(rf + __int_as_float(lk.z)) *
                                fmaf(__int_as_float(lk.w), rinv, inv2a);
                float w_el = fmaf(q58.x, u, q14.w);
                w_el = fmaf(w_el, u, q14.z);
                w_el = fmaf(w_el, u, q14.y);
                w_el = fmaf(w_el, u, q14.x);
                w_el *= u;
                const float eb0 = fmaf(eb.w, qb4.y - qb4.x, qb4.x), eb1 = fmaf(eb.w, qb4.w - qb4.z, qb4.z);
                const float v1 = fmaf(w_r, eb1 - eb0, eb0);
                v = lerp_fill(w_el, v, v1, fill32);
                // between two sweeps and at least one gate away from both ends of their (shared) range table:
                // observable whatever the few-ulp difference between the NumPy twin's r and this one
                obs = (gi > 1) & (gi < ngm1);
                if (!obs) obs = net2_obs_edge(r2, np_t + lower, x, y, &R, lv + iz, false, &ties);
            }
            merge(iz, v, obs, rf, q58.y, q58.z);
        }

        // ---- the marked voxels: the reference's own chain (pcg_fast.cuh fast_voxel) ----
        if (slow_mask) {
            SlowArgs sa;
            sa.s_sw = R.vol.sweeps; sa.s_pr = R.pairs; sa.my_az = my_az; sa.my_iaz = nullptr;
            sa.rg_t = reinterpret_cast<const double2*>(R.vol.rng); sa.val = R.vol.val; sa.az_stride = T; sa.ne = ne;
            sa.a2 = R.a2; sa.twoa = R.twoa; sa.guard = R.guard; sa.inv_twoa = R.inv_twoa; sa.fill32 = fill32;
            sa.blind = (p.nearest_gate != 0 ? 1 : 0) | (lowest_sweep ? 2 : 0) | (highest_sweep ? 4 : 0);
            int k6 = 0;
#pragma unroll 1
            while (slow_mask) {
                const int iz = __ffsll((long long)slow_mask) - 1;
                slow_mask &= slow_mask - 1;
                VoxelInfo vi;
                float tmp[1];
                r3_slow_voxel<1, false>(&sa, &lv[iz].cy, cphi, &k6, tmp, &vi);
                const bool is_pair = vi.bracket == kPair;
                const int kk = vi.lower < ne - 1 ? (vi.lower < 0 ? 0 : vi.lower) : ne - 2;
                const NetPair* np = np_t + (is_pair ? kk : 0);
                bool obs = false;
                if (is_pair || vi.tie || !(vi.r2 > 0.0))
                    obs = net2_obs_edge(vi.r2, np, x, y, &R, lv + iz, vi.tie || !is_pair, &ties) && (is_pair || vi.tie);
                const NetPairQ* pq = pq_t + (is_pair ? kk : 0);
                merge(iz, tmp[0], obs, (float)vi.r, pq->beam_k, pq->vert_w);
            }
        }

        if (p.cr != nullptr) {
            // get_CR_xy on the local grid: every sweep at its fixed elevation, blind "mask"
            // (RadarGridC.pyx:596-624), then the NaN-aware max over radars (RadarInterp.py:2047-2063)
            const double sphi = sin_small(phi);
            if (UNI) {
                // Uniform-table volumes, three sweeps at a time: the geometry chain of the reference in FP64 up to
                // r2 (bit-identical), the gate bracket in r2 space like the level loop (r >= R[j] <=> r2 >= tg[j]:
                // no FP64 square root, one band load), and the three bands and three quads requested together.
                // A guess the band does not confirm takes the reference chain (net2_cr_exact).
                const double2 ends = __ldg(reinterpret_cast<const double2*>(&fs_t[0].tg_first));   // {tg_first, tg_gt_last}
#pragma unroll 1
                for (int ie0 = 0; ie0 < ne; ie0 += NET2_GRP) {
                    double r2v[NET2_GRP];
                    float rfv[NET2_GRP];
                    int giv[NET2_GRP];
                    bool okv[NET2_GRP];
                    double2 tgv[NET2_GRP];
                    float4 qv[NET2_GRP];
                    AzEntry ev[NET2_GRP];
#pragma unroll
                    for (int j = 0; j < NET2_GRP; ++j) {
                        const int ie = min(ie0 + j, ne - 1);
                        const double denom = sub(cphi, mul(sphi, w_tan[ie]));
                        const double g = div(R.a, denom);
                        double r2 = sub(add(R.a2, mul(g, g)), mul(mul(R.twoa, g), cphi));
                        if (r2 < 0.0) r2 = 0.0;
                        // r in [R[0], R[ng-1]] (RadarGridC.pyx:275-280, blind "mask"); a NaN fails both tests
                        okv[j] = (ie0 + j < ne) & (denom != 0.0) & (r2 >= ends.x) & (r2 < ends.y);
                        r2v[j] = r2;
                        const float f = okv[j] ? (float)r2 : 1.f;
                        rfv[j] = f * rsqrt_approx(f);
                        giv[j] = min(max(__float2int_rd(fmaf(rfv[j], u_inv_step, u_c0)), 1), u_ngm1);
                        tgv[j] = ldg_vol(g_t + (u_rng_off + (unsigned)(giv[j] - 1)));
                        ev[j] = lds_az(az_addr + (unsigned)ie * az_pitch);
                        qv[j] = ldg_vol(val0 + (ev[j].row + (unsigned)giv[j]));
                    }
#pragma unroll
                    for (int j = 0; j < NET2_GRP; ++j) {
                        if (!okv[j]) continue;
                        float v;
                        if ((r2v[j] >= tgv[j].x) & (r2v[j] < tgv[j].y)) {
                            const float rj = fmaf((float)(giv[j] - 1), u_step, u_r_first);
                            const float w_r = (float)sub(r2v[j], tgv[j].x) * u_inv_step * rcp_approx(rfv[j] + rj);
                            const float er0 = fmaf(ev[j].w, qv[j].y - qv[j].x, qv[j].x);
                            const float er1 = fmaf(ev[j].w, qv[j].w - qv[j].z, qv[j].z);
                            v = lerp_fill(w_r, er0, er1, fill32);
                        } else {
                            v = net2_cr_exact(R.vol.sweeps + (ie0 + j), reinterpret_cast<const double2*>(R.vol.rng), val0,
                                              my_az + (size_t)(ie0 + j) * T, r2v[j], fill32);
                        }
                        if (v == fill32 || v != v) continue;
                        if (!(cr_best >= v)) cr_best = v;     // NaN (nothing yet) or smaller
                    }
                }
            } else {
                const double2* rg_t = reinterpret_cast<const double2*>(R.vol.rng);
                for (int ie = 0; ie < ne; ++ie) {
                    const SweepDesc& d = R.vol.sweeps[ie];
                    if (d.naz < 2 || d.ng < 2) continue;
                    const double denom = sub(cphi, mul(sphi, d.tan_e));
                    if (denom == 0.0) continue;
                    const double g = div(R.a, denom);
                    double r2 = sub(add(R.a2, mul(g, g)), mul(mul(R.twoa, g), cphi));
                    if (r2 < 0.0) r2 = 0.0;
                    const float v = net2_cr_exact(&d, rg_t, val0, my_az + (size_t)ie * T, r2, fill32);
                    if (v == fill32 || v != v) continue;
                    if (!(cr_best >= v)) cr_best = v;         // NaN (nothing yet) or smaller
                }
            }
        }
    }

    // ---- 4. the tile's outputs, once ----
    if (active) {
        size_t o = (size_t)col;
        for (int iz = 0; iz < nz; ++iz, o += (size_t)p.ncol) {
            const size_t c = (size_t)iz * T + tid;
            const float2 acc = a_acc[c];
            const float awv = acc.x, aw = acc.y;
            const unsigned cc = a_cc[c];
            double out = p.out_missing;
            if (METHOD == kNetMax || METHOD == kNetNearest) {
                if (aw != 0.f) out = (double)awv;
            } else {
                if (aw > 0.f) out = (double)awv / (double)aw;            // has_data = total_weight > 0
            }
            if (p.composite) p.composite[o] = out;
            if (p.npeer) {
                const size_t at = (size_t)iz * (size_t)p.peer_plane + (size_t)p.peer_off + (size_t)col;
                for (int q = 0; q < p.npeer; ++q) p.peer[q][at] = out;
            }
            if (p.valid_count) p.valid_count[o] = (int16_t)(cc & 0xffu);
            if (p.coverage) p.coverage[o] = (int16_t)(cc >> 8);
            if (p.blind) p.blind[o] = (int8_t)((cc >> 8) == 0u);
        }
        if (p.cr) p.cr[col] = (double)cr_best;               // NaN where no radar has an echo
    }
    if (ties && p.tie_count) atomicAdd(p.tie_count, ties);
}

}  // namespace pcg
