// pcg_vvp.cuh — local moving-window VVP retrieval on one sweep:
//   _run_local_vvp (pycwr/retrieve/WindField.py:1843-1941) with its helpers
//   _azimuth_coverage_deg (:58-68), _azimuth_sector_count (:71-81), _effective_min_sector_count (:84-90),
//   _design_matrix (:93-106), _vvp_window_weights (:1832-1840), _solve_weighted_least_squares (:115-200)
//   and _wind_direction_from_uv (:109-112).
//
// The reference runs a Python loop over every (ray, gate) and, per window, NumPy calls for the window
// statistics, up to six 3x3 weighted least-squares solves (initial + 5 Huber re-weightings) and two
// medians per re-weighting.  Here one WARP owns one window (az_num x bin_num samples, 819 by default):
//   * the window's velocities are staged once in shared memory; the rows' design terms, azimuths and
//     Gaussian weights are staged once per block (the block's warps are consecutive gates of one ray);
//   * normal equations are accumulated per lane and reduced by butterfly shuffles (every lane ends with
//     the same sums, so the 3x3 work — Jacobi eigenvalues for the 2-norm condition number, LU with
//     partial pivoting like LAPACK's gesv — runs redundantly without further communication);
//   * medians are exact: bitwise selection on order-preserving 64-bit keys (high words first, the low
//     words only among keys that share the selected high word), counts via the warp reduction instruction.
// FP64 throughout (velocities: 1e-5 relative bar).  Window statistics (counts, valid fraction, azimuth
// coverage, sector count) are bit-exact; the fits agree to rounding (different summation order).
#pragma once
#include "pcg_math.cuh"
#include "pcg_network.cuh"

namespace pcg {

constexpr int kVvpMaxWarps = 4;       // windows (consecutive gates of one ray) per block

struct VvpArgs {
    const double* vel;        // [naz][nrange], NaN = missing
    const double* d0;         // [naz] sin(az) cos(el)
    const double* d1;         // [naz] cos(az) cos(el)
    const double* azm;        // [naz] azimuth mod 360 (NaN when the azimuth is not finite)
    const double* w_az;       // [az_num] window weight factors
    const double* w_bin;      // [bin_num]
    int naz, nrange, az_num, bin_num;
    double min_valid_fraction;
    int min_valid_count;
    double min_coverage;
    int min_sector_count;     // as requested (capped per window by the sectors its rays can occupy)
    double max_cond;
    double* u; double* v; double* speed; double* direction; double* offset; double* rmse; double* cond;
    double* valid_fraction; double* coverage;
    int* sector_count; int* valid_count;
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

__device__ __forceinline__ double warp_min(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// order_key (pcg_network.cuh): doubles -> unsigned keys with the same order; key_value inverts it
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// c += (w >= cand): the carry-out of w - cand feeds an add-with-carry (two integer instructions, no select)
__device__ __forceinline__ unsigned add_not_below(unsigned c, unsigned w, unsigned cand) {
    unsigned t;
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %2, %3;\n\taddc.u32 %0, %1, 0;\n\t}" : "=r"(t) : "r"(c), "r"(w), "r"(cand));
    return t;
}

// #{w[s] < cand : s = lane, lane + 32, ...} summed over the warp = n - #{w[s] >= cand}: four independent
// partial counts, immediate-offset loads
__device__ __forceinline__ unsigned count_below(const unsigned* w, int n, unsigned cand, int lane) {
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int s = lane;
    for (; s + 96 < n; s += 128) {
        c0 = add_not_below(c0, w[s], cand);
        c1 = add_not_below(c1, w[s + 32], cand);
        c2 = add_not_below(c2, w[s + 64], cand);
        c3 = add_not_below(c3, w[s + 96], cand);
    }
    for (; s < n; s += 32) c0 = add_not_below(c0, w[s], cand);
    return (unsigned)n - __reduce_add_sync(0xffffffffu, (c0 + c1) + (c2 + c3));
}

// Keys live in shared memory as separate high / low 32-bit words (conflict-free 32-bit loads); slots that
// are not part of the fit hold hi = 0xffffffff (the key of a negative NaN: never produced by a residual).
// k-th smallest (0-based): bitwise selection on the high words first — the largest P with #{hi < P} <= k —
// then, only if several keys share that high word, the same on their low words.
//
// The bit-by-bit narrowing on the high words stops as soon as at most 32 keys remain in the surviving
// interval (a handful of iterations for residual-like data: sign and exponent settle quickly and only a
// few per cent of a window share the median's binade); those keys are compacted to one per lane and
// ranked with shuffles.  Only degenerate windows (more than 32 keys sharing a complete high word) run
// the selection down to the last bit and repeat it on the low words.
struct Select2 { unsigned long long kth, next; };     // order statistics k and k + 1 (next: when asked for)

__device__ Select2 warp_select(const unsigned* hi, const unsigned* lo, int n, int m, int k, bool want_next,
                               unsigned long long* cand_buf, int lane) {
    Select2 out;
    unsigned all_or = 0u, all_and = ~0u;
    for (int s = lane; s < n; s += 32) {
        const unsigned h = hi[s];
        if (h != ~0u) { all_or |= h; all_and &= h; }
    }
    all_or = __reduce_or_sync(0xffffffffu, all_or);
    all_and = __reduce_and_sync(0xffffffffu, all_and);
    const unsigned diff = all_or ^ all_and;
    int b = diff ? 31 - __clz((int)diff) : -1;          // bits b..0 of the high word are still undecided
    unsigned H = (b >= 0) ? (all_and & ~((2u << b) - 1u)) : all_and;
    int below = 0, upto = m;                             // #keys < [H, last], #keys <= last
    while (upto - below > 32 && b >= 0) {
        const unsigned cand = H | (1u << b);
        const int cnt = (int)count_below(hi, n, cand, lane);
        if (cnt <= k) { H = cand; below = cnt; } else { upto = cnt; }
        --b;
    }
    if (upto - below <= 32) {
        const unsigned last = H | ((b >= 0) ? ((2u << b) - 1u) : 0u);
        // compact the survivors, one per lane
        int filled = 0;
        for (int s0 = 0; s0 < n; s0 += 32) {
            const int s = s0 + lane;
            const unsigned h = s < n ? hi[s] : ~0u;
            const bool in = (h >= H) & (h <= last) & (h != ~0u);
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) cand_buf[filled + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)h << 32) | lo[s];
            filled += __popc(bal);
        }
        __syncwarp();
        const unsigned long long key = lane < filled ? cand_buf[lane] : ~0ull;
        __syncwarp();
        int rank = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const unsigned long long o = __shfl_sync(0xffffffffu, key, j);
            rank += (o < key || (o == key && j < lane)) ? 1 : 0;
        }
        const int r = k - below;
        unsigned who = __ballot_sync(0xffffffffu, rank == r);
        out.kth = __shfl_sync(0xffffffffu, key, __ffs((int)who) - 1);
        out.next = out.kth;
        if (want_next) {
            if (r + 1 < filled) {
                who = __ballot_sync(0xffffffffu, rank == r + 1);
                out.next = __shfl_sync(0xffffffffu, key, __ffs((int)who) - 1);
            } else {                                     // the smallest key above the interval
                unsigned long long up = ~0ull;
                for (int s = lane; s < n; s += 32) {
                    const unsigned h = hi[s];
                    if (h > last && h != ~0u) {
                        const unsigned long long q = ((unsigned long long)h << 32) | lo[s];
                        up = q < up ? q : up;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long t = __shfl_xor_sync(0xffffffffu, up, o);
                    up = t < up ? t : up;
                }
                out.next = up;
            }
        }
        return out;
    }
    // ---- more than 32 keys share the high word H: select on their low words ----------------------
    unsigned l_or = 0u, l_and = ~0u;
    for (int s = lane; s < n; s += 32)
        if (hi[s] == H) { const unsigned l = lo[s]; l_or |= l; l_and &= l; }
    l_or = __reduce_or_sync(0xffffffffu, l_or);
    l_and = __reduce_and_sync(0xffffffffu, l_and);
    const unsigned ldiff = l_or ^ l_and;
    unsigned L = l_and;
    if (ldiff) {
        const int k2 = k - below;
        const int top = 31 - __clz((int)ldiff);
        L = l_and & ~((2u << top) - 1u);
        for (int bb = top; bb >= 0; --bb) {
            const unsigned cand = L | (1u << bb);
            unsigned cnt = 0;
            for (int s = lane; s < n; s += 32) cnt += (hi[s] == H && lo[s] < cand) ? 1u : 0u;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if ((int)cnt <= k2) L = cand;
        }
    }
    out.kth = ((unsigned long long)H << 32) | L;
    out.next = out.kth;
    if (want_next) {
        // `kth` again when it is repeated beyond position k, else the smallest key above it
        unsigned le = 0;
        unsigned long long up = ~0ull;
        for (int s = lane; s < n; s += 32) {
            const unsigned long long q = ((unsigned long long)hi[s] << 32) | lo[s];
            le += q <= out.kth ? 1u : 0u;
            if (q > out.kth && q < up) up = q;
        }
        le = __reduce_add_sync(0xffffffffu, le);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, up, o);
            up = t < up ? t : up;
        }
        out.next = ((int)le >= k + 2) ? out.kth : up;
    }
    return out;
}

// np.median of the m keys taking part: the middle one, or the mean of the middle two
__device__ double warp_median(const unsigned* hi, const unsigned* lo, int n, int m, unsigned long long* cand_buf,
                              int lane) {
    const Select2 s2 = warp_select(hi, lo, n, m, (m - 1) >> 1, !(m & 1), cand_buf, lane);
    if (m & 1) return key_value(s2.kth);
    return (key_value(s2.kth) + key_value(s2.next)) / 2.0;
}

// eigenvalues of a symmetric 3x3 by cyclic Jacobi rotations; returns max|lambda| / min|lambda|
__device__ double sym3_condition(double a00, double a01, double a02, double a11, double a12, double a22) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = A[p][q];
            if (apq == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            const int r = 3 - p - q;
            const double arp = A[r][p], arq = A[r][q];
            A[p][p] -= t * apq;
            A[q][q] += t * apq;
            A[p][q] = A[q][p] = 0.0;
            A[r][p] = A[p][r] = c * arp - s * arq;
            A[r][q] = A[q][r] = s * arp + c * arq;
        }
    }
    const double e0 = fabs(A[0][0]), e1 = fabs(A[1][1]), e2 = fabs(A[2][2]);
    const double hi = fmax(e0, fmax(e1, e2)), lo = fmin(e0, fmin(e1, e2));
    if (!(hi == hi) || !(lo == lo)) return __longlong_as_double(0x7ff8000000000000LL);
    return lo == 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : hi / lo;
}

// 3x3 solve, LU with partial pivoting (the algorithm of LAPACK gesv behind np.linalg.solve)
__device__ bool solve3(double N[3][3], double b[3], double x[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        double best = fabs(N[c][c]);
#pragma unroll
        for (int r = c + 1; r < 3; ++r)
            if (fabs(N[r][c]) > best) { best = fabs(N[r][c]); piv = r; }
        if (best == 0.0 || !(best == best)) return false;
        if (piv != c) {
#pragma unroll
            for (int j = 0; j < 3; ++j) { const double t = N[c][j]; N[c][j] = N[piv][j]; N[piv][j] = t; }
            const double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            const double f = N[r][c] / N[c][c];
#pragma unroll
            for (int j = c; j < 3; ++j) N[r][j] -= f * N[c][j];
            b[r] -= f * b[c];
        }
    }
    x[2] = b[2] / N[2][2];
    x[1] = (b[1] - N[1][2] * x[2]) / N[1][1];
    x[0] = (b[0] - N[0][1] * x[1] - N[0][2] * x[2]) / N[0][0];
    return true;
}

struct VvpWindow {
    const double* resp;   // [n] velocities of the window (NaN = not part of the fit)
    const double* d0;     // [az_num] per window row
    const double* d1;
    const double* wa;
    const double* wb;
    int n, bin_num;
    float inv_bin;
};

// one weighted solve (_solve_once): Huber factors from the residuals about `prev` when scale > 0
__device__ bool vvp_solve(const VvpWindow& w, int lane, const double prev[3], double scale, double max_cond,
                          double coef[3], double& cond) {
    double n00 = 0, n01 = 0, n02 = 0, n11 = 0, n12 = 0, n22 = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int s = lane; s < w.n; s += 32) {
        const double y = w.resp[s];
        if (!(y == y)) continue;
        const int row = __float2int_rd(((float)s + 0.5f) * w.inv_bin);
        const int col = s - row * w.bin_num;
        const double a0 = w.d0[row], a1 = w.d1[row];
        double wt = w.wa[row] * w.wb[col];
        if (scale > 0.0) {
            const double res = y - (a0 * prev[0] + a1 * prev[1] + prev[2]);
            // np.where(normalized <= 1, 1, 1 / normalized), normalized = |res| / (1.5 scale): one division
            const double lim = 1.5 * scale, ar = fabs(res);
            if (!(ar <= lim)) wt *= lim / ar;
        }
        const double sq = sqrt(wt);
        const double x0 = a0 * sq, x1 = a1 * sq, yy = y * sq;
        n00 += x0 * x0; n01 += x0 * x1; n02 += x0 * sq;
        n11 += x1 * x1; n12 += x1 * sq; n22 += sq * sq;
        b0 += x0 * yy; b1 += x1 * yy; b2 += sq * yy;
    }
    n00 = warp_sum(n00); n01 = warp_sum(n01); n02 = warp_sum(n02);
    n11 = warp_sum(n11); n12 = warp_sum(n12); n22 = warp_sum(n22);
    b0 = warp_sum(b0); b1 = warp_sum(b1); b2 = warp_sum(b2);
    cond = sym3_condition(n00, n01, n02, n11, n12, n22);
    if (!isfinite(cond) || cond > max_cond) return false;
    double N[3][3] = {{n00, n01, n02}, {n01, n11, n12}, {n02, n12, n22}};
    double b[3] = {b0, b1, b2};
    return solve3(N, b, coef);
}

__global__ void __launch_bounds__(kVvpMaxWarps * 32) vvp_kernel(const __grid_constant__ VvpArgs a) {
    extern __shared__ double sm[];
    const int A = a.az_num, B = a.bin_num, n = A * B;
    const int half_az = A / 2, half_bin = B / 2;
    double* s_d0 = sm;
    double* s_d1 = s_d0 + A;
    double* s_azm = s_d1 + A;
    double* s_wa = s_azm + A;
    double* s_wb = s_wa + A;
    double* per_warp = s_wb + B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* s_resp = per_warp + (size_t)warp * (2 * (size_t)n + 32 + (size_t)((A + 1) / 2));
    unsigned* s_hi = (unsigned*)(s_resp + n);
    unsigned* s_lo = s_hi + n;
    unsigned long long* s_cand = (unsigned long long*)(s_lo + n);        // 32 compacted keys
    int* s_rowv = (int*)(s_cand + 32);

    const int az_index = blockIdx.y;
    const int gate = blockIdx.x * (blockDim.x >> 5) + warp;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    // rows of the window: np.pad(..., mode="wrap") about ray az_index (WindField.py:1876-1878)
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        int row = (az_index - half_az + i) % a.naz;
        if (row < 0) row += a.naz;
        s_d0[i] = a.d0[row];
        s_d1[i] = a.d1[row];
        s_azm[i] = a.azm[row];
        s_wa[i] = a.w_az[i];
    }
    for (int i = threadIdx.x; i < B; i += blockDim.x) s_wb[i] = a.w_bin[i];
    __syncthreads();
    if (gate >= a.nrange) return;

    const size_t o = (size_t)az_index * a.nrange + gate;
    double r_u = nan, r_v = nan, r_speed = nan, r_dir = nan, r_off = nan, r_rmse = nan, r_cond = nan;
    double r_vf = nan, r_cov = nan;
    int r_sec = 0, r_vc = 0;

    if (gate >= half_bin && gate < a.nrange - half_bin) {
        // ---- stage the window, count its finite velocities (:1903-1906) ----------------------------
        for (int i = lane; i < A; i += 32) s_rowv[i] = 0;
        __syncwarp();
        const float inv_bin = 1.0f / (float)B;
        unsigned vc = 0, nfit = 0;
        for (int s = lane; s < n; s += 32) {
            const int row = __float2int_rd(((float)s + 0.5f) * inv_bin);
            const int col = s - row * B;
            int src = (az_index - half_az + row) % a.naz;
            if (src < 0) src += a.naz;
            const double y = a.vel[(size_t)src * a.nrange + (gate - half_bin + col)];
            const bool fin = isfinite(y);
            const bool fit = fin && isfinite(s_d0[row]) && isfinite(s_d1[row]);
            s_resp[s] = fit ? y : nan;
            if (fin) { s_rowv[row] = 1; ++vc; }
            if (fit) ++nfit;
        }
        vc = __reduce_add_sync(0xffffffffu, vc);
        nfit = __reduce_add_sync(0xffffffffu, nfit);
        __syncwarp();
        r_vc = (int)vc;
        r_vf = (double)r_vc / (double)n;

        // ---- azimuth coverage (:58-68) and occupied 45-degree sectors (:71-81) ----------------------
        // coverage = 360 - largest gap between consecutive distinct azimuths of the rays holding a
        // finite sample; the successor of each ray is found by direct search (no ordering assumed)
        double gap = -1.0, lo_az = __longlong_as_double(0x7ff0000000000000LL);
        unsigned sectors = 0, capacity = 0;
        for (int i = lane; i < A; i += 32) {
            const double z = s_azm[i];
            if (!(z == z)) continue;
            capacity |= 1u << (int)floor(z / 45.0);
            if (!s_rowv[i]) continue;
            sectors |= 1u << (int)floor(z / 45.0);
            lo_az = fmin(lo_az, z);
        }
        lo_az = warp_min(lo_az);
        sectors = __reduce_or_sync(0xffffffffu, sectors);
        capacity = __reduce_or_sync(0xffffffffu, capacity);
        int distinct = 0;
        for (int i = lane; i < A; i += 32) {
            const double z = s_azm[i];
            if (!(z == z) || !s_rowv[i]) continue;
            double succ = __longlong_as_double(0x7ff0000000000000LL);
            for (int j = 0; j < A; ++j) {
                const double q = s_azm[j];
                if (s_rowv[j] && q > z && q < succ) succ = q;
            }
            if (succ < __longlong_as_double(0x7ff0000000000000LL)) { gap = fmax(gap, sub(succ, z)); distinct = 1; }
            else gap = fmax(gap, sub(add(lo_az, 360.0), z));
        }
        gap = warp_max(gap);
        distinct = __reduce_or_sync(0xffffffffu, (unsigned)distinct);
        r_cov = (distinct && gap >= 0.0) ? fmax(0.0, sub(360.0, gap)) : 0.0;
        r_sec = __popc(sectors);
        int required = 0;
        if (a.min_sector_count > 0) required = min(a.min_sector_count, max(__popc(capacity), 1));

        if (!(r_vc < a.min_valid_count || r_vf < a.min_valid_fraction || r_cov < a.min_coverage || r_sec < required) &&
            nfit > 0) {
            // ---- weighted least squares with Huber re-weighting (:115-200) ---------------------------
            VvpWindow w;
            w.resp = s_resp; w.d0 = s_d0; w.d1 = s_d1; w.wa = s_wa; w.wb = s_wb;
            w.n = n; w.bin_num = B; w.inv_bin = inv_bin;
            double coef[3] = {0.0, 0.0, 0.0}, cond = nan;
            const double zero3[3] = {0.0, 0.0, 0.0};
            if (vvp_solve(w, lane, zero3, 0.0, a.max_cond, coef, cond)) {
#pragma unroll 1
                for (int it = 0; it < 5; ++it) {
                    // scale = 1.4826 * median(|res - median(res)|)
                    for (int s = lane; s < n; s += 32) {
                        const double y = s_resp[s];
                        const int row = __float2int_rd(((float)s + 0.5f) * inv_bin);
                        const unsigned long long q =
                            (y == y) ? order_key(y - (s_d0[row] * coef[0] + s_d1[row] * coef[1] + coef[2])) : ~0ull;
                        s_hi[s] = (unsigned)(q >> 32);
                        s_lo[s] = (unsigned)q;
                    }
                    __syncwarp();
                    const double med = warp_median(s_hi, s_lo, n, (int)nfit, s_cand, lane);
                    __syncwarp();
                    for (int s = lane; s < n; s += 32) {
                        const unsigned h = s_hi[s];
                        if (h == ~0u) continue;
                        const unsigned long long q = order_key(fabs(key_value(((unsigned long long)h << 32) | s_lo[s]) - med));
                        s_hi[s] = (unsigned)(q >> 32);
                        s_lo[s] = (unsigned)q;
                    }
                    __syncwarp();
                    const double scale = 1.4826 * warp_median(s_hi, s_lo, n, (int)nfit, s_cand, lane);
                    __syncwarp();
                    if (!isfinite(scale) || scale < 1.0e-6) break;
                    double nc[3], ncond;
                    if (!vvp_solve(w, lane, coef, scale, a.max_cond, nc, ncond)) break;
                    bool close = true;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        close = close && (fabs(nc[k] - coef[k]) <= 1.0e-5 + 1.0e-5 * fabs(coef[k]));
                    coef[0] = nc[0]; coef[1] = nc[1]; coef[2] = nc[2];
                    cond = ncond;
                    if (close) break;
                }
                double ss = 0.0;
                for (int s = lane; s < n; s += 32) {
                    const double y = s_resp[s];
                    if (!(y == y)) continue;
                    const int row = __float2int_rd(((float)s + 0.5f) * inv_bin);
                    const double res = y - (s_d0[row] * coef[0] + s_d1[row] * coef[1] + coef[2]);
                    ss += res * res;
                }
                ss = warp_sum(ss);
                r_u = coef[0];
                r_v = coef[1];
                r_off = coef[2];
                r_speed = hypot(r_u, r_v);
                double dir = fmod(sub(270.0, mul(atan2(r_v, r_u), kRad2Deg)), 360.0);
                if (dir != 0.0) { if (dir < 0.0) dir = add(dir, 360.0); } else dir = 0.0;   // np.mod: sign of the divisor
                r_dir = dir;
                r_rmse = sqrt(ss / (double)nfit);
            }
            r_cond = cond;
        }
    }
    if (lane == 0) {
        a.u[o] = r_u; a.v[o] = r_v; a.speed[o] = r_speed; a.direction[o] = r_dir; a.offset[o] = r_off;
        a.rmse[o] = r_rmse; a.cond[o] = r_cond; a.valid_fraction[o] = r_vf; a.coverage[o] = r_cov;
        a.sector_count[o] = r_sec; a.valid_count[o] = r_vc;
    }
}

}  // namespace pcg
