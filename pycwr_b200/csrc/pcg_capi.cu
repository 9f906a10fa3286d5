// pcg_capi.cu — C ABI (include/pycwr_b200.h) over the sm_100a gridding kernels.
//
// Host-side responsibilities: pack the reference's list-of-arrays volume into device arenas
// (the role of PRD.get_vol_data's output, pycwr/core/NRadar.py:1931-1961), evaluate the few
// per-level / per-sweep constants that need glibc or plain IEEE arithmetic, stage host buffers
// and launch.  There is no CPU compute path: without a CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <memory>
#include <unordered_map>
#include <chrono>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/pycwr_b200.h"
#include "pcg_fast.cuh"
#include "pcg_kernels.cuh"
#include "pcg_r2.cuh"
#include "pcg_r3.cuh"
#include "pcg_network.cuh"
#include "pcg_net2.cuh"
#include "pcg_products.cuh"
#include "pcg_wind.cuh"
#include "pcg_section.cuh"
#include "pcg_vvp.cuh"

using namespace pcg;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(PCG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

// grow-only device scratch buffers for the host-buffer entry points
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
    int device = -1;
    int ensure(size_t bytes) {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (ptr && (cap < bytes || device != dev)) {
            cudaFree(ptr);
            ptr = nullptr;
            cap = 0;
        }
        if (!ptr) {
            size_t want = bytes < 256 ? 256 : bytes;
            cudaError_t e = cudaMalloc(&ptr, want);
            if (e != cudaSuccess) {
                ptr = nullptr;
                return fail(PCG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            }
            cap = want;
            device = dev;
        }
        return PCG_OK;
    }
};


// Device-memory pool for the per-call volume arenas and tables: a gridding call that packs its own volume
// (the reference's calling convention) would otherwise pay a dozen cudaMalloc / cudaFree pairs, and every
// cudaFree synchronises the device.  Exact-size reuse, bounded; blocks return after a device synchronise.
struct DevPool {
    std::mutex m;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_list;
    std::unordered_map<void*, std::pair<int, size_t>> live;
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)8 << 30;

    cudaError_t alloc(void** out, size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (bytes == 0) bytes = 256;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(m);
        auto it = free_list.find({dev, bytes});
        if (it != free_list.end() && !it->second.empty()) {
            *out = it->second.back();
            it->second.pop_back();
            cached -= bytes;
        } else {
            e = cudaMalloc(out, bytes);
            if (e == cudaErrorMemoryAllocation && cached) {       // give the cache back and retry once
                for (auto& kv : free_list) for (void* q : kv.second) cudaFree(q);
                free_list.clear();
                cached = 0;
                cudaGetLastError();
                e = cudaMalloc(out, bytes);
            }
            if (e != cudaSuccess) { *out = nullptr; return e; }
        }
        live[*out] = {dev, bytes};
        return cudaSuccess;
    }

    void release(void* ptr) {
        if (!ptr) return;
        std::lock_guard<std::mutex> lk(m);
        auto it = live.find(ptr);
        if (it == live.end()) { cudaFree(ptr); return; }
        const std::pair<int, size_t> key = it->second;
        live.erase(it);
        if (cached + key.second <= kMaxCached) {
            free_list[key].push_back(ptr);
            cached += key.second;
        } else {
            cudaFree(ptr);
        }
    }
};
DevPool g_pool;

enum { S_GX = 0, S_GY, S_OUT, S_IDX, S_TMP0, S_TMP1, S_TMP2, S_NET, S_AUX0, S_AUX1, S_AUX2, S_AUX3, S_AUX4, S_ACC0, S_ACC4, S_ACC5, S_ACC6, S_COUNT };

constexpr int kMaxDevices = 16;
constexpr int kNetStreams = 6;
struct NetStreams {
    cudaStream_t side[kNetStreams] = {};
    cudaEvent_t ev_side[kNetStreams] = {};
    cudaEvent_t ev_class = nullptr;
    bool ready = false;
};
// peers of the next enqueue_network call (set and cleared by pcg_network_compose_gather_dev under the device lock)
struct NetPeers { int n = 0; long long plane = 0, off = 0; double* ptr[kNetMaxPeers] = {}; };

// Everything a call needs besides its arguments lives in the context of the CURRENT device (cudaSetDevice is
// per host thread): scratch arena, library stream, table cache, side streams — one host thread per GPU can
// therefore drive its device without meeting the other threads (the reference fans out over processes,
// pycwr/interp/RadarInterp.py:1816-1861; here the fan-out is over the GPUs of one box).
struct DeviceCtx {
    std::mutex mutex;                   // guards this device's scratch arena, library stream and table cache
    Scratch scratch[S_COUNT];
    cudaStream_t stream = nullptr;
    int* nonfinite_flag = nullptr;      // raised by the packed kernels on NaN / inf targets (host entry points)
    std::vector<char> net_key;          // what the cached network tables were built from
    void* net_ptr = nullptr;
    NetPeers net_peers;
    NetStreams net_streams;
    // the last enqueue that used the shared scratch: a following call on ANOTHER stream waits for it
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
};
DeviceCtx g_ctx[kMaxDevices];

DeviceCtx& cur_ctx() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return g_ctx[dev];
}
#define g_scratch (cur_ctx().scratch)
#define g_mutex (cur_ctx().mutex)
#define g_nonfinite_flag (cur_ctx().nonfinite_flag)
#define g_net_key (cur_ctx().net_key)
#define g_net_ptr (cur_ctx().net_ptr)
#define g_net_peers (cur_ctx().net_peers)

// Stream order across calls that share the device's scratch: `st` first waits for the previous enqueue if that
// ran on a different stream; `scratch_done` marks the end of this one.
int scratch_begin(cudaStream_t st) {
    DeviceCtx& c = cur_ctx();
    if (c.has_last && c.last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, c.ev_last, 0));
    return PCG_OK;
}
int scratch_done(cudaStream_t st) {
    DeviceCtx& c = cur_ctx();
    if (!c.ev_last) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_last, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(c.ev_last, st));
    c.last_stream = st;
    c.has_last = true;
    return PCG_OK;
}

int lib_stream(cudaStream_t* out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(PCG_ERR_ARG, "device index %d out of range", dev);
    DeviceCtx& c = g_ctx[dev];
    if (!c.stream) CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    *out = c.stream;
    return PCG_OK;
}


}  // namespace

std::atomic<unsigned long long> g_volume_uid{0};

struct pcg_volume {
    // never reused, unlike the handle's address: what the network table cache is keyed by
    unsigned long long uid = ++g_volume_uid;
    int ne = 0;
    int nfield = 0;
    int dtype = PCG_VALUE_F64;
    int device = 0;
    double* d_az = nullptr;
    double* d_rng = nullptr;
    void* d_val[kMaxFields] = {nullptr, nullptr, nullptr, nullptr};
    SweepDesc* d_sweeps = nullptr;
    PairDesc* d_pairs = nullptr;     // [ne-1], packed (F32) volumes only
    FastSweep* d_fs = nullptr;       // [ne]   r2-space descriptors (packed volumes only)
    double2* d_gt = nullptr;         // gate thresholds in r2 space, same offsets as the range arena
    GateAux* d_ht = nullptr;
    double2* d_azt = nullptr;        // {A[i], 1 / gap_i} per ray, same offsets as the azimuth arena
    // last (height, Re, levels) threshold table: rebuilt only when the call geometry changes
    mutable std::mutex ls_mutex;
    mutable std::vector<double> ls_key;
    mutable LevelSweep* d_ls = nullptr;      // followed in the same allocation by the LevelPair table
    mutable std::vector<unsigned char> ls_slow;
    mutable std::vector<LevelSweep> ls_host;  // host copy of the same table (rows go into the kernel parameters)
    std::vector<FastSweep> fs_host;           // host copy of d_fs (pair polynomials go into the kernel parameters)
    bool uniform = false;            // one range table shared by every sweep, every sweep / pair usable + validated
    unsigned uni_rng_off = 0;
    int uni_ngm1 = 0;
    float uni_inv_step = 0.f, uni_c0 = 0.f, uni_step = 0.f, uni_r_first = 0.f;
    std::vector<SweepDesc> sweeps;
    std::vector<PairDesc> pairs;     // host copy of d_pairs (the network path rescales it per call)
    double fill = -999.0;            // the missing-value marker the moments were packed with
    size_t bytes = 0;
    // valid (finite, != fill) source values per moment: the range-sanity window of the network path
    double src_min[kMaxFields] = {0, 0, 0, 0};
    double src_max[kMaxFields] = {0, 0, 0, 0};
    unsigned long long src_count[kMaxFields] = {0, 0, 0, 0};
};

namespace {

VolumeView view_of(const pcg_volume* v) {
    VolumeView w;
    w.az = v->d_az;
    w.rng = v->d_rng;
    for (int f = 0; f < kMaxFields; ++f) w.val[f] = v->d_val[f];
    w.sweeps = v->d_sweeps;
    w.ne = v->ne;
    w.nfield = v->nfield;
    return w;
}

int check_volume(const pcg_volume* v) {
    if (!v) return fail(PCG_ERR_ARG, "volume handle is NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != v->device)
        return fail(PCG_ERR_ARG, "volume lives on device %d but the current device is %d", v->device, dev);
    return PCG_OK;
}

int check_common(double re, int nx, int ny) {
    if (!(re > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    if (nx < 0 || ny < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    return PCG_OK;
}

void fill_level(LevelConst* lc, double z, double Re, double a2, double twoa) {
    // RadarGridC.pyx:145-146: gate_radius = Re + z; a*a + g*g - 2.0*a*g*cos(phi)
    volatile double g = Re + z;
    volatile double g2 = g * g;
    volatile double a2g2 = a2 + g2;
    volatile double twoag = twoa * g;
    lc->z = z;
    lc->g2 = g2;
    lc->a2g2 = a2g2;
    lc->twoag = twoag;
}

template <typename VT, bool EMIT>
int launch_cappi_nf(const CappiArgs& args, int nf, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    switch (nf) {
        case 1: cappi3d_kernel<VT, 1, EMIT><<<grid, block, smem, st>>>(args); break;
        case 2: cappi3d_kernel<VT, 2, EMIT><<<grid, block, smem, st>>>(args); break;
        case 3: cappi3d_kernel<VT, 3, EMIT><<<grid, block, smem, st>>>(args); break;
        case 4: cappi3d_kernel<VT, 4, EMIT><<<grid, block, smem, st>>>(args); break;
        default: return fail(PCG_ERR_ARG, "nfield must be 1..%d", kMaxFields);
    }
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}


// Elevation weight of sweep pair (k, k+1) as a degree-5 polynomial in u = c - c_mid, where
// c = cos(pi/2 + el) = -sin(el): w(u) = (el(u) - e_k) / (e_{k+1} - e_k), el(u) = -asin(c_mid + u)
// in degrees (RadarGridC.pyx:155-157,469-476).  Interpolated at Chebyshev nodes in long double and
// validated on a dense set; pairs that miss 1e-9 (or have e_{k+1} <= e_k) are marked !ok and their
// voxels take the reference's own acos chain.
void fit_pair(PairDesc* pr, double e0, double e1, double c0, double c1) {
    memset(pr, 0, sizeof(*pr));
    pr->c_lo = c0;
    pr->c_hi = c1;
    pr->c_mid = 0.5 * (c0 + c1);
    if (!(e1 > e0) || !(c0 > c1)) return;
    const long double mid = pr->c_mid;
    const long double half = 0.5L * ((long double)c0 - (long double)c1) * 1.02L + 1e-12L;
    const long double r2d = 180.0L / 3.14159265358979323846264338327950288L;
    auto wfun = [&](long double u) -> long double {
        long double c = mid + u;
        if (c > 1.0L) c = 1.0L;
        if (c < -1.0L) c = -1.0L;
        return (-asinl(c) * r2d - (long double)e0) / ((long double)e1 - (long double)e0);
    };
    const int n = 6;
    long double A[n][n + 1];
    for (int j = 0; j < n; ++j) {
        const long double t = cosl((2 * j + 1) * 3.14159265358979323846264338327950288L / (2 * n));
        long double pw = 1.0L;
        for (int i = 0; i < n; ++i) { A[j][i] = pw; pw *= t; }
        A[j][n] = wfun(t * half);
    }
    for (int c = 0; c < n; ++c) {           // Gaussian elimination with partial pivoting
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        for (int k = 0; k <= n; ++k) { long double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const long double f = A[r][c] / A[c][c];
            for (int k = c; k <= n; ++k) A[r][k] -= f * A[c][k];
        }
    }
    long double hp = 1.0L;
    for (int i = 0; i < n; ++i) { pr->q[i] = (double)(A[i][n] / A[i][i] / hp); hp *= half; }
    long double worst = 0.0L;
    for (int j = 0; j <= 128; ++j) {
        const long double u = half * (2.0L * j / 128.0L - 1.0L);
        long double v = 0.0L;
        for (int i = n - 1; i >= 0; --i) v = v * u + (long double)pr->q[i];
        const long double e = fabsl(v - wfun(u));
        if (e > worst) worst = e;
    }
    if (worst <= 1e-9L) pr->flags |= kPairPolyOk;
}


// ---- r2-space tables (pcg_r2.cuh) ----------------------------------------------------------------

// smallest double t with sqrt(t) >= R  (r >= R  <=>  r2 >= t: the IEEE square root is monotone)
double sqrt_threshold_ge(double R) {
    if (!(R > 0.0)) return 0.0;
    volatile double t = R * R;
    while (sqrt((double)t) >= R) t = nextafter((double)t, -HUGE_VAL);
    while (sqrt((double)t) < R) t = nextafter((double)t, HUGE_VAL);
    return t;
}

// smallest double t with sqrt(t) > R   (r <= R  <=>  r2 < t)
double sqrt_threshold_gt(double R) {
    if (!(R >= 0.0)) return 0.0;
    volatile double t = R * R;
    while (sqrt((double)t) > R) t = nextafter((double)t, -HUGE_VAL);
    while (!(sqrt((double)t) > R)) t = nextafter((double)t, HUGE_VAL);
    return t;
}

// w(u) = (el(C_k + u) - e_k) / (e_{k+1} - e_k), el(c) = -asin(c) in degrees, as u * (q1 + .. + q5 u^4):
// interpolated at Chebyshev nodes in long double, validated with the float-rounded coefficients.
bool fit_pair_r2(FastSweep* f, double e0, double e1) {
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double c0 = -sinl((long double)e0 * pi / 180.0L), c1 = -sinl((long double)e1 * pi / 180.0L);
    if (!(e1 > e0) || !(c0 > c1)) return false;
    const long double span = (c1 - c0);                 // < 0
    auto wfun = [&](long double u) -> long double {
        long double c = c0 + u;
        if (c > 1.0L) c = 1.0L;
        if (c < -1.0L) c = -1.0L;
        return (-asinl(c) * 180.0L / pi - (long double)e0) / ((long double)e1 - (long double)e0);
    };
    const int n = 5;
    long double A[n][n + 1];
    for (int j = 0; j < n; ++j) {
        // nodes t in (0, 1.02): u = t * span
        const long double t = 0.51L * (1.0L + cosl((2 * j + 1) * pi / (2 * n)));
        long double pw = 1.0L;
        for (int i = 0; i < n; ++i) { A[j][i] = pw; pw *= t; }
        A[j][n] = wfun(t * span) / (t * span);
    }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        for (int k = 0; k <= n; ++k) { long double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const long double g = A[r][c] / A[c][c];
            for (int k = c; k <= n; ++k) A[r][k] -= g * A[c][k];
        }
    }
    float q[n];
    long double sp = 1.0L;
    for (int i = 0; i < n; ++i) { q[i] = (float)(A[i][n] / A[i][i] / sp); sp *= span; }
    long double worst = 0.0L;
    for (int j = 0; j <= 256; ++j) {
        const long double u = span * 1.01L * j / 256.0L;
        long double v = 0.0L;
        for (int i = n - 1; i >= 0; --i) v = v * u + (long double)q[i];
        v *= u;
        const long double e = fabsl(v - wfun(u));
        if (e > worst) worst = e;
    }
    f->q1 = q[0]; f->q2 = q[1]; f->q3 = q[2]; f->q4 = q[3]; f->q5 = q[4];
    return worst <= 1e-7L;
}

// Per (level, sweep) thresholds for one radar: ls[(nz)][ne + 2], flags in slow_bits (bit per level).
void build_level_sweeps(const pcg_volume* v, double Re, double av, const double* levels, int nz,
                        LevelSweep* ls, size_t row_stride, std::vector<unsigned char>* slow) {
    const int ne = v->ne;
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double a = (long double)av;
    slow->assign(nz, 0);
    long double cmax = -2.0L;                                  // C_k = -sin(e_k) of the LOWEST sweep
    for (int k = 0; k < ne; ++k) {
        const long double ck = -sinl((long double)v->sweeps[k].elev * pi / 180.0L);
        if (ck > cmax) cmax = ck;
    }
    const double inf = HUGE_VAL;
    for (int iz = 0; iz < nz; ++iz) {
        LevelSweep* row = ls + (size_t)iz * row_stride;
        memset(row, 0, sizeof(LevelSweep) * (ne + 2));
        row[0].t_lo = inf; row[0].t_hi = inf;                  // "sweep -1": always el >= e_-1, never below it

        row[ne + 1].t_lo = -inf; row[ne + 1].t_hi = 0.0;       // "sweep ne": never reached; 0 so that r2 <= 0 confirms nothing
        volatile double gd = Re + levels[iz];                  // the reference's gate radius
        const long double g = (long double)gd;
        if (g > a) {
            const long double D = (g - a) * (g + a) / (2.0L * a);
            for (int k = 0; k < ne; ++k) {
                const long double e = (long double)v->sweeps[k].elev * pi / 180.0L;
                const long double se = sinl(e), ce = cosl(e);
                const long double rho = -a * se + sqrtl((g - a * ce) * (g + a * ce));
                LevelSweep& L = row[k + 1];
                if (!(rho > 0.0L) || !(fabsl(e) < 1.55L)) { (*slow)[iz] = 1; continue; }
                const long double rho2 = rho * rho;
                // bound on |r2_ref-equivalent - r2_exact| at the decision point (pcg_r2.cuh), with 4x safety:
                // the reference's numerator a2 + r2 - g2 carries ~2 ulp(max(a^2, g^2)) of rounding, worth twice
                // that in r2; its division / acos / scaling ~8e-16 in the cosine, worth 4 a rho times that
                const long double big = (g > a ? g * g : a * a);
                const long double margin = 16.0L * 2.220446049250313e-16L * big + 1.3e-14L * a * rho + 1e-14L * rho2;
                L.t_lo = nextafter((double)(rho2 - margin), -HUGE_VAL);
                L.t_hi = nextafter((double)(rho2 + margin), HUGE_VAL);
                L.rho2 = (double)rho2;
                L.rho = (float)rho;
                L.b1 = (float)(D / rho);
            }
        } else {
            // level at or below the antenna: el <= 0 everywhere; if the lowest sweep is clearly above the
            // largest reachable elevation every voxel is "below the lowest sweep", else take the slow chain
            const long double d2 = (a - g) * (a + g);
            const long double cmin = sqrtl(d2 > 0.0L ? d2 : 0.0L) / a;
            if (cmin - cmax > 1e-6L) {
                for (int k = 0; k < ne; ++k) { row[k + 1].t_lo = -1.0; row[k + 1].t_hi = -1.0; }
            } else {
                (*slow)[iz] = 1;
            }
        }
    }
}

// {t_lo[k], t_hi[k+1], rho2[k], rho[k], b1[k]} per (level, k + 1) from the per-sweep bands
// One row per level: [LevelPair x (ne + 1)][LevelSweep x (ne + 2)].  `rows` already holds the LevelSweep
// halves (build_level_sweeps with row_stride = 2 ne + 3 and ls = rows + ne + 1); this fills the pair halves
// {t_lo[k], t_hi[k+1], rho2[k], rho[k], b1[k]} and rewrites flagged levels so that nothing is ever confirmed
// there and the walk ends in the reference chain at once.
void build_level_pairs(LevelSweep* rows, int nz, int ne, const std::vector<unsigned char>& slow) {
    static_assert(sizeof(LevelSweep) == 32 && sizeof(LevelPair) == 32, "threshold tables are 32-byte records");
    const size_t stride = (size_t)2 * ne + 3;
    for (int iz = 0; iz < nz; ++iz) {
        LevelPair* out = reinterpret_cast<LevelPair*>(rows + (size_t)iz * stride);
        LevelSweep* row = rows + (size_t)iz * stride + (ne + 1);
        if (slow[iz])
            for (int s = 0; s < ne + 2; ++s) { row[s].t_lo = -HUGE_VAL; row[s].t_hi = HUGE_VAL; }
        for (int s = 0; s <= ne; ++s) {          // s = k + 1
            out[s].t_lo = slow[iz] ? -HUGE_VAL : row[s].t_lo;
            out[s].t_hi = row[s + 1].t_hi;
            out[s].rho2 = row[s].rho2;
            out[s].rho = row[s].rho;
            out[s].b1 = row[s].b1;
        }
    }
}

template <int NF, bool EMIT, bool UNI, bool LAZY>
int launch_fast_one(const R2Args& args, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024)   // opt in to large dynamic shared memory (per device, cheap)
        CUDA_TRY(cudaFuncSetAttribute(cappi3d_r2_kernel<NF, EMIT, UNI, LAZY>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cappi3d_r2_kernel<NF, EMIT, UNI, LAZY><<<grid, block, smem, st>>>(args);
    return PCG_OK;
}

template <int NF, bool EMIT>
int launch_fast_flags(const R2Args& args, bool uni, bool lazy, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    if (uni) return lazy ? launch_fast_one<NF, EMIT, true, true>(args, grid, block, smem, st)
                         : launch_fast_one<NF, EMIT, true, false>(args, grid, block, smem, st);
    return lazy ? launch_fast_one<NF, EMIT, false, true>(args, grid, block, smem, st)
                : launch_fast_one<NF, EMIT, false, false>(args, grid, block, smem, st);
}

// kernel v8 (pcg_r3.cuh): a [sweep] azimuth table per column up to 16 sweeps, the current pair's brackets in
// registers beyond (LAZY; threshold rows in global memory)
template <int NF, bool EMIT, bool UNI, bool MERGE, bool CT, bool LAZY>
int launch_r3_one(const R3Args& args, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(cappi3d_r3_kernel<NF, EMIT, UNI, MERGE, CT, LAZY>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cappi3d_r3_kernel<NF, EMIT, UNI, MERGE, CT, LAZY><<<grid, block, smem, st>>>(args);
    return PCG_OK;
}

template <int NF, bool EMIT, bool CT, bool LAZY>
int launch_r3_flags(const R3Args& args, bool uni, bool merge, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    if (uni) return merge ? launch_r3_one<NF, EMIT, true, true, CT, LAZY>(args, grid, block, smem, st)
                          : launch_r3_one<NF, EMIT, true, false, CT, LAZY>(args, grid, block, smem, st);
    return merge ? launch_r3_one<NF, EMIT, false, true, CT, LAZY>(args, grid, block, smem, st)
                 : launch_r3_one<NF, EMIT, false, false, CT, LAZY>(args, grid, block, smem, st);
}

template <int NF, bool EMIT>
int launch_r3_ct(const R3Args& args, bool ct, bool uni, bool merge, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    if (args.a.lazy_az) return launch_r3_flags<NF, EMIT, false, true>(args, uni, merge, grid, block, smem, st);
    return ct ? launch_r3_flags<NF, EMIT, true, false>(args, uni, merge, grid, block, smem, st)
              : launch_r3_flags<NF, EMIT, false, false>(args, uni, merge, grid, block, smem, st);
}

template <bool EMIT>
int launch_r3_nf(const R3Args& args, bool ct, int nf, bool uni, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    const bool merge = args.a.base.merge_max != 0;
    int rc;
    switch (nf) {
        case 1: rc = launch_r3_ct<1, EMIT>(args, ct, uni, merge, grid, block, smem, st); break;
        case 2: rc = launch_r3_ct<2, EMIT>(args, ct, uni, merge, grid, block, smem, st); break;
        case 3: rc = launch_r3_ct<3, EMIT>(args, ct, uni, merge, grid, block, smem, st); break;
        case 4: rc = launch_r3_ct<4, EMIT>(args, ct, uni, merge, grid, block, smem, st); break;
        default: return fail(PCG_ERR_ARG, "nfield must be 1..%d", kMaxFields);
    }
    if (rc != PCG_OK) return rc;
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}

// PCG_KERNEL=r2 keeps the v7 kernel (A/B runs); default: v8 wherever it applies
bool use_r3_kernel() {
    static const int choice = [] {
        const char* e = getenv("PCG_KERNEL");
        return (e && strcmp(e, "r2") == 0) ? 0 : 1;
    }();
    return choice != 0;
}

// PCG_LAZY_KERNEL=r2 keeps the v7 kernel for volumes with more than 16 sweeps (A/B runs)
bool use_r3_lazy_kernel() {
    static const int choice = [] {
        const char* e = getenv("PCG_LAZY_KERNEL");
        return (e && strcmp(e, "r2") == 0) ? 0 : 1;
    }();
    return choice != 0;
}

// PCG_TABLES=global keeps the threshold rows in global memory (A/B runs)
bool use_const_tables() {
    static const int choice = [] {
        const char* e = getenv("PCG_TABLES");
        return (e && strcmp(e, "global") == 0) ? 0 : 1;
    }();
    return choice != 0;
}

template <bool EMIT>
int launch_fast_nf(const R2Args& args, int nf, bool uni, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    const bool lazy = args.lazy_az != 0;
    int rc;
    switch (nf) {
        case 1: rc = launch_fast_flags<1, EMIT>(args, uni, lazy, grid, block, smem, st); break;
        case 2: rc = launch_fast_flags<2, EMIT>(args, uni, lazy, grid, block, smem, st); break;
        case 3: rc = launch_fast_flags<3, EMIT>(args, uni, lazy, grid, block, smem, st); break;
        case 4: rc = launch_fast_flags<4, EMIT>(args, uni, lazy, grid, block, smem, st); break;
        default: return fail(PCG_ERR_ARG, "nfield must be 1..%d", kMaxFields);
    }
    if (rc != PCG_OK) return rc;
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}

int enqueue_cappi_fast(const pcg_volume* vol, double h, double Re, const double* d_gx, const double* d_gy,
                       long long ncol, int ny, const double* levels, int nz, double fill, int blind_code,
                       double radar_x, double radar_y, int shift, int merge_max, double* d_out,
                       int32_t* d_idx, cudaStream_t st, int nfield) {
    if (!(fill == vol->fill))
        return fail(PCG_ERR_ARG, "fillvalue %.17g differs from the value the volume was packed with (%.17g)",
                    fill, vol->fill);
    R2Args RA;
    memset(&RA, 0, sizeof(RA));
    FastArgs& A = RA.base;
    A.vol = view_of(vol);
    A.pairs = vol->d_pairs;
    A.grid.gx = d_gx; A.grid.gy = d_gy; A.grid.ncol = ncol;
    A.grid.radar_x = radar_x; A.grid.radar_y = radar_y; A.grid.shift = shift;
    // warp footprint on the (nx, ny) grid: 4 rows x 8 columns when the grid is at least 4 rows tall (a compact
    // patch touches fewer rays / gate lines per gather), else 1 x 32
    const int nxg = (int)(ncol / (ny > 0 ? ny : 1));
    A.grid.ny = ny; A.grid.nx = nxg;
    A.grid.py_log2 = (nxg >= 4 && ny >= 8) ? 3 : 5;
    volatile double av = Re + h;
    volatile double a2 = av * av;
    volatile double twoa = 2.0 * av;
    A.a = av; A.a2 = a2; A.twoa = twoa; A.Re = Re; A.fill = fill; A.fill32 = (float)fill;
    A.guard = kGuardDelta * twoa;
    A.inv_twoa = 1.0 / twoa;
    A.nearest_gate = (blind_code == 1 || blind_code == 3 || blind_code == 4);
    A.lowest_sweep = (blind_code == 2 || blind_code == 3 || blind_code == 4);
    A.highest_sweep = (blind_code == 4);
    A.merge_max = merge_max;
    A.field_stride = (unsigned long long)nz * (unsigned long long)ncol;
    // r2-space thresholds of this (radar height, level set): built on the host, stream-ordered upload
    const int ne = vol->ne;
    LevelSweep* d_ls = nullptr;
    std::vector<LevelSweep> tab_host;
    {
        std::lock_guard<std::mutex> lk(vol->ls_mutex);
        std::vector<double> key;
        key.reserve(nz + 3);
        key.push_back(h); key.push_back(Re); key.push_back((double)nz);
        key.insert(key.end(), levels, levels + nz);
        if (!vol->d_ls || key.size() != vol->ls_key.size() ||
            memcmp(key.data(), vol->ls_key.data(), key.size() * sizeof(double)) != 0) {
            const size_t stride = (size_t)2 * ne + 3;
            std::vector<LevelSweep> ls((size_t)nz * stride);
            build_level_sweeps(vol, Re, av, levels, nz, ls.data() + (ne + 1), stride, &vol->ls_slow);
            build_level_pairs(ls.data(), nz, ne, vol->ls_slow);
            LevelSweep* fresh = nullptr;
            CUDA_TRY(g_pool.alloc((void**)&fresh, ls.size() * sizeof(LevelSweep)));
            cudaError_t e = cudaMemcpy(fresh, ls.data(), ls.size() * sizeof(LevelSweep), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { g_pool.release(fresh); return fail(PCG_ERR_CUDA, "threshold upload failed: %s", cudaGetErrorString(e)); }
            if (vol->d_ls) {
                cudaDeviceSynchronize();  // kernels may still be reading the previous table
                g_pool.release(vol->d_ls);
            }
            vol->d_ls = fresh;
            vol->ls_key.swap(key);
            vol->ls_host.swap(ls);
        }
        d_ls = vol->d_ls;
        tab_host = vol->ls_host;       // (a copy: another thread may rebuild the volume's table meanwhile)
    }
    RA.nonfinite = g_nonfinite_flag;
    RA.r2.fs = vol->d_fs;
    RA.r2.g_t = vol->d_gt;
    RA.r2.h_t = vol->d_ht;
    RA.r2.az_t = vol->d_azt;
    RA.r2.inv2a = (float)(1.0 / twoa);
    RA.r2.u_rng_off = vol->uni_rng_off; RA.r2.u_ngm1 = vol->uni_ngm1;
    RA.r2.u_inv_step = vol->uni_inv_step; RA.r2.u_c0 = vol->uni_c0;
    RA.r2.u_step = vol->uni_step; RA.r2.u_r_first = vol->uni_r_first;
    // shared memory: descriptors + the per-thread azimuth table [thread][sweep]; pick the block
    // size so that several blocks still fit an SM (227 KB)
    const size_t fixed = (size_t)vol->ne * (sizeof(SweepDesc) + sizeof(FastSweep)) + (size_t)(vol->ne - 1) * sizeof(PairDesc);
    // azimuth brackets: a [sweep] table per column up to 16 sweeps, on demand (2-entry cache) beyond
    RA.lazy_az = vol->ne > 16 ? 1 : 0;
    const bool r3 = use_r3_kernel() && (!RA.lazy_az || use_r3_lazy_kernel());
    const size_t per_thread = RA.lazy_az ? 2 * (sizeof(AzEntry) + sizeof(int) + (d_idx ? sizeof(int) : 0))
                                         : (size_t)vol->ne * (sizeof(AzEntry) + (d_idx ? sizeof(int) : 0));
    int threads = 256;
    const size_t fixed_all = fixed + (r3 ? (size_t)vol->ne * sizeof(AzDesc) : 0);
    while (threads > 32 && fixed_all + per_thread * threads > 56 * 1024) threads >>= 1;
    const size_t smem = fixed_all + per_thread * threads;
    const dim3 block(threads);
    // block tile: (warps/2 ... ) see cappi3d_r2_kernel: tile_y = 4 warps x PY columns, tile_x = rest
    const int PY = 1 << A.grid.py_log2, PX = 32 / PY, warps = threads / 32;
    const int wy = warps >= 4 ? 4 : warps, wx = warps / wy;
    const long long tiles_y = (ny + (long long)wy * PY - 1) / ((long long)wy * PY);
    const long long tiles_x = (nxg + (long long)wx * PX - 1) / ((long long)wx * PX);
    A.grid.tiles_y = (int)tiles_y;
    const dim3 grid((unsigned)(tiles_x * tiles_y));
    int rc = PCG_OK;
    // v8: the threshold rows of a launch's levels ride in the kernel parameters when at least 16 levels fit
    const size_t row_bytes = (size_t)(2 * ne + 3) * sizeof(LevelSweep);
    const int tab_levels = (int)(kR3TabBytes / row_bytes);
    const bool ct = r3 && use_const_tables() && tab_levels >= 16 && ne <= kR3PolySweeps &&
                    vol->fs_host.size() == (size_t)ne && tab_host.size() == (size_t)nz * (2 * ne + 3);
    const int per_launch = ct && tab_levels < kMaxLevelsPerLaunch ? tab_levels : kMaxLevelsPerLaunch;
    std::unique_ptr<R3Args> QA;
    if (r3) {
        QA.reset(new R3Args);
        memset(QA->poly, 0, sizeof(QA->poly));
        if (ct)
            for (int k = 0; k < ne; ++k) {
                const FastSweep& f = vol->fs_host[k];
                QA->poly[k][0] = f.q1; QA->poly[k][1] = f.q2; QA->poly[k][2] = f.q3; QA->poly[k][3] = f.q4;
                QA->poly[k][4] = f.q5;
            }
    }
    for (int z0 = 0; z0 < nz && rc == PCG_OK; z0 += per_launch) {
        const int cnt = (nz - z0 < per_launch) ? nz - z0 : per_launch;
        A.nz = cnt;
        for (int k = 0; k < cnt; ++k) fill_level(&A.levels[k], levels[z0 + k], Re, a2, twoa);
        RA.r2.lp = reinterpret_cast<const LevelPair*>(d_ls) + (size_t)z0 * (2 * ne + 3);
        A.out = d_out + (size_t)z0 * (size_t)ncol;
        A.idx = d_idx ? d_idx + (size_t)z0 * (size_t)ncol * kIdxStride : nullptr;
        if (r3) {
            QA->a = RA;
            if (ct) memcpy(QA->tab, tab_host.data() + (size_t)z0 * (2 * ne + 3), (size_t)cnt * row_bytes);
            rc = d_idx ? launch_r3_nf<true>(*QA, ct, nfield, vol->uniform, grid, block, smem, st)
                       : launch_r3_nf<false>(*QA, ct, nfield, vol->uniform, grid, block, smem, st);
        } else {
            rc = d_idx ? launch_fast_nf<true>(RA, nfield, vol->uniform, grid, block, smem, st)
                       : launch_fast_nf<false>(RA, nfield, vol->uniform, grid, block, smem, st);
        }
    }
    return rc;
}

int enqueue_cappi(const pcg_volume* vol, double h, double Re, const double* d_gx, const double* d_gy,
                  long long ncol, int ny, const double* levels, int nz, double fill, int blind_code,
                  double beam_width_deg, double radar_x, double radar_y, int shift, int merge_max,
                  double* d_out, int32_t* d_idx, cudaStream_t st, int nfield_override = 0) {
    const int nfield = nfield_override > 0 ? nfield_override : vol->nfield;
    if (blind_code < 0 || blind_code > 4)
        return fail(PCG_ERR_ARG,
                    "blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', 'hybrid', or "
                    "'nearest_valid'.");
    if (nz < 0) return fail(PCG_ERR_ARG, "nz must be non-negative");
    if (ncol == 0 || nz == 0) return PCG_OK;
    if (vol->ne > kMaxSweeps) return fail(PCG_ERR_ARG, "at most %d sweeps are supported", kMaxSweeps);
    if (vol->dtype == PCG_VALUE_F32)
        return enqueue_cappi_fast(vol, h, Re, d_gx, d_gy, ncol, ny, levels, nz, fill, blind_code, radar_x, radar_y,
                                  shift, merge_max, d_out, d_idx, st, nfield);
    CappiArgs a;
    memset(&a, 0, sizeof(a));
    a.vol = view_of(vol);
    a.grid.gx = d_gx;
    a.grid.gy = d_gy;
    a.grid.ncol = ncol;
    a.grid.radar_x = radar_x;
    a.grid.radar_y = radar_y;
    a.grid.shift = shift;
    volatile double av = Re + h;           // RadarGridC.pyx:144
    volatile double a2 = av * av;
    volatile double twoa = 2.0 * av;
    a.a = av; a.a2 = a2; a.twoa = twoa; a.Re = Re; a.fill = fill;
    if (beam_width_deg <= 0.0) beam_width_deg = 1.0;     // RadarGridC.pyx:400-404
    double beam_half = beam_width_deg * 0.5;
    if (beam_half < 0.1) beam_half = 0.1;
    a.beam_half = beam_half;
    a.nearest_gate = (blind_code == 1 || blind_code == 3 || blind_code == 4);
    a.lowest_sweep = (blind_code == 2 || blind_code == 3 || blind_code == 4);
    a.highest_sweep = (blind_code == 4);
    a.merge_max = merge_max;
    a.field_stride = (unsigned long long)nz * (unsigned long long)ncol;
    const size_t smem = (size_t)(vol->ne > 0 ? vol->ne : 1) * sizeof(SweepDesc);
    const dim3 block(256);
    const dim3 grid((unsigned)((ncol + 255) / 256));
    for (int z0 = 0; z0 < nz; z0 += kMaxLevelsPerLaunch) {
        const int cnt = (nz - z0 < kMaxLevelsPerLaunch) ? nz - z0 : kMaxLevelsPerLaunch;
        a.nz = cnt;
        for (int k = 0; k < cnt; ++k) fill_level(&a.levels[k], levels[z0 + k], Re, a2, twoa);
        a.out = d_out + (size_t)z0 * (size_t)ncol;
        a.idx = d_idx ? d_idx + (size_t)z0 * (size_t)ncol * kIdxStride : nullptr;
        const int rc = d_idx ? launch_cappi_nf<double, true>(a, nfield, grid, block, smem, st)
                             : launch_cappi_nf<double, false>(a, nfield, grid, block, smem, st);
        if (rc != PCG_OK) return rc;
    }
    return PCG_OK;
}

int enqueue_cr(const pcg_volume* vol, double h, double Re, const double* d_gx, const double* d_gy,
               long long ncol, double fill, int single_sweep, double elevation_deg, int blind_code,
               double radar_x, double radar_y, int shift, double* d_out, int32_t* d_idx,
               cudaStream_t st) {
    if (ncol == 0) return PCG_OK;
    if (vol->ne > kMaxSweeps) return fail(PCG_ERR_ARG, "at most %d sweeps are supported", kMaxSweeps);
    if (vol->dtype == PCG_VALUE_F32) {
        if (single_sweep >= 0) return fail(PCG_ERR_ARG, "ppi_to_grid needs a volume packed with PCG_VALUE_F64");
        if (!(fill == vol->fill))
            return fail(PCG_ERR_ARG, "fillvalue %.17g differs from the value the volume was packed with (%.17g)",
                        fill, vol->fill);
        CrFastArgs f;
        memset(&f, 0, sizeof(f));
        f.vol = view_of(vol);
        f.grid.gx = d_gx; f.grid.gy = d_gy; f.grid.ncol = ncol;
        f.grid.radar_x = radar_x; f.grid.radar_y = radar_y; f.grid.shift = shift;
        volatile double fav = Re + h;
        volatile double fa2 = fav * fav;
        volatile double ftwoa = 2.0 * fav;
        f.a = fav; f.a2 = fa2; f.twoa = ftwoa; f.Re = Re; f.fill = fill; f.fill32 = (float)fill;
        f.out = d_out;
        f.nonfinite = g_nonfinite_flag;
        // small grids: two warps share the sweeps of every 32 columns (see cr_fast_kernel)
        const int W = (vol->ne >= 2 && ncol < 148LL * 2048) ? 2 : 1;
        const long long per_block = 32LL * (kCrWarps / W);
        cr_fast_kernel<<<(unsigned)((ncol + per_block - 1) / per_block), 32 * kCrWarps,
                         (size_t)vol->ne * sizeof(SweepDesc), st>>>(f, W);
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        return PCG_OK;
    }
    CrArgs a;
    memset(&a, 0, sizeof(a));
    a.vol = view_of(vol);
    a.grid.gx = d_gx;
    a.grid.gy = d_gy;
    a.grid.ncol = ncol;
    a.grid.radar_x = radar_x;
    a.grid.radar_y = radar_y;
    a.grid.shift = shift;
    volatile double av = Re + h;
    volatile double a2 = av * av;
    volatile double twoa = 2.0 * av;
    a.a = av; a.a2 = a2; a.twoa = twoa; a.Re = Re; a.fill = fill;
    a.single_sweep = single_sweep;
    a.tan_override = tan(elevation_deg * kDeg2Rad);   // RadarGridC.pyx:110,113
    a.nearest_gate = (blind_code == 1 || blind_code == 3 || blind_code == 4);
    a.out = d_out;
    a.idx = d_idx;
    const size_t smem = (size_t)(vol->ne > 0 ? vol->ne : 1) * sizeof(SweepDesc);
    const dim3 block(256);
    const dim3 grid((unsigned)((ncol + 255) / 256));
    if (d_idx) cr_kernel<double, true><<<grid, block, smem, st>>>(a);
    else cr_kernel<double, false><<<grid, block, smem, st>>>(a);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}

}  // namespace

// the product wrappers' `np.where(grid == fillvalue, np.nan, grid)` (NRadar.py:1339,1392,...) on the device, before
// the result crosses PCIe: a host pass over a 461 MB volume costs more than the whole gridding call
__global__ void relabel_kernel(double* __restrict__ data, size_t n, double from, double to) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (data[i] == from) data[i] = to;
}

static int relabel_missing(double* d_data, size_t n, double fill, double out_missing, cudaStream_t st) {
    if (n == 0 || memcmp(&fill, &out_missing, sizeof(double)) == 0) return PCG_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    relabel_kernel<<<blocks, 256, 0, st>>>(d_data, n, fill, out_missing);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}

extern "C" {

const char* pcg_last_error(void) { return g_err; }

const char* pcg_version(void) { return "pycwr_b200 0.1 (sm_100a)"; }

int pcg_device_count(int* count) {
    if (!count) return fail(PCG_ERR_ARG, "count is NULL");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return PCG_OK;
}

int pcg_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return PCG_OK;
}

int pcg_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) return fail(PCG_ERR_ARG, "ptr is NULL");
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(PCG_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return PCG_OK;
}

int pcg_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return PCG_OK;
}

int64_t pcg_launch_count(void) { return (int64_t)g_launches.load(); }

}  // extern "C"

// Shared body of pcg_volume_create / pcg_volume_create_raw.  `az` is sorted; when `ray_order` is given the
// moments are still in scan order (row i of the packed sweep is raw row ray_order[s][i]) with NaN for
// missing, `val_ld[s]` elements between consecutive raw rows: the permutation, the range clip and the
// NaN -> fill substitution of the reference's packer happen inside the pack kernels.
static int volume_create_impl(int ne, const int* naz, const int* ng, const double* const* az,
                              const double* const* rng, int nfield, const double* const* const* val,
                              const long long* val_ld, const int* const* ray_order,
                              const double* fix_elev, double fill, int value_dtype, pcg_volume** out) {
    const bool raw = ray_order != nullptr;
    if (!out) return fail(PCG_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (ne < 0 || ne > kMaxSweeps) return fail(PCG_ERR_ARG, "ne must be in [0, %d]", kMaxSweeps);
    if (nfield < 1 || nfield > kMaxFields) return fail(PCG_ERR_ARG, "nfield must be in [1, %d]", kMaxFields);
    if (value_dtype != PCG_VALUE_F64 && value_dtype != PCG_VALUE_F32)
        return fail(PCG_ERR_ARG, "value_dtype must be PCG_VALUE_F64 or PCG_VALUE_F32");
    if (ne > 0 && (!naz || !ng || !az || !rng || !val || !fix_elev))
        return fail(PCG_ERR_ARG, "NULL table pointer");
    for (int s = 0; s < ne; ++s) {
        if (naz[s] < 0 || ng[s] < 0) return fail(PCG_ERR_ARG, "negative table length in sweep %d", s);
        for (int f = 0; f < nfield; ++f)
            if ((size_t)naz[s] * (size_t)ng[s] > 0 && (!val[f] || !val[f][s]))
                return fail(PCG_ERR_ARG, "NULL value array (field %d, sweep %d)", f, s);
    }
    pcg_volume* v = new (std::nothrow) pcg_volume();
    if (!v) return fail(PCG_ERR_NOMEM, "out of host memory");
    v->ne = ne;
    v->nfield = nfield;
    v->fill = fill;
    cudaError_t e = cudaGetDevice(&v->device);
    if (e != cudaSuccess) { delete v; return fail(PCG_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e)); }

    // ---- descriptors; identical range tables share storage (one bracket serves both sweeps) ----
    size_t n_az = 0, n_rng = 0, n_val = 0;
    bool strict = true;        // every table strictly increasing (no duplicate coordinates)
    v->sweeps.resize(ne > 0 ? ne : 1);
    memset(v->sweeps.data(), 0, v->sweeps.size() * sizeof(SweepDesc));
    std::vector<int> rng_owner(ne > 0 ? ne : 1, -1);
    for (int s = 0; s < ne; ++s) {
        SweepDesc& d = v->sweeps[s];
        d.naz = naz[s];
        d.ng = ng[s];
        d.az_off = (unsigned)n_az;
        d.val_off = (unsigned long long)n_val;
        d.elev = fix_elev[s];
        d.tan_e = tan(fix_elev[s] * kDeg2Rad);     // RadarGridC.pyx:110,113 (host glibc)
        d.cthr = -sin(fix_elev[s] * kDeg2Rad);     // cos(pi/2 + el): elevation threshold in cosine space
        d.sorted = 1;
        if (s > 0 && !(fix_elev[s - 1] <= fix_elev[s])) strict = false;   // packer sorts by fixed angle
        if (naz[s] > 0) {
            d.az_first = az[s][0];
            const double span = az[s][naz[s] - 1] - az[s][0];
            d.az_inv_step = (naz[s] > 1 && span > 0.0) ? (double)(naz[s] - 1) / span : 0.0;
            for (int i = 1; i < naz[s]; ++i) {
                if (!(az[s][i - 1] <= az[s][i])) d.sorted = 0;
                if (!(az[s][i - 1] < az[s][i])) strict = false;
            }
            // the wrap interval (az[last]-360, az[0]) must be non-degenerate too
            if (naz[s] > 1 && !(az[s][naz[s] - 1] - 360.0 < az[s][0])) strict = false;
        }
        if (ng[s] > 0) {
            d.r_first = rng[s][0];
            d.r_last = rng[s][ng[s] - 1];
            const double span = d.r_last - d.r_first;
            d.rng_inv_step = (ng[s] > 1 && span > 0.0) ? (double)(ng[s] - 1) / span : 0.0;
            for (int i = 1; i < ng[s]; ++i) {
                if (!(rng[s][i - 1] <= rng[s][i])) d.sorted = 0;
                if (!(rng[s][i - 1] < rng[s][i])) strict = false;
            }
        }
        if (!d.sorted) strict = false;
        for (int t = 0; t < s && rng_owner[s] < 0; ++t)
            if (rng_owner[t] == t && ng[t] == ng[s] &&
                (ng[s] == 0 || memcmp(rng[t], rng[s], (size_t)ng[s] * sizeof(double)) == 0))
                rng_owner[s] = t;
        if (rng_owner[s] < 0) {
            rng_owner[s] = s;
            d.rng_off = (unsigned)n_rng;
            n_rng += (size_t)ng[s];
        } else {
            d.rng_off = v->sweeps[rng_owner[s]].rng_off;
        }
        n_az += (size_t)naz[s];
        n_val += (size_t)naz[s] * (size_t)ng[s];
        n_val = (n_val + 1) & ~(size_t)1;   // keep every sweep tile 16-byte aligned
    }
    if (n_az > 0xffffffffull || n_rng > 0xffffffffull) { delete v; return fail(PCG_ERR_ARG, "tables too large"); }
    const bool fits32 = n_val < 0xfffffff0ull;     // packed layout addresses pair rows with 32 bits

    // packed (fast) storage needs >= 2 sweeps and clean tables; anything else keeps the float64
    // layout and runs the kernels that restate the reference operation by operation.
    // (the packed kernels widen float results back to double: the fill marker must survive the round trip)
    const bool packed = (value_dtype == PCG_VALUE_F32) && ne >= 2 && ne <= 120 && strict && fits32 &&
                        ((double)(float)fill == fill);
    v->dtype = packed ? PCG_VALUE_F32 : PCG_VALUE_F64;
    const size_t vsz = packed ? sizeof(float4) : sizeof(double);   // quad {A,B}[g-1], {A,B}[g] per (radial pair, gate)
    const size_t rsz = packed ? sizeof(double2) : sizeof(double);

    auto cleanup = [&]() { pcg_volume_destroy(v); };
#define VOL_TRY(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            cleanup();                                                                         \
            return fail(e__ == cudaErrorMemoryAllocation ? PCG_ERR_NOMEM : PCG_ERR_CUDA,       \
                        "%s failed: %s", #expr, cudaGetErrorString(e__));                      \
        }                                                                                      \
    } while (0)
    VOL_TRY(g_pool.alloc((void**)&v->d_az, (n_az ? n_az : 1) * sizeof(double)));
    // (+2 entries: the straight-line samplers read entries gi-1 and gi before they know the sweep is usable)
    VOL_TRY(g_pool.alloc((void**)&v->d_rng, (n_rng + 2) * rsz));
    VOL_TRY(cudaMemset(v->d_rng, 0, (n_rng + 2) * rsz));
    VOL_TRY(g_pool.alloc((void**)&v->d_sweeps, v->sweeps.size() * sizeof(SweepDesc)));
    for (int f = 0; f < nfield; ++f) VOL_TRY(g_pool.alloc(&v->d_val[f], (n_val ? n_val : 1) * vsz));
    v->bytes = n_az * sizeof(double) + n_rng * rsz + v->sweeps.size() * sizeof(SweepDesc) +
               (size_t)nfield * n_val * vsz;

    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if (lib_stream(&st) != PCG_OK) { cleanup(); return PCG_ERR_CUDA; }
    size_t max_sweep = 1;
    for (int s = 0; s < ne; ++s) {
        const size_t cnt = (size_t)naz[s] * (size_t)ng[s];
        if (cnt > max_sweep) max_sweep = cnt;
        if ((size_t)ng[s] > max_sweep) max_sweep = (size_t)ng[s];
    }
    unsigned long long* d_mm = nullptr;      // [nfield][3]: min key, max key, count
    if (g_scratch[S_TMP1].ensure(kMaxFields * 3 * sizeof(unsigned long long)) != PCG_OK) { cleanup(); return PCG_ERR_NOMEM; }
    d_mm = (unsigned long long*)g_scratch[S_TMP1].ptr;
    {
        unsigned long long init[kMaxFields * 3];
        for (int f = 0; f < kMaxFields; ++f) { init[3 * f] = ~0ull; init[3 * f + 1] = 0ull; init[3 * f + 2] = 0ull; }
        VOL_TRY(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    double* tmp = nullptr;
    int* d_order = nullptr;
    if (packed || raw) {
        // one float64 staging tile; uploads and pack kernels are ordered by the stream
        if (g_scratch[S_TMP0].ensure(max_sweep * sizeof(double)) != PCG_OK) { cleanup(); return PCG_ERR_NOMEM; }
        tmp = (double*)g_scratch[S_TMP0].ptr;
    }
    if (raw) {
        size_t max_naz = 1;
        for (int s = 0; s < ne; ++s) if ((size_t)naz[s] > max_naz) max_naz = (size_t)naz[s];
        if (g_scratch[S_TMP2].ensure(max_naz * sizeof(int)) != PCG_OK) { cleanup(); return PCG_ERR_NOMEM; }
        d_order = (int*)g_scratch[S_TMP2].ptr;
    }
    for (int s = 0; s < ne; ++s) {
        const SweepDesc& d = v->sweeps[s];
        if (d.naz) VOL_TRY(cudaMemcpyAsync(v->d_az + d.az_off, az[s], (size_t)d.naz * sizeof(double), cudaMemcpyHostToDevice, st));
        if (d.ng && rng_owner[s] == s) {
            if (packed) {
                VOL_TRY(cudaMemcpyAsync(tmp, rng[s], (size_t)d.ng * sizeof(double), cudaMemcpyHostToDevice, st));
                pack_range_kernel<<<(d.ng + 255) / 256, 256, 0, st>>>(tmp, (double2*)v->d_rng + d.rng_off, d.ng);
                g_launches.fetch_add(1);
                VOL_TRY(cudaGetLastError());
            } else {
                VOL_TRY(cudaMemcpyAsync(v->d_rng + d.rng_off, rng[s], (size_t)d.ng * sizeof(double), cudaMemcpyHostToDevice, st));
            }
        }
        const size_t cnt = (size_t)d.naz * (size_t)d.ng;
        if (!cnt) continue;
        if (raw) VOL_TRY(cudaMemcpyAsync(d_order, ray_order[s], (size_t)d.naz * sizeof(int), cudaMemcpyHostToDevice, st));
        for (int f = 0; f < nfield; ++f) {
            if (raw) {
                // scan-order rows (clipped to the first ng gates of each ray) -> staging tile -> packed layout
                const size_t ld = val_ld ? (size_t)val_ld[s] : (size_t)d.ng;
                VOL_TRY(cudaMemcpy2DAsync(tmp, (size_t)d.ng * sizeof(double), val[f][s], ld * sizeof(double),
                                          (size_t)d.ng * sizeof(double), (size_t)d.naz, cudaMemcpyHostToDevice, st));
                const unsigned blocks = (unsigned)((cnt + 255) / 256 < 148 * 16 ? (cnt + 255) / 256 : 148 * 16);
                if (packed)
                    pack_quads_kernel<<<blocks, 256, 0, st>>>(tmp, (float4*)v->d_val[f] + d.val_off, d.naz, d.ng,
                                                              fill, (float)fill, d_order, 1);
                else
                    permute_rows_kernel<<<blocks, 256, 0, st>>>(tmp, (double*)v->d_val[f] + d.val_off, d.naz, d.ng,
                                                                fill, d_order);
                minmax_kernel<<<blocks, 256, 0, st>>>(tmp, cnt, fill, d_mm + 3 * f);
                g_launches.fetch_add(2);
                VOL_TRY(cudaGetLastError());
            } else if (packed) {
                VOL_TRY(cudaMemcpyAsync(tmp, val[f][s], cnt * sizeof(double), cudaMemcpyHostToDevice, st));
                const unsigned blocks = (unsigned)((cnt + 255) / 256 < 148 * 16 ? (cnt + 255) / 256 : 148 * 16);
                pack_quads_kernel<<<blocks, 256, 0, st>>>(tmp, (float4*)v->d_val[f] + d.val_off, d.naz, d.ng,
                                                          fill, (float)fill, nullptr, 0);
                minmax_kernel<<<blocks, 256, 0, st>>>(tmp, cnt, fill, d_mm + 3 * f);
                g_launches.fetch_add(2);
                VOL_TRY(cudaGetLastError());
            } else {
                double* dst = (double*)v->d_val[f] + d.val_off;
                VOL_TRY(cudaMemcpyAsync(dst, val[f][s], cnt * sizeof(double), cudaMemcpyHostToDevice, st));
                const unsigned blocks = (unsigned)((cnt + 255) / 256 < 148 * 16 ? (cnt + 255) / 256 : 148 * 16);
                minmax_kernel<<<blocks, 256, 0, st>>>(dst, cnt, fill, d_mm + 3 * f);
                g_launches.fetch_add(1);
                VOL_TRY(cudaGetLastError());
            }
        }
    }
    VOL_TRY(cudaMemcpyAsync(v->d_sweeps, v->sweeps.data(), v->sweeps.size() * sizeof(SweepDesc), cudaMemcpyHostToDevice, st));
    std::vector<PairDesc> pairs;
    if (packed) {
        pairs.resize(ne - 1);
        for (int k = 0; k + 1 < ne; ++k) {
            const SweepDesc& d0 = v->sweeps[k];
            const SweepDesc& d1 = v->sweeps[k + 1];
            fit_pair(&pairs[k], d0.elev, d1.elev, d0.cthr, d1.cthr);
            pairs[k].r_first = d0.r_first;
            pairs[k].r_last = d0.r_last;
            pairs[k].inv_step = d0.rng_inv_step;
            pairs[k].rng_off = d0.rng_off;
            pairs[k].ngm1 = d0.ng - 1;
            if (d0.rng_off == d1.rng_off && d0.ng == d1.ng && d0.naz >= 2 && d1.naz >= 2 && d0.ng >= 2)
                pairs[k].flags |= kPairShared;
        }
        v->pairs = pairs;
        {   // r2-space descriptors: gate thresholds per range table, per-sweep / per-pair constants
            std::vector<double2> gt(n_rng + 2, make_double2(0.0, 0.0));
            std::vector<GateAux> ht(n_rng + 2);
            memset(ht.data(), 0, ht.size() * sizeof(GateAux));
            std::vector<FastSweep> fs(ne);
            memset(fs.data(), 0, fs.size() * sizeof(FastSweep));
            for (int s = 0; s < ne; ++s) {
                const SweepDesc& d = v->sweeps[s];
                FastSweep& f = fs[s];
                f.rng_off = d.rng_off;
                f.ngm1 = d.ng - 1;
                f.inv_step = (float)d.rng_inv_step;
                f.c0 = (float)(1.0 - d.r_first * d.rng_inv_step);
                if (d.naz >= 2 && d.ng >= 2) f.flags |= kFsUsable;
                if (d.ng > 0) {
                    f.tg_first = sqrt_threshold_ge(d.r_first);
                    f.tg_gt_last = sqrt_threshold_gt(d.r_last);
                }
                if (rng_owner[s] == s) {
                    double prev = d.ng > 0 ? sqrt_threshold_ge(rng[s][0]) : 0.0;
                    for (int j = 0; j < d.ng; ++j) {
                        // interval j serves r in [R[j], R[j+1]); the last usable one (j = ng-2) extends to
                        // r <= R[ng-1], the reference's clamp of the gate index (RadarGridC.pyx:291-295)
                        const double next = (j + 2 < d.ng) ? sqrt_threshold_ge(rng[s][j + 1]) : f.tg_gt_last;
                        gt[d.rng_off + j] = make_double2(prev, next);
                        GateAux& hA = ht[d.rng_off + j];
                        volatile double rs = rng[s][j] * rng[s][j];
                        hA.rsq = rs;
                        hA.r = (float)rng[s][j];
                        hA.inv_dr = (j + 1 < d.ng) ? (float)(1.0 / (rng[s][j + 1] - rng[s][j])) : 0.f;
                        prev = next;
                    }
                }
                if (s + 1 < ne) {
                    const SweepDesc& d1 = v->sweeps[s + 1];
                    if (fit_pair_r2(&f, d.elev, d1.elev)) f.flags |= kFsPolyOk;
                    if (d.rng_off == d1.rng_off && d.ng == d1.ng && d.naz >= 2 && d1.naz >= 2 && d.ng >= 2)
                        f.flags |= kFsShared;
                }
            }
            std::vector<double2> azt(n_az ? n_az : 1);
            for (int s = 0; s < ne; ++s) {
                const SweepDesc& d = v->sweeps[s];
                for (int i = 0; i < d.naz; ++i) {
                    const double gap = (i + 1 < d.naz) ? az[s][i + 1] - az[s][i] : az[s][0] + 360.0 - az[s][i];
                    azt[d.az_off + i] = make_double2(az[s][i], gap > 0.0 ? 1.0 / gap : 0.0);
                }
            }
            VOL_TRY(g_pool.alloc((void**)&v->d_azt, azt.size() * sizeof(double2)));
            VOL_TRY(cudaMemcpy(v->d_azt, azt.data(), azt.size() * sizeof(double2), cudaMemcpyHostToDevice));
            v->bytes += azt.size() * sizeof(double2);
            bool uni = ne >= 2;
            for (int s = 0; s < ne && uni; ++s) {
                const int need = (s + 1 < ne) ? (kFsPolyOk | kFsShared | kFsUsable) : kFsUsable;
                if ((fs[s].flags & need) != need || fs[s].rng_off != fs[0].rng_off || fs[s].ngm1 != fs[0].ngm1) uni = false;
            }
            // ... with constant gate spacing: R[j] = R[0] + j * step to within 1e-6 of a step
            if (uni) {
                const SweepDesc& d0 = v->sweeps[0];
                const double step = (d0.r_last - d0.r_first) / (double)(d0.ng - 1);
                for (int j = 0; j < d0.ng && uni; ++j)
                    if (fabs(rng[0][j] - (d0.r_first + j * step)) > 1e-6 * step) uni = false;
                v->uni_step = (float)step;
                v->uni_r_first = (float)d0.r_first;
            }
            v->uniform = uni;
            v->uni_rng_off = fs[0].rng_off; v->uni_ngm1 = fs[0].ngm1;
            v->uni_inv_step = fs[0].inv_step; v->uni_c0 = fs[0].c0;
            VOL_TRY(g_pool.alloc((void**)&v->d_fs, fs.size() * sizeof(FastSweep)));
            VOL_TRY(g_pool.alloc((void**)&v->d_gt, gt.size() * sizeof(double2)));
            VOL_TRY(g_pool.alloc((void**)&v->d_ht, ht.size() * sizeof(GateAux)));
            VOL_TRY(cudaMemcpyAsync(v->d_fs, fs.data(), fs.size() * sizeof(FastSweep), cudaMemcpyHostToDevice, st));
            v->fs_host = fs;
            VOL_TRY(cudaMemcpyAsync(v->d_gt, gt.data(), gt.size() * sizeof(double2), cudaMemcpyHostToDevice, st));
            VOL_TRY(cudaMemcpyAsync(v->d_ht, ht.data(), ht.size() * sizeof(GateAux), cudaMemcpyHostToDevice, st));
            VOL_TRY(cudaStreamSynchronize(st));      // the host vectors die with this block
            v->bytes += fs.size() * sizeof(FastSweep) + gt.size() * sizeof(double2) + ht.size() * sizeof(GateAux);
        }
        VOL_TRY(g_pool.alloc((void**)&v->d_pairs, pairs.size() * sizeof(PairDesc)));
        VOL_TRY(cudaMemcpyAsync(v->d_pairs, pairs.data(), pairs.size() * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
        v->bytes += pairs.size() * sizeof(PairDesc);
    }
    unsigned long long mm[kMaxFields * 3];
    VOL_TRY(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
    VOL_TRY(cudaStreamSynchronize(st));
#undef VOL_TRY
    for (int f = 0; f < nfield; ++f) {
        v->src_count[f] = mm[3 * f + 2];
        if (v->src_count[f]) {
            auto unkey = [](unsigned long long k) {
                const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
                double d;
                memcpy(&d, &b, sizeof(d));
                return d;
            };
            v->src_min[f] = unkey(mm[3 * f]);
            v->src_max[f] = unkey(mm[3 * f + 1]);
        }
    }
    *out = v;
    return PCG_OK;
}

extern "C" {

int pcg_volume_create(int ne, const int* naz, const int* ng, const double* const* az,
                      const double* const* rng, int nfield, const double* const* const* val,
                      const double* fix_elev, double fill, int value_dtype, pcg_volume** out) {
    return volume_create_impl(ne, naz, ng, az, rng, nfield, val, nullptr, nullptr, fix_elev, fill, value_dtype, out);
}

int pcg_volume_create_raw(int ne, const int* naz, const int* ng, const double* const* az,
                          const double* const* rng, int nfield, const double* const* const* val,
                          const long long* val_ld, const int* const* ray_order, const double* fix_elev,
                          double fill, int value_dtype, pcg_volume** out) {
    if (!out) return fail(PCG_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (ne < 0 || ne > kMaxSweeps) return fail(PCG_ERR_ARG, "ne must be in [0, %d]", kMaxSweeps);
    if (ne > 0 && (!naz || !ng || !az || !ray_order)) return fail(PCG_ERR_ARG, "NULL table pointer");
    // the azimuth tables in packed order (host: a few hundred values per sweep)
    std::vector<std::vector<double>> sorted(ne > 0 ? ne : 1);
    std::vector<const double*> azp(ne > 0 ? ne : 1, nullptr);
    for (int s = 0; s < ne; ++s) {
        if (naz[s] < 0 || ng[s] < 0) return fail(PCG_ERR_ARG, "negative table length in sweep %d", s);
        if (naz[s] > 0 && (!az[s] || !ray_order[s])) return fail(PCG_ERR_ARG, "NULL azimuth / ray order (sweep %d)", s);
        if (val_ld && val_ld[s] < ng[s]) return fail(PCG_ERR_ARG, "val_ld[%d] is smaller than ng[%d]", s, s);
        sorted[s].resize(naz[s]);
        for (int i = 0; i < naz[s]; ++i) {
            const int r = ray_order[s][i];
            if (r < 0 || r >= naz[s]) return fail(PCG_ERR_ARG, "ray_order[%d][%d] = %d out of range", s, i, r);
            sorted[s][i] = az[s][r];
        }
        azp[s] = sorted[s].data();
    }
    return volume_create_impl(ne, naz, ng, azp.data(), rng, nfield, val, val_ld, ray_order, fix_elev, fill,
                              value_dtype, out);
}

int pcg_volume_destroy(pcg_volume* v) {
    if (!v) return PCG_OK;
    // the arenas belong to v->device, whatever device the calling thread happens to be on
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != v->device) cudaSetDevice(v->device);
    cudaDeviceSynchronize();          // nothing may still be reading the arenas when they go back to the pool
    if (v->device >= 0 && v->device < kMaxDevices) {
        // cached network tables hold this volume's arena pointers: drop them with it
        DeviceCtx& c = g_ctx[v->device];
        std::lock_guard<std::mutex> lock(c.mutex);
        c.net_key.clear();
        c.net_ptr = nullptr;
    }
    g_pool.release(v->d_az);
    g_pool.release(v->d_rng);
    g_pool.release(v->d_sweeps);
    g_pool.release(v->d_pairs);
    g_pool.release(v->d_fs);
    g_pool.release(v->d_azt);
    g_pool.release(v->d_ls);
    g_pool.release(v->d_gt);
    g_pool.release(v->d_ht);
    for (int f = 0; f < kMaxFields; ++f) g_pool.release(v->d_val[f]);
    if (prev >= 0 && prev != v->device) cudaSetDevice(prev);
    delete v;
    return PCG_OK;
}

int pcg_volume_info(const pcg_volume* v, int* ne, int* nfield, int* value_dtype, size_t* device_bytes) {
    if (!v) return fail(PCG_ERR_ARG, "volume handle is NULL");
    if (ne) *ne = v->ne;
    if (nfield) *nfield = v->nfield;
    if (value_dtype) *value_dtype = v->dtype;
    if (device_bytes) *device_bytes = v->bytes;
    return PCG_OK;
}

int pcg_get_cappi3d_dev(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                        const double* d_grid_x, const double* d_grid_y, int nx, int ny,
                        const double* level_heights, int nz, double fill, int blind_code,
                        double beam_width_deg, double radar_x, double radar_y, int merge_max,
                        double* d_out, int32_t* d_index_out, void* stream) {
    int rc = check_volume(vol);
    if (rc) return rc;
    rc = check_common(effective_earth_radius, nx, ny);
    if (rc) return rc;
    const int shift = (radar_x != 0.0 || radar_y != 0.0 || merge_max) ? 1 : 0;
    return enqueue_cappi(vol, radar_height, effective_earth_radius, d_grid_x, d_grid_y,
                         (long long)nx * ny, ny, level_heights, nz, fill, blind_code, beam_width_deg,
                         radar_x, radar_y, shift, merge_max, d_out, d_index_out, (cudaStream_t)stream);
}

int pcg_get_cr_dev(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                   const double* d_grid_x, const double* d_grid_y, int nx, int ny, double fill,
                   double radar_x, double radar_y, double* d_out, void* stream) {
    int rc = check_volume(vol);
    if (rc) return rc;
    rc = check_common(effective_earth_radius, nx, ny);
    if (rc) return rc;
    const int shift = (radar_x != 0.0 || radar_y != 0.0) ? 1 : 0;
    return enqueue_cr(vol, radar_height, effective_earth_radius, d_grid_x, d_grid_y, (long long)nx * ny,
                      fill, -1, 0.0, 0, radar_x, radar_y, shift, d_out, nullptr, (cudaStream_t)stream);
}

int pcg_get_cappi3d(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                    const double* grid_x, const double* grid_y, int nx, int ny,
                    const double* level_heights, int nz, double fill, int blind_code,
                    double beam_width_deg, double* out, int32_t* index_out) {
    return pcg_get_cappi3d_ex(vol, radar_height, effective_earth_radius, grid_x, grid_y, nx, ny, level_heights, nz,
                              fill, blind_code, beam_width_deg, fill, out, index_out);
}

int pcg_get_cappi3d_ex(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                       const double* grid_x, const double* grid_y, int nx, int ny,
                       const double* level_heights, int nz, double fill, int blind_code,
                       double beam_width_deg, double out_missing, double* out, int32_t* index_out) {
    int rc = check_volume(vol);
    if (rc) return rc;
    rc = check_common(effective_earth_radius, nx, ny);
    if (rc) return rc;
    const size_t ncol = (size_t)nx * (size_t)ny;
    const size_t nout = (size_t)vol->nfield * (size_t)nz * ncol;
    if (ncol == 0 || nz <= 0) return nz < 0 ? fail(PCG_ERR_ARG, "nz must be non-negative") : PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(nout * sizeof(double)))) return rc;
    int32_t* d_idx = nullptr;
    if (index_out) {
        if ((rc = g_scratch[S_IDX].ensure((size_t)nz * ncol * kIdxStride * sizeof(int32_t)))) return rc;
        d_idx = (int32_t*)g_scratch[S_IDX].ptr;
    }
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    if ((rc = g_scratch[S_TMP2].ensure(sizeof(int)))) return rc;
    CUDA_TRY(cudaMemsetAsync(g_scratch[S_TMP2].ptr, 0, sizeof(int), st));
    g_nonfinite_flag = (int*)g_scratch[S_TMP2].ptr;
    rc = enqueue_cappi(vol, radar_height, effective_earth_radius, (const double*)g_scratch[S_GX].ptr,
                       (const double*)g_scratch[S_GY].ptr, (long long)ncol, ny, level_heights, nz, fill,
                       blind_code, beam_width_deg, 0.0, 0.0, 0, 0, (double*)g_scratch[S_OUT].ptr, d_idx, st);
    g_nonfinite_flag = nullptr;
    if (rc) return rc;
    if ((rc = relabel_missing((double*)g_scratch[S_OUT].ptr, nout, fill, out_missing, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, g_scratch[S_OUT].ptr, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (index_out)
        CUDA_TRY(cudaMemcpyAsync(index_out, d_idx, (size_t)nz * ncol * kIdxStride * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, st));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, g_scratch[S_TMP2].ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return flag ? PCG_WARN_NONFINITE : PCG_OK;
}

int pcg_get_cr(const pcg_volume* vol, double radar_height, double effective_earth_radius,
               const double* grid_x, const double* grid_y, int nx, int ny, double fill, double* out) {
    return pcg_get_cr_ex(vol, radar_height, effective_earth_radius, grid_x, grid_y, nx, ny, fill, fill, out);
}

int pcg_get_cr_ex(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                  const double* grid_x, const double* grid_y, int nx, int ny, double fill, double out_missing,
                  double* out) {
    int rc = check_volume(vol);
    if (rc) return rc;
    rc = check_common(effective_earth_radius, nx, ny);
    if (rc) return rc;
    const size_t ncol = (size_t)nx * (size_t)ny;
    if (ncol == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(ncol * sizeof(double)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    if ((rc = g_scratch[S_TMP2].ensure(sizeof(int)))) return rc;
    CUDA_TRY(cudaMemsetAsync(g_scratch[S_TMP2].ptr, 0, sizeof(int), st));
    g_nonfinite_flag = (int*)g_scratch[S_TMP2].ptr;
    rc = enqueue_cr(vol, radar_height, effective_earth_radius, (const double*)g_scratch[S_GX].ptr,
                    (const double*)g_scratch[S_GY].ptr, (long long)ncol, fill, -1, 0.0, 0, 0.0, 0.0, 0,
                    (double*)g_scratch[S_OUT].ptr, nullptr, st);
    g_nonfinite_flag = nullptr;
    if (rc) return rc;
    if ((rc = relabel_missing((double*)g_scratch[S_OUT].ptr, ncol, fill, out_missing, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, g_scratch[S_OUT].ptr, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, g_scratch[S_TMP2].ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return flag ? PCG_WARN_NONFINITE : PCG_OK;
}

int pcg_ppi_to_grid(const pcg_volume* vol, int sweep, double elevation_deg, double radar_height,
                    double effective_earth_radius, const double* grid_x, const double* grid_y,
                    int nx, int ny, double fill, int blind_code, double* out, int32_t* index_out) {
    int rc = check_volume(vol);
    if (rc) return rc;
    rc = check_common(effective_earth_radius, nx, ny);
    if (rc) return rc;
    if (sweep < 0 || sweep >= vol->ne) return fail(PCG_ERR_ARG, "sweep index %d out of range", sweep);
    if (blind_code < 0 || blind_code > 4)
        return fail(PCG_ERR_ARG,
                    "blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', 'hybrid', or "
                    "'nearest_valid'.");
    const size_t ncol = (size_t)nx * (size_t)ny;
    if (ncol == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(ncol * sizeof(double)))) return rc;
    int32_t* d_idx = nullptr;
    if (index_out) {
        if ((rc = g_scratch[S_IDX].ensure(ncol * kIdxStride * sizeof(int32_t)))) return rc;
        d_idx = (int32_t*)g_scratch[S_IDX].ptr;
    }
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = enqueue_cr(vol, radar_height, effective_earth_radius, (const double*)g_scratch[S_GX].ptr,
                    (const double*)g_scratch[S_GY].ptr, (long long)ncol, fill, sweep, elevation_deg,
                    blind_code, 0.0, 0.0, 0, (double*)g_scratch[S_OUT].ptr, d_idx, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, g_scratch[S_OUT].ptr, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (index_out)
        CUDA_TRY(cudaMemcpyAsync(index_out, d_idx, ncol * kIdxStride * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

int pcg_get_mosaic_cappi3d(int nradar, const pcg_volume* const* vols, const double* radar_x,
                           const double* radar_y, const double* radar_height,
                           const double* effective_earth_radius, const double* grid_x,
                           const double* grid_y, int nx, int ny, const double* level_heights,
                           int nz, double fill, double* out) {
    if (nradar < 0) return fail(PCG_ERR_ARG, "nradar must be non-negative");
    if (nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    const size_t ncol = (size_t)nx * (size_t)ny;
    const size_t nout = (size_t)nz * ncol;
    if (nout == 0) return PCG_OK;
    int rc;
    for (int i = 0; i < nradar; ++i) {
        if ((rc = check_volume(vols[i]))) return rc;
        if (!(effective_earth_radius[i] > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(nout * sizeof(double)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    // GridValue = full(fill) (RadarGridC.pyx:548): merging the first radar into an all-fill grid
    // is a plain store, so radar 0 writes and the others merge (RadarGridC.pyx:586-592).
    if (nradar == 0) {
        for (size_t i = 0; i < nout; ++i) out[i] = fill;
        return PCG_OK;
    }
    for (int i = 0; i < nradar; ++i) {
        // get_mosaic_CAPPI_3d uses field 0 of each radar, blind "mask", default beam width
        rc = enqueue_cappi(vols[i], radar_height[i], effective_earth_radius[i], (const double*)g_scratch[S_GX].ptr,
                           (const double*)g_scratch[S_GY].ptr, (long long)ncol, ny, level_heights, nz, fill, 0,
                           1.0, radar_x[i], radar_y[i], 1, i > 0 ? 1 : 0, (double*)g_scratch[S_OUT].ptr, nullptr,
                           st, 1);
        if (rc) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(out, g_scratch[S_OUT].ptr, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

int pcg_debug_cartesian_to_antenna(const double* x, const double* y, size_t n, double z, double h,
                                   double effective_earth_radius, double* az, double* r, double* el) {
    if (!(effective_earth_radius > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    if (n == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    int rc;
    if ((rc = lib_stream(&st))) return rc;
    Scratch* s = g_scratch;
    const size_t b = n * sizeof(double);
    if ((rc = s[S_GX].ensure(b)) || (rc = s[S_GY].ensure(b)) || (rc = s[S_TMP0].ensure(b)) ||
        (rc = s[S_TMP1].ensure(b)) || (rc = s[S_TMP2].ensure(b)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(s[S_GX].ptr, x, b, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s[S_GY].ptr, y, b, cudaMemcpyHostToDevice, st));
    volatile double av = effective_earth_radius + h;
    volatile double a2 = av * av;
    volatile double twoa = 2.0 * av;
    LevelConst lc;
    fill_level(&lc, z, effective_earth_radius, a2, twoa);
    debug_c2a_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        (const double*)s[S_GX].ptr, (const double*)s[S_GY].ptr, n, lc, a2, twoa, effective_earth_radius,
        (double*)s[S_TMP0].ptr, (double*)s[S_TMP1].ptr, (double*)s[S_TMP2].ptr);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(az, s[S_TMP0].ptr, b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(r, s[S_TMP1].ptr, b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(el, s[S_TMP2].ptr, b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

int pcg_debug_sqrt_check(const double* a, size_t n, uint64_t* mismatches, double* worst_rinv_err) {
    if (!a || !mismatches || !worst_rinv_err) return fail(PCG_ERR_ARG, "NULL pointer");
    *mismatches = 0;
    *worst_rinv_err = 0.0;
    if (n == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    int rc;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(n * sizeof(double))) || (rc = g_scratch[S_TMP0].ensure(16))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, a, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(g_scratch[S_TMP0].ptr, 0, 16, st));
    debug_sqrt_kernel<<<148 * 8, 256, 0, st>>>((const double*)g_scratch[S_GX].ptr, n,
                                               (unsigned long long*)g_scratch[S_TMP0].ptr,
                                               (double*)g_scratch[S_TMP0].ptr + 1);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    unsigned long long host[2];
    CUDA_TRY(cudaMemcpyAsync(host, g_scratch[S_TMP0].ptr, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *mismatches = host[0];
    memcpy(worst_rinv_err, &host[1], sizeof(double));
    return PCG_OK;
}

int pcg_debug_xye_to_antenna(const double* x, const double* y, size_t n, double elevation_deg, double h,
                             double effective_earth_radius, double* az, double* r) {
    if (!(effective_earth_radius > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    if (n == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    int rc;
    if ((rc = lib_stream(&st))) return rc;
    Scratch* s = g_scratch;
    const size_t b = n * sizeof(double);
    if ((rc = s[S_GX].ensure(b)) || (rc = s[S_GY].ensure(b)) || (rc = s[S_TMP0].ensure(b)) ||
        (rc = s[S_TMP1].ensure(b)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(s[S_GX].ptr, x, b, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s[S_GY].ptr, y, b, cudaMemcpyHostToDevice, st));
    volatile double av = effective_earth_radius + h;
    volatile double a2 = av * av;
    volatile double twoa = 2.0 * av;
    debug_xye_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        (const double*)s[S_GX].ptr, (const double*)s[S_GY].ptr, n, tan(elevation_deg * kDeg2Rad), av, a2,
        twoa, effective_earth_radius, (double*)s[S_TMP0].ptr, (double*)s[S_TMP1].ptr);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(az, s[S_TMP0].ptr, b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(r, s[S_TMP1].ptr, b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

}  // extern "C"

// ---- network compositing (K_net) ------------------------------------------------------------

namespace {


__global__ void fill_f64_kernel(double* p, size_t n, double v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int M, bool U>
int launch_radar_pass(const NetArgs& a, const NetAcc& acc, int ir, int ntiles, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(net_radar_kernel<M, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    net_radar_kernel<M, U><<<ntiles, kNetThreads, smem, st>>>(a, acc, ir);
    return PCG_OK;
}

int launch_radar_pass_any(const NetArgs& a, const NetAcc& acc, int method, bool uni, int ir, int ntiles, size_t smem,
                          cudaStream_t st) {
    switch (method * 2 + (uni ? 1 : 0)) {
        case kNetMax * 2: return launch_radar_pass<kNetMax, false>(a, acc, ir, ntiles, smem, st);
        case kNetMax * 2 + 1: return launch_radar_pass<kNetMax, true>(a, acc, ir, ntiles, smem, st);
        case kNetNearest * 2: return launch_radar_pass<kNetNearest, false>(a, acc, ir, ntiles, smem, st);
        case kNetNearest * 2 + 1: return launch_radar_pass<kNetNearest, true>(a, acc, ir, ntiles, smem, st);
        case kNetExp * 2: return launch_radar_pass<kNetExp, false>(a, acc, ir, ntiles, smem, st);
        case kNetExp * 2 + 1: return launch_radar_pass<kNetExp, true>(a, acc, ir, ntiles, smem, st);
        case kNetQuality * 2: return launch_radar_pass<kNetQuality, false>(a, acc, ir, ntiles, smem, st);
        case kNetQuality * 2 + 1: return launch_radar_pass<kNetQuality, true>(a, acc, ir, ntiles, smem, st);
        default: return fail(PCG_ERR_ARG, "Unsupported composite_method code %d", method);
    }
}

template <int M, bool U>
int launch_tile(const NetArgs& a, unsigned ntile, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(net_tile_kernel<M, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    net_tile_kernel<M, U><<<ntile, kNet2Threads, smem, st>>>(a);
    return PCG_OK;
}

int launch_tile_any(const NetArgs& a, int method, bool uni, unsigned ntile, size_t smem, cudaStream_t st) {
    switch (method * 2 + (uni ? 1 : 0)) {
        case kNetMax * 2: return launch_tile<kNetMax, false>(a, ntile, smem, st);
        case kNetMax * 2 + 1: return launch_tile<kNetMax, true>(a, ntile, smem, st);
        case kNetNearest * 2: return launch_tile<kNetNearest, false>(a, ntile, smem, st);
        case kNetNearest * 2 + 1: return launch_tile<kNetNearest, true>(a, ntile, smem, st);
        case kNetExp * 2: return launch_tile<kNetExp, false>(a, ntile, smem, st);
        case kNetExp * 2 + 1: return launch_tile<kNetExp, true>(a, ntile, smem, st);
        case kNetQuality * 2: return launch_tile<kNetQuality, false>(a, ntile, smem, st);
        case kNetQuality * 2 + 1: return launch_tile<kNetQuality, true>(a, ntile, smem, st);
        default: return fail(PCG_ERR_ARG, "Unsupported composite_method code %d", method);
    }
}

// PCG_NET=passes keeps the round-1 radar-major passes (A/B runs)
bool use_tile_kernel() {
    static const int choice = [] {
        const char* e = getenv("PCG_NET");
        return (e && strcmp(e, "passes") == 0) ? 0 : 1;
    }();
    return choice != 0;
}

int launch_finalize(const NetArgs& a, const NetAcc& acc, int method, cudaStream_t st) {
    const size_t nvox = (size_t)a.nz * (size_t)a.ncol;
    const unsigned blocks = (unsigned)((nvox + 255) / 256 < 148 * 32 ? (nvox + 255) / 256 : 148 * 32);
    switch (method) {
        case kNetMax: net_finalize_kernel<kNetMax><<<blocks, 256, 0, st>>>(a, acc); break;
        case kNetNearest: net_finalize_kernel<kNetNearest><<<blocks, 256, 0, st>>>(a, acc); break;
        case kNetExp: net_finalize_kernel<kNetExp><<<blocks, 256, 0, st>>>(a, acc); break;
        default: net_finalize_kernel<kNetQuality><<<blocks, 256, 0, st>>>(a, acc); break;
    }
    return PCG_OK;
}

int net_streams(NetStreams** out) {
    NetStreams& n = cur_ctx().net_streams;
    if (!n.ready) {
        for (int k = 0; k < kNetStreams; ++k) {
            CUDA_TRY(cudaStreamCreateWithFlags(&n.side[k], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&n.ev_side[k], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&n.ev_class, cudaEventDisableTiming));
        n.ready = true;
    }
    *out = &n;
    return PCG_OK;
}

// d_* pointers are DEVICE pointers; tables are built on the host and uploaded through S_NET.  The blob of
// the last call is kept: repeated calls with the same radars / geometry / options only launch.

template <typename T>
void key_put(std::vector<char>* key, const T* p, size_t n) {
    const char* b = reinterpret_cast<const char*>(p);
    key->insert(key->end(), b, b + n * sizeof(T));
}


int enqueue_network(int nradar, const pcg_volume* const* vols, const double* radar_x, const double* radar_y,
                    const double* radar_height, const double* re, const double* const* beam_widths,
                    const pcg_network_params* prm, const double* d_gx, const double* d_gy, long long ncol, int ny,
                    const double* levels, int nz, double fill, double* d_comp, int16_t* d_valid,
                    int16_t* d_cov, int8_t* d_blind, double* d_cr, double* d_quality,
                    unsigned long long* d_ties, cudaStream_t st, double out_missing, bool* relabelled) {
    if (relabelled) *relabelled = false;
    if (!prm) return fail(PCG_ERR_ARG, "params is NULL");
    if (nradar < 1) return fail(PCG_ERR_ARG, "radar_volumes must contain at least one volume.");
    if (nradar > 255) return fail(PCG_ERR_ARG, "at most 255 radars per composite");
    if (prm->method < 0 || prm->method > 3) return fail(PCG_ERR_ARG, "Unsupported composite_method code %d", prm->method);
    if (!(prm->influence_radius_m > 0.0)) return fail(PCG_ERR_ARG, "influence_radius_m must be positive.");
    if (prm->blind_code < 0 || prm->blind_code > 4)
        return fail(PCG_ERR_ARG,
                    "blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', 'hybrid', or "
                    "'nearest_valid'.");
    if (nz < 0) return fail(PCG_ERR_ARG, "nz must be non-negative");
    if (ncol == 0 || (nz == 0 && !d_cr)) return PCG_OK;       // (no levels: the CR plane is still wanted)
    int rc;
    if ((rc = scratch_begin(st))) return rc;                  // a previous call on another stream may still use the scratch
    bool uni = true;
    size_t npair_total = 0, n_ls = 0, nsweep_total = 0;
    for (int i = 0; i < nradar; ++i) {
        if ((rc = check_volume(vols[i]))) return rc;
        if (!(re[i] > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
        if (vols[i]->dtype != PCG_VALUE_F32)
            return fail(PCG_ERR_ARG, "radar %d: network compositing needs a packed volume (>= 2 sweeps, strictly "
                                     "increasing tables, value_dtype f32)", i);
        if (vols[i]->ne > kNetMaxSweeps) return fail(PCG_ERR_ARG, "radar %d: at most %d sweeps", i, kNetMaxSweeps);
        if (!(fill == vols[i]->fill))
            return fail(PCG_ERR_ARG, "fillvalue %.17g differs from the value volume %d was packed with (%.17g)", fill, i,
                        vols[i]->fill);
        uni = uni && vols[i]->uniform;
        npair_total += (size_t)(vols[i]->ne - 1);
        nsweep_total += (size_t)vols[i]->ne;
        n_ls += (size_t)nz * (2 * vols[i]->ne + 3);
    }
    const size_t bytes_radars = sizeof(NetRadar) * (size_t)nradar;
    const size_t bytes_pairs = sizeof(NetPair) * npair_total;
    const size_t bytes_levels = sizeof(NetLevel) * (size_t)nradar * nz;
    const size_t bytes_spairs = sizeof(PairDesc) * npair_total;
    const size_t bytes_ls = sizeof(LevelSweep) * n_ls, bytes_lp = 0;
    const size_t off_pairs = bytes_radars, off_levels = off_pairs + bytes_pairs, off_spairs = off_levels + bytes_levels;
    const size_t off_ls = (off_spairs + bytes_spairs + 31) & ~(size_t)31, off_lp = off_ls + bytes_ls;
    const size_t off_pq = (off_lp + bytes_lp + 31) & ~(size_t)31;
    const size_t total = off_pq + sizeof(NetPairQ) * nsweep_total;
    if ((rc = g_scratch[S_NET].ensure(total + 64))) return rc;
    char* d_base = (char*)g_scratch[S_NET].ptr;

    // reach of each radar on the ground: slant range r >= 0.999 * s * min(a, g) / Re for phi < 0.15, so no
    // gate is reachable beyond reach[i] (HUGE_VAL: no culling for absurd heights or ranges)
    std::vector<double> reach(nradar);
    {
        double zmin = nz > 0 ? levels[0] : 0.0;
        for (int k = 1; k < nz; ++k) zmin = levels[k] < zmin ? levels[k] : zmin;
        for (int i = 0; i < nradar; ++i) {
            double rlast = 0.0;
            for (int k = 0; k < vols[i]->ne; ++k) rlast = vols[i]->sweeps[k].r_last > rlast ? vols[i]->sweeps[k].r_last : rlast;
            const double min_ag = re[i] + (radar_height[i] < zmin ? radar_height[i] : zmin);
            const double factor = 0.999 * min_ag / re[i];
            reach[i] = (factor < 0.9 || !(rlast < 1.0e6)) ? HUGE_VAL : rlast / factor * 1.001 + 1.0;
        }
    }
    // everything the tables depend on
    std::vector<char> key;
    key_put(&key, &nradar, 1); key_put(&key, &nz, 1); key_put(&key, prm, 1); key_put(&key, &fill, 1);
    for (int i = 0; i < nradar; ++i) key_put(&key, &vols[i]->uid, 1);
    key_put(&key, radar_x, nradar); key_put(&key, radar_y, nradar);
    key_put(&key, radar_height, nradar); key_put(&key, re, nradar); key_put(&key, levels, nz);
    for (int i = 0; i < nradar; ++i) {
        const double* bw = beam_widths ? beam_widths[i] : nullptr;
        const int has = bw ? 1 : 0;
        key_put(&key, &has, 1);
        if (bw) key_put(&key, bw, vols[i]->ne);
        key_put(&key, &vols[i]->src_min[0], 1); key_put(&key, &vols[i]->src_max[0], 1);
    }
    const bool cached = (g_net_ptr == (void*)d_base) && key.size() == g_net_key.size() &&
                        memcmp(key.data(), g_net_key.data(), key.size()) == 0;
    if (!cached) {
        std::vector<char> blob(total, 0);
        NetRadar* radars = reinterpret_cast<NetRadar*>(blob.data());
        NetPair* npairs = reinterpret_cast<NetPair*>(blob.data() + off_pairs);
        NetLevel* nlevels = reinterpret_cast<NetLevel*>(blob.data() + off_levels);
        PairDesc* spairs = reinterpret_cast<PairDesc*>(blob.data() + off_spairs);
        LevelSweep* ls_all = reinterpret_cast<LevelSweep*>(blob.data() + off_ls);
        NetPairQ* pq_all = reinterpret_cast<NetPairQ*>(blob.data() + off_pq);
        size_t poff = 0, lsoff = 0, qoff = 0;
        std::vector<unsigned char> slow;
        for (int i = 0; i < nradar; ++i) {
            const pcg_volume* v = vols[i];
            const int ne = v->ne;
            NetRadar& R = radars[i];
            R.vol = view_of(v);
            R.pairs = reinterpret_cast<const PairDesc*>(d_base + off_spairs) + poff;
            R.npairs = reinterpret_cast<const NetPair*>(d_base + off_pairs) + poff;
            R.levels = reinterpret_cast<const NetLevel*>(d_base + off_levels) + (size_t)i * nz;
            R.x = radar_x[i];
            R.y = radar_y[i];
            R.Re = re[i];
            volatile double av = re[i] + radar_height[i];   // RadarGridC.pyx:144
            volatile double a2 = av * av;
            volatile double twoa = 2.0 * av;
            R.a = av; R.a2 = a2; R.twoa = twoa;
            R.guard = kGuardDelta * twoa;
            R.inv_twoa = 1.0 / twoa;
            R.a2_np = pow((double)av, 2.0);                 // radar_radius ** 2 (transforms.py:213)
            R.ne = ne;
            // r2-space tables
            R.r2.lp = reinterpret_cast<const LevelPair*>(d_base + off_ls) + lsoff;
            R.r2.fs = v->d_fs; R.r2.g_t = v->d_gt; R.r2.h_t = v->d_ht; R.r2.az_t = v->d_azt;
            R.r2.inv2a = (float)(1.0 / twoa);
            R.r2.u_rng_off = v->uni_rng_off; R.r2.u_ngm1 = v->uni_ngm1; R.r2.u_inv_step = v->uni_inv_step;
            R.r2.u_c0 = v->uni_c0; R.r2.u_step = v->uni_step; R.r2.u_r_first = v->uni_r_first;
            build_level_sweeps(v, re[i], av, levels, nz, ls_all + lsoff + (ne + 1), (size_t)2 * ne + 3, &slow);
            build_level_pairs(ls_all + lsoff, nz, ne, slow);
            lsoff += (size_t)nz * (2 * ne + 3);
            // range-sanity window (RadarInterp.py:1003-1011)
            if (v->src_count[0] == 0) {
                R.src_lo = 1.0; R.src_hi = 0.0;             // no valid source value: everything -> fill
            } else {
                const double mn = v->src_min[0], mx = v->src_max[0];
                double m = fabs(mn) > fabs(mx) ? fabs(mn) : fabs(mx);
                if (m < 1.0) m = 1.0;
                double tol = 1e-6 * m;
                if (tol < 1e-6) tol = 1e-6;
                R.src_lo = mn - tol;
                R.src_hi = mx + tol;
            }
            if (R.src_lo <= R.src_hi) {
                float lo = (float)R.src_lo, hi = (float)R.src_hi;
                if ((double)lo < R.src_lo) lo = nextafterf(lo, HUGE_VALF);
                if ((double)hi > R.src_hi) hi = nextafterf(hi, -HUGE_VALF);
                R.src_lo_f = lo; R.src_hi_f = hi;
            } else {
                R.src_lo_f = HUGE_VALF; R.src_hi_f = -HUGE_VALF;
            }
            R.pq = reinterpret_cast<const NetPairQ*>(d_base + off_pq) + qoff;
            R.inv_decay = (prm->quality_terms & kTermRange) ? (float)(1.0 / prm->range_decay_m) : 0.f;
            R.inv_r2inf4 = (float)(4.0 / (prm->influence_radius_m * prm->influence_radius_m));
            R.cull2 = reach[i] < HUGE_VAL ? reach[i] * reach[i] : HUGE_VAL;
            const double* bw = beam_widths ? beam_widths[i] : nullptr;
            for (int k = 0; k + 1 < ne; ++k) {
                NetPair& np = npairs[poff + k];
                const SweepDesc& d0 = v->sweeps[k];
                const SweepDesc& d1 = v->sweeps[k + 1];
                np.obs_first = d0.r_first > d1.r_first ? d0.r_first : d1.r_first;
                np.obs_last = d0.r_last < d1.r_last ? d0.r_last : d1.r_last;
                const double f2 = np.obs_first * np.obs_first, l2 = np.obs_last * np.obs_last;
                np.of2_lo = f2 * (1.0 - 1e-9); np.of2_hi = f2 * (1.0 + 1e-9);
                np.ol2_lo = l2 * (1.0 - 1e-9); np.ol2_hi = l2 * (1.0 + 1e-9);
                const double b0 = bw ? bw[k] : 1.0, b1 = bw ? bw[k + 1] : 1.0;
                const double rep = 0.5 * (b0 + b1);
                np.beam_k = (prm->quality_terms & kTermBeam)
                                ? (float)(tan(rep * kDeg2Rad * 0.5) / prm->beam_radius_ref_m) : 0.f;
                const double gap_half = 0.5 * (d1.elev - d0.elev);
                const double gq = gap_half / prm->vertical_gap_ref_deg;
                np.vert_w = (prm->quality_terms & kTermVertical) ? (float)(1.0 / (1.0 + gq * gq)) : 1.f;
                PairDesc sp = v->pairs[k];          // what stage_pairs does for the single-radar kernel
                volatile double s_lo = sp.c_lo * (double)twoa, s_hi = sp.c_hi * (double)twoa, s_mid = sp.c_mid * (double)twoa;
                sp.c_lo = s_lo; sp.c_hi = s_hi; sp.c_mid = s_mid;
                spairs[poff + k] = sp;
            }
            for (int k = 0; k < ne; ++k) {          // (entry ne - 1 is only ever loaded speculatively)
                NetPairQ& q = pq_all[qoff + k];
                const FastSweep& f = v->fs_host[k];
                q.q1 = f.q1; q.q2 = f.q2; q.q3 = f.q3; q.q4 = f.q4; q.q5 = f.q5;
                q.beam_k = k + 1 < ne ? npairs[poff + k].beam_k : 0.f;
                q.vert_w = k + 1 < ne ? npairs[poff + k].vert_w : 1.f;
                q.pad = 0.f;
            }
            qoff += (size_t)ne;
            poff += (size_t)(ne - 1);
            for (int k = 0; k < nz; ++k) {
                NetLevel& L = nlevels[(size_t)i * nz + k];
                fill_level(&L.cy, levels[k], re[i], a2, twoa);
                const double g = re[i] + levels[k];
                L.g2_np = pow(g, 2.0);
                L.a2g2_np = R.a2_np + L.g2_np;
                volatile double t = 2.0 * (double)av;
                volatile double tg = t * g;                 // 2 * radar_radius * gate_radius (left to right)
                L.twoag_np = tg;
            }
        }
        // the previous launch may still be reading the old tables
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(d_base, blob.data(), total, cudaMemcpyHostToDevice));
        g_net_key.swap(key);
        g_net_ptr = (void*)d_base;
    }

    NetArgs A;
    memset(&A, 0, sizeof(A));
    A.radars = (const NetRadar*)d_base;
    A.nradar = nradar;
    A.nz = nz;
    A.gx = d_gx; A.gy = d_gy; A.ncol = ncol;
    A.ny = ny; A.nx = (int)(ncol / (ny > 0 ? ny : 1));
    A.py_log2 = (A.nx >= 4 && ny >= 8) ? 3 : 5;
    const int PY = 1 << A.py_log2, PX = 32 / PY, nwarp = kNetThreads / 32;
    A.tiles_y = (int)((ny + (long long)nwarp * PY - 1) / ((long long)nwarp * PY));
    const long long tiles_x = (A.nx + PX - 1) / PX;
    const long long ntile = tiles_x * A.tiles_y;
    if (ntile > 0x7fffffffll) return fail(PCG_ERR_ARG, "grid too large for one network call");
    A.fill = fill; A.fill32 = (float)fill;
    A.quality = 1;
    A.sanity = prm->sanity_filter ? 1 : 0;
    A.nearest_gate = (prm->blind_code == 1 || prm->blind_code == 3 || prm->blind_code == 4);
    A.lowest_sweep = (prm->blind_code == 2 || prm->blind_code == 3 || prm->blind_code == 4);
    A.highest_sweep = (prm->blind_code == 4);
    A.composite = d_comp; A.valid_count = d_valid; A.coverage = d_cov; A.blind = d_blind; A.cr = d_cr;
    A.npeer = g_net_peers.n; A.peer_plane = g_net_peers.plane; A.peer_off = g_net_peers.off;
    for (int q = 0; q < g_net_peers.n; ++q) A.peer[q] = g_net_peers.ptr[q];
    A.quality_out = d_quality; A.tie_count = d_ties;
    A.out_missing = out_missing;
    A.ne_max = 0;
    for (int i = 0; i < nradar; ++i) A.ne_max = vols[i]->ne > A.ne_max ? vols[i]->ne : A.ne_max;

    // (the per-radar quality diagnostics are written by the radar-major passes only)
    if (use_tile_kernel() && !d_quality) {
        // ---- one tile-persistent launch (pcg_net2.cuh): no lists, no accumulators in HBM, no finalize ----
        const bool nearest = prm->method == kNetNearest;
        int chunk = nz > 0 ? nz : 1;
        // levels per launch: the accumulators of a block live in shared memory, and how many blocks fit an SM
        // decides how much latency the kernel hides (PCG_NET_SMEM_KB: tuning runs)
        static const size_t smem_budget = [] {
            const char* e = getenv("PCG_NET_SMEM_KB");
            const long kb = e ? atol(e) : 0;
            return (size_t)((kb >= 8 && kb <= 200) ? kb : kNet2SmemKB) * 1024;
        }();
        while (chunk > 1 && net2_smem_bytes(A.ne_max, chunk, nearest) > smem_budget) --chunk;
        for (int z0 = 0; z0 < (nz > 0 ? nz : 1); z0 += chunk) {
            NetArgs B = A;
            B.z0 = z0;
            B.nz = nz - z0 < chunk ? nz - z0 : chunk;
            const size_t off = (size_t)z0 * (size_t)ncol;
            if (d_comp) B.composite = d_comp + off;
            if (d_valid) B.valid_count = d_valid + off;
            if (d_cov) B.coverage = d_cov + off;
            if (d_blind) B.blind = d_blind + off;
            if (z0 > 0) B.cr = nullptr;
            if (B.npeer) for (int q = 0; q < B.npeer; ++q) B.peer[q] = A.peer[q] + (size_t)z0 * (size_t)A.peer_plane;
            const size_t smem = net2_smem_bytes(A.ne_max, B.nz, nearest);
            if ((rc = launch_tile_any(B, prm->method, uni, (unsigned)ntile, smem, st))) return rc;
            g_launches.fetch_add(1);
            CUDA_TRY(cudaGetLastError());
        }
        if (relabelled) *relabelled = true;
        return scratch_done(st);
    }
    if (nz == 0) {                                           // radar passes need at least one level: CR alone
        NetArgs B = A;
        B.nz = 0; B.z0 = 0;
        if ((rc = launch_tile_any(B, prm->method, uni, (unsigned)ntile, net2_smem_bytes(A.ne_max, 0, false), st))) return rc;
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        return scratch_done(st);
    }

    // ---- round-1 path (PCG_NET=passes): accumulators (one 16-byte cell per voxel) and per-radar tile lists
    const size_t nvox = (size_t)nz * (size_t)ncol;
    NetAcc acc;
    memset(&acc, 0, sizeof(acc));
    if ((rc = g_scratch[S_ACC0].ensure(nvox * sizeof(NetCell))) ||
        (rc = g_scratch[S_ACC5].ensure((size_t)ncol * sizeof(float))) ||
        (rc = g_scratch[S_ACC6].ensure(((size_t)nradar * (size_t)ntile + (size_t)nradar) * sizeof(int))))
        return rc;
    acc.cell = (NetCell*)g_scratch[S_ACC0].ptr;
    acc.cr = (float*)g_scratch[S_ACC5].ptr;
    acc.tile_count = (int*)g_scratch[S_ACC6].ptr;
    acc.tiles = acc.tile_count + nradar;
    acc.ntile = (int)ntile;
    if (prm->method == kNetNearest) {
        if ((rc = g_scratch[S_ACC4].ensure(nvox * sizeof(double)))) return rc;
        acc.best_d2 = (double*)g_scratch[S_ACC4].ptr;
        fill_f64_kernel<<<148 * 8, 256, 0, st>>>(acc.best_d2, nvox, HUGE_VAL);
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaMemsetAsync(acc.cell, 0, nvox * sizeof(NetCell), st));
    CUDA_TRY(cudaMemsetAsync(acc.cr, 0xff, (size_t)ncol * sizeof(float), st));      // 0xffffffff is a NaN: no echo yet
    CUDA_TRY(cudaMemsetAsync(acc.tile_count, 0, (size_t)nradar * sizeof(int), st));
    if (d_quality) CUDA_TRY(cudaMemsetAsync(d_quality, 0, nvox * sizeof(double), st));
    net_cull_kernel<<<(unsigned)ntile, kNetThreads, 0, st>>>(A, acc);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    std::vector<int> counts(nradar);
    CUDA_TRY(cudaMemcpyAsync(counts.data(), acc.tile_count, (size_t)nradar * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // Radars whose reach discs are disjoint touch disjoint voxels, so their passes may overlap in time:
    // greedy classes in radar order (class c holds radars that conflict with none already in it), the
    // classes run one after the other, the radars of a class on up to kNetStreams side streams.  The
    // schedule is a function of the geometry alone, so results are reproducible run to run.
    std::vector<int> cls(nradar, -1);
    int nclass = 0;
    for (int i = 0; i < nradar; ++i) {
        // (every radar is classed, reached or not: the order of accumulation at a voxel is then the same
        //  whether the call covers the whole grid or a slab of it)
        for (int c = 0; c <= nclass && cls[i] < 0; ++c) {
            bool ok = true;
            for (int j = 0; j < i && ok; ++j) {
                if (cls[j] != c) continue;
                const double dx = radar_x[i] - radar_x[j], dy = radar_y[i] - radar_y[j], rr = reach[i] + reach[j];
                if (!(dx * dx + dy * dy > rr * rr)) ok = false;
            }
            if (ok) cls[i] = c;
        }
        if (cls[i] == nclass) ++nclass;
    }
    NetStreams* ns = nullptr;
    if ((rc = net_streams(&ns))) return rc;
    CUDA_TRY(cudaEventRecord(ns->ev_class, st));
    for (int c = 0; c < nclass; ++c) {
        int slot = 0, used = 0;
        for (int i = 0; i < nradar; ++i) {
            if (cls[i] != c || counts[i] <= 0) continue;
            const int ne = vols[i]->ne;
            const size_t smem = (size_t)ne * (sizeof(SweepDesc) + sizeof(FastSweep)) +
                                (size_t)(ne - 1) * (sizeof(PairDesc) + sizeof(NetPair)) +
                                (size_t)ne * kNetThreads * sizeof(AzEntry);
            cudaStream_t s2 = ns->side[slot];
            if (slot >= used) { CUDA_TRY(cudaStreamWaitEvent(s2, ns->ev_class, 0)); used = slot + 1; }
            if ((rc = launch_radar_pass_any(A, acc, prm->method, uni, i, counts[i], smem, s2))) return rc;
            g_launches.fetch_add(1);
            CUDA_TRY(cudaGetLastError());
            slot = (slot + 1) % kNetStreams;
        }
        for (int k = 0; k < used; ++k) {
            CUDA_TRY(cudaEventRecord(ns->ev_side[k], ns->side[k]));
            CUDA_TRY(cudaStreamWaitEvent(st, ns->ev_side[k], 0));
        }
        CUDA_TRY(cudaEventRecord(ns->ev_class, st));
    }
    if ((rc = launch_finalize(A, acc, prm->method, st))) return rc;
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return scratch_done(st);
}

}  // namespace

extern "C" int pcg_volume_value_range(const pcg_volume* v, int field, double* vmin, double* vmax, uint64_t* count) {
    if (!v) return fail(PCG_ERR_ARG, "volume handle is NULL");
    if (field < 0 || field >= v->nfield) return fail(PCG_ERR_ARG, "field index %d out of range", field);
    if (vmin) *vmin = v->src_min[field];
    if (vmax) *vmax = v->src_max[field];
    if (count) *count = v->src_count[field];
    return PCG_OK;
}

extern "C" int pcg_network_compose_dev(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                       const double* radar_y, const double* radar_height,
                                       const double* effective_earth_radius, const double* const* beam_widths,
                                       const pcg_network_params* params, const double* d_grid_x,
                                       const double* d_grid_y, int nx, int ny, const double* level_heights, int nz,
                                       double fill, double* d_composite, int16_t* d_valid_count,
                                       int16_t* d_coverage_count, int8_t* d_blind_mask, double* d_cr,
                                       double* d_quality, void* stream) {
    if (nx < 0 || ny < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    std::lock_guard<std::mutex> lock(g_mutex);   // the table scratch is shared
    return enqueue_network(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths, params,
                           d_grid_x, d_grid_y, (long long)nx * ny, ny, level_heights, nz, fill, d_composite,
                           d_valid_count, d_coverage_count, d_blind_mask, d_cr, d_quality, nullptr,
                           (cudaStream_t)stream, fill, nullptr);
}

namespace {
int enqueue_columns(const double* d_vol, const double* d_levels, int nz, long long ncol, double fill, double min_dbz,
                    double max_dbz_cap, double threshold, double* d_cr, double* d_vil, double* d_et,
                    unsigned char* d_topped, cudaStream_t st, int nan_as_fill = 0);
}

// ---- multi-GPU: the finished slab goes straight into every peer's grid --------------------------
extern "C" int pcg_peer_alloc(size_t bytes, void** dptr, unsigned char* handle64) {
    if (!dptr || !handle64) return fail(PCG_ERR_ARG, "NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    *dptr = nullptr;
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(PCG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *dptr);
    if (e != cudaSuccess) { cudaFree(*dptr); *dptr = nullptr; return fail(PCG_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    return PCG_OK;
}

extern "C" int pcg_peer_open(const unsigned char* handle64, void** dptr) {
    if (!dptr || !handle64) return fail(PCG_ERR_ARG, "NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PCG_OK;
}

extern "C" int pcg_peer_close(void* dptr) {
    if (dptr) CUDA_TRY(cudaIpcCloseMemHandle(dptr));
    return PCG_OK;
}

extern "C" int pcg_peer_free(void* dptr) {
    if (dptr) CUDA_TRY(cudaFree(dptr));
    return PCG_OK;
}

extern "C" int pcg_network_compose_gather_dev(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                              const double* radar_y, const double* radar_height,
                                              const double* effective_earth_radius, const double* const* beam_widths,
                                              const pcg_network_params* params, const double* d_grid_x,
                                              const double* d_grid_y, int nx, int ny, const double* level_heights,
                                              int nz, double fill, double* d_composite, int16_t* d_valid_count,
                                              int16_t* d_coverage_count, int8_t* d_blind_mask, double* d_cr,
                                              int npeer, double* const* peer_full, long long full_rows,
                                              long long row_offset, void* stream) {
    if (nx < 0 || ny < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    if (npeer < 0 || npeer > kNetMaxPeers) return fail(PCG_ERR_ARG, "npeer must be in [0, %d]", kNetMaxPeers);
    if (npeer && (!peer_full || row_offset < 0 || row_offset + nx > full_rows))
        return fail(PCG_ERR_ARG, "slab rows [%lld, %lld) do not fit a grid of %lld rows", row_offset, row_offset + nx, full_rows);
    std::lock_guard<std::mutex> lock(g_mutex);
    g_net_peers.n = npeer;
    g_net_peers.plane = full_rows * (long long)ny;
    g_net_peers.off = row_offset * (long long)ny;
    for (int q = 0; q < npeer; ++q) g_net_peers.ptr[q] = peer_full[q];
    const int rc = enqueue_network(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths,
                                   params, d_grid_x, d_grid_y, (long long)nx * ny, ny, level_heights, nz, fill,
                                   d_composite, d_valid_count, d_coverage_count, d_blind_mask, d_cr, nullptr, nullptr,
                                   (cudaStream_t)stream, fill, nullptr);
    g_net_peers.n = 0;
    return rc;
}

extern "C" int pcg_network_compose_ex(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                      const double* radar_y, const double* radar_height,
                                      const double* effective_earth_radius, const double* const* beam_widths,
                                      const pcg_network_params* params, const double* grid_x, const double* grid_y,
                                      int nx, int ny, const double* level_heights, int nz, double fill,
                                      double out_missing, double* composite, int16_t* valid_count,
                                      int16_t* coverage_count, int8_t* blind_mask, double* cr, double* quality,
                                      uint64_t* tie_voxels, const double* column_params, double* vil, double* et,
                                      unsigned char* et_topped);

extern "C" int pcg_network_compose(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                   const double* radar_y, const double* radar_height,
                                   const double* effective_earth_radius, const double* const* beam_widths,
                                   const pcg_network_params* params, const double* grid_x, const double* grid_y,
                                   int nx, int ny, const double* level_heights, int nz, double fill,
                                   double* composite, int16_t* valid_count, int16_t* coverage_count,
                                   int8_t* blind_mask, double* cr, double* quality, uint64_t* tie_voxels) {
    if (!composite) return fail(PCG_ERR_ARG, "composite is NULL");
    return pcg_network_compose_ex(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths,
                                  params, grid_x, grid_y, nx, ny, level_heights, nz, fill, fill, composite,
                                  valid_count, coverage_count, blind_mask, cr, quality, tie_voxels, nullptr, nullptr,
                                  nullptr, nullptr);
}

// Host-buffer network composite of rows [row_offset, row_offset + nx) of a grid of full_rows rows: grid_x / grid_y
// and every output are the FULL host arrays; only the slab's rows are uploaded, computed and written back (each
// level's slab is one contiguous run, so the volumes go back with one pitched copy per output).  One host thread
// per GPU calls this for its slab: every GPU returns its rows over its own PCIe link into the same pinned arrays —
// the device analogue of the reference's fork pool (RadarInterp.py:1816-1861).
// PCG_TRACE=1: the host-buffer network call reports where its time goes (stderr, one line per call; the stream is
// synchronised after each stage, so the stages no longer overlap — a diagnostic, not a timing mode)
struct StageTrace {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    char line[512];
    int len;
    explicit StageTrace(cudaStream_t s) : on(false), st(s), len(0) {
        static const bool enabled = [] { const char* e = getenv("PCG_TRACE"); return e && *e && *e != '0'; }();
        on = enabled;
        line[0] = 0;
        if (on) t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        len += snprintf(line + len, sizeof(line) - (size_t)len, " %s %.2f ms", what,
                        std::chrono::duration<double, std::milli>(t1 - t0).count());
        if (len > (int)sizeof(line) - 1) len = (int)sizeof(line) - 1;
        t0 = t1;
    }
    void flush(const char* who, int device) {
        if (on) fprintf(stderr, "[pcg trace] %s dev %d:%s\n", who, device, line);
    }
};

static int network_compose_host(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                const double* radar_y, const double* radar_height,
                                const double* effective_earth_radius, const double* const* beam_widths,
                                const pcg_network_params* params, const double* grid_x, const double* grid_y,
                                int nx, int ny, const double* level_heights, int nz, double fill,
                                double out_missing, double* composite, int16_t* valid_count,
                                int16_t* coverage_count, int8_t* blind_mask, double* cr, double* quality,
                                uint64_t* tie_voxels, const double* column_params, double* vil, double* et,
                                unsigned char* et_topped, long long full_rows, long long row_offset) {
    if (row_offset < 0 || full_rows < 0 || row_offset + nx > full_rows)
        return fail(PCG_ERR_ARG, "slab rows [%lld, %lld) do not fit a grid of %lld rows", row_offset, row_offset + nx, full_rows);
    const size_t row0 = (size_t)row_offset * (size_t)(ny > 0 ? ny : 0);      // first element of the slab in a plane
    const size_t plane = (size_t)full_rows * (size_t)(ny > 0 ? ny : 0);      // elements of one full level
    grid_x = grid_x ? grid_x + row0 : grid_x;
    grid_y = grid_y ? grid_y + row0 : grid_y;
    if (nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    const bool columns = column_params && (vil || et || et_topped);
    if (!composite && !columns) return fail(PCG_ERR_ARG, "composite is NULL");
    const size_t ncol = (size_t)nx * (size_t)ny;
    const size_t nvox = (size_t)nz * ncol;
    if (tie_voxels) *tie_voxels = 0;
    if (nvox == 0 && !(cr && ncol)) return PCG_OK;
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    StageTrace trace(st);
    // The slab is taken in ROW CHUNKS through two sets of device buffers: while chunk c's results cross PCIe on the
    // copy stream, chunk c + 1's grid rows are uploaded and its kernels run on the compute stream (tiles are
    // independent, so chunking by rows repeats no work; the radar tables are built once and cached).  The call is
    // then bound by the result's trip over PCIe alone instead of upload + kernels + download one after the other.
    // PCG_TRACE, the per-radar quality output and small slabs use one chunk.
    int nchunk = 1;
    if (!trace.on && !quality && use_tile_kernel() && nz > 0 && nx >= 256) {
        nchunk = nx / 128;
        if (nchunk > 6) nchunk = 6;
    }
    if (const char* e = getenv("PCG_NET_CHUNKS")) {          // tests: force a chunk count on small grids
        const int want = atoi(e);
        if (want >= 1 && !quality && use_tile_kernel() && nz > 0) nchunk = want < nx ? want : (nx > 0 ? nx : 1);
    }
    const int rows_cap = (nx + nchunk - 1) / nchunk;
    const size_t ccol = (size_t)rows_cap * (size_t)ny, cvox = (size_t)nz * ccol;     // capacity of one buffer set
    const int nset = nchunk > 1 ? 2 : 1;
    if ((rc = g_scratch[S_GX].ensure(nset * ccol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(nset * ccol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(nset * (cvox ? cvox : 1) * sizeof(double)))) return rc;
    if (valid_count && (rc = g_scratch[S_AUX0].ensure(nset * cvox * sizeof(int16_t)))) return rc;
    if (coverage_count && (rc = g_scratch[S_AUX1].ensure(nset * cvox * sizeof(int16_t)))) return rc;
    if (blind_mask && (rc = g_scratch[S_AUX2].ensure(nset * cvox * sizeof(int8_t)))) return rc;
    if (cr && (rc = g_scratch[S_AUX3].ensure(nset * ccol * sizeof(double)))) return rc;
    if (quality && (rc = g_scratch[S_AUX4].ensure(cvox * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_TMP2].ensure(sizeof(unsigned long long)))) return rc;
    const size_t col_bytes = 2 * ccol * sizeof(double) + ((ccol + 7) & ~(size_t)7);     // VIL, ET, topped of one set
    if (columns && ncol) {
        if ((rc = g_scratch[S_TMP0].ensure((size_t)(nz ? nz : 1) * sizeof(double)))) return rc;
        if ((rc = g_scratch[S_ACC0].ensure(nset * col_bytes))) return rc;
        if (nz) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_TMP0].ptr, level_heights, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemsetAsync(g_scratch[S_TMP2].ptr, 0, sizeof(unsigned long long), st));
    trace.mark("scratch");
    NetStreams* ns = nullptr;
    cudaStream_t cp = st;                      // copy-out stream (the compute stream itself when there is one chunk)
    if (nchunk > 1) {
        if ((rc = net_streams(&ns))) return rc;
        cp = ns->side[0];
    }
    // the tile kernel stores the caller's missing-value marker itself; the column products read the composite
    // afterwards and know `fill` and NaN as "no data" — any other marker is applied after them, as before
    const bool marker_ok = !columns || out_missing != out_missing || memcmp(&out_missing, &fill, sizeof(double)) == 0;
    const int nan_as_fill = (columns && out_missing != out_missing) ? 1 : 0;
    const double tile_missing = marker_ok ? out_missing : fill;
    for (int c = 0; c < nchunk; ++c) {
        const int b = c & 1;
        const int r_lo = (int)(((long long)nx * c) / nchunk), r_hi = (int)(((long long)nx * (c + 1)) / nchunk);
        const int rows = r_hi - r_lo;
        const size_t kcol = (size_t)rows * (size_t)ny, kvox = (size_t)nz * kcol;     // this chunk
        const size_t h0 = row0 + (size_t)r_lo * (size_t)ny;                          // its first element in a host plane
        double* d_gx = (double*)g_scratch[S_GX].ptr + b * ccol;
        double* d_gy = (double*)g_scratch[S_GY].ptr + b * ccol;
        double* d_out = (double*)g_scratch[S_OUT].ptr + b * cvox;
        int16_t* d_valid = valid_count ? (int16_t*)g_scratch[S_AUX0].ptr + b * cvox : nullptr;
        int16_t* d_cov = coverage_count ? (int16_t*)g_scratch[S_AUX1].ptr + b * cvox : nullptr;
        int8_t* d_blind = blind_mask ? (int8_t*)g_scratch[S_AUX2].ptr + b * cvox : nullptr;
        double* d_cr = cr ? (double*)g_scratch[S_AUX3].ptr + b * ccol : nullptr;
        // this buffer set is free again once the copies of chunk c - 2 have left it
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(st, ns->ev_side[2 + b], 0));
        CUDA_TRY(cudaMemcpyAsync(d_gx, grid_x + (size_t)r_lo * (size_t)ny, kcol * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_gy, grid_y + (size_t)r_lo * (size_t)ny, kcol * sizeof(double), cudaMemcpyHostToDevice, st));
        trace.mark("grid H2D");
        bool relabelled = false;
        rc = enqueue_network(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths, params,
                             d_gx, d_gy, (long long)kcol, ny, level_heights, nz, fill, d_out, d_valid, d_cov, d_blind, d_cr,
                             quality ? (double*)g_scratch[S_AUX4].ptr : nullptr,
                             (unsigned long long*)g_scratch[S_TMP2].ptr, st, tile_missing, &relabelled);
        if (rc) return rc;
        trace.mark("tables + kernels");
        double* d2 = nullptr;
        unsigned char* dt = nullptr;
        if (columns && kcol) {
            // VIL / ET of the composite while it is still on the device (RadarInterp.py:2064-2119 -> derive_vil /
            // derive_et, RadarProduct.py:33-85): no second trip of the volume over PCIe
            d2 = (double*)((char*)g_scratch[S_ACC0].ptr + b * col_bytes);
            dt = (unsigned char*)(d2 + 2 * ccol);
            rc = enqueue_columns(d_out, (const double*)g_scratch[S_TMP0].ptr, nz, (long long)kcol, fill, column_params[0],
                                 column_params[1], column_params[2], nullptr, vil ? d2 : nullptr, et ? d2 + ccol : nullptr,
                                 et_topped ? dt : nullptr, st, relabelled ? nan_as_fill : 0);
            if (rc) return rc;
        }
        if (composite && kvox && !(relabelled && marker_ok) &&
            (rc = relabel_missing(d_out, kvox, fill, out_missing, st))) return rc;
        if (nchunk > 1) {
            CUDA_TRY(cudaEventRecord(ns->ev_side[b], st));
            CUDA_TRY(cudaStreamWaitEvent(cp, ns->ev_side[b], 0));
        }
        if (d2) {
            if (vil) CUDA_TRY(cudaMemcpyAsync(vil + h0, d2, kcol * sizeof(double), cudaMemcpyDeviceToHost, cp));
            if (et) CUDA_TRY(cudaMemcpyAsync(et + h0, d2 + ccol, kcol * sizeof(double), cudaMemcpyDeviceToHost, cp));
            if (et_topped) CUDA_TRY(cudaMemcpyAsync(et_topped + h0, dt, kcol, cudaMemcpyDeviceToHost, cp));
        }
        trace.mark("column products + D2H");
        if (composite && kvox)
            CUDA_TRY(cudaMemcpy2DAsync(composite + h0, plane * sizeof(double), d_out, kcol * sizeof(double), kcol * sizeof(double), (size_t)nz, cudaMemcpyDeviceToHost, cp));
        trace.mark("composite D2H");
        if (valid_count && kvox) CUDA_TRY(cudaMemcpy2DAsync(valid_count + h0, plane * sizeof(int16_t), d_valid, kcol * sizeof(int16_t), kcol * sizeof(int16_t), (size_t)nz, cudaMemcpyDeviceToHost, cp));
        if (coverage_count && kvox) CUDA_TRY(cudaMemcpy2DAsync(coverage_count + h0, plane * sizeof(int16_t), d_cov, kcol * sizeof(int16_t), kcol * sizeof(int16_t), (size_t)nz, cudaMemcpyDeviceToHost, cp));
        if (blind_mask && kvox) CUDA_TRY(cudaMemcpy2DAsync(blind_mask + h0, plane * sizeof(int8_t), d_blind, kcol * sizeof(int8_t), kcol * sizeof(int8_t), (size_t)nz, cudaMemcpyDeviceToHost, cp));
        if (cr && kcol) CUDA_TRY(cudaMemcpyAsync(cr + h0, d_cr, kcol * sizeof(double), cudaMemcpyDeviceToHost, cp));
        if (quality && kvox) CUDA_TRY(cudaMemcpy2DAsync(quality + h0, plane * sizeof(double), g_scratch[S_AUX4].ptr, kcol * sizeof(double), kcol * sizeof(double), (size_t)nz, cudaMemcpyDeviceToHost, cp));
        if (nchunk > 1) CUDA_TRY(cudaEventRecord(ns->ev_side[2 + b], cp));
    }
    unsigned long long ties = 0;
    CUDA_TRY(cudaMemcpyAsync(&ties, g_scratch[S_TMP2].ptr, sizeof(ties), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (cp != st) {
        CUDA_TRY(cudaStreamSynchronize(cp));
        // later calls on the library stream must also see the copy stream drained (it is, here) — nothing to record
    }
    trace.mark("counts / masks / CR D2H");
    {
        int dev = 0;
        cudaGetDevice(&dev);
        trace.flush("network_compose_host", dev);
    }
    if (tie_voxels) *tie_voxels = ties;
    return PCG_OK;
}

extern "C" int pcg_network_compose_ex(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                      const double* radar_y, const double* radar_height,
                                      const double* effective_earth_radius, const double* const* beam_widths,
                                      const pcg_network_params* params, const double* grid_x, const double* grid_y,
                                      int nx, int ny, const double* level_heights, int nz, double fill,
                                      double out_missing, double* composite, int16_t* valid_count,
                                      int16_t* coverage_count, int8_t* blind_mask, double* cr, double* quality,
                                      uint64_t* tie_voxels, const double* column_params, double* vil, double* et,
                                      unsigned char* et_topped) {
    return network_compose_host(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths, params,
                                grid_x, grid_y, nx, ny, level_heights, nz, fill, out_missing, composite, valid_count,
                                coverage_count, blind_mask, cr, quality, tie_voxels, column_params, vil, et, et_topped,
                                (long long)nx, 0);
}

extern "C" int pcg_network_compose_rows(int nradar, const pcg_volume* const* vols, const double* radar_x,
                                        const double* radar_y, const double* radar_height,
                                        const double* effective_earth_radius, const double* const* beam_widths,
                                        const pcg_network_params* params, const double* grid_x, const double* grid_y,
                                        long long full_rows, long long row_offset, int nrows, int ny,
                                        const double* level_heights, int nz, double fill, double out_missing,
                                        double* composite, int16_t* valid_count, int16_t* coverage_count,
                                        int8_t* blind_mask, double* cr, uint64_t* tie_voxels,
                                        const double* column_params, double* vil, double* et,
                                        unsigned char* et_topped) {
    return network_compose_host(nradar, vols, radar_x, radar_y, radar_height, effective_earth_radius, beam_widths, params,
                                grid_x, grid_y, nrows, ny, level_heights, nz, fill, out_missing, composite, valid_count,
                                coverage_count, blind_mask, cr, nullptr, tie_voxels, column_params, vil, et, et_topped,
                                full_rows, row_offset);
}

// ---- compose_network_volume on gridded volumes, column products ---------------------------------

extern "C" int pcg_compose_volumes(int nradar, const double* const* volumes, const double* const* quality,
                                   const double* weight2d, int method, int nz, int nx, int ny, double fill,
                                   double* composite, int16_t* valid_count) {
    if (nradar < 1) return fail(PCG_ERR_ARG, "radar_volumes must contain at least one volume.");
    if (nradar > kComposeMaxRadars) return fail(PCG_ERR_ARG, "at most %d radar volumes per call", kComposeMaxRadars);
    if (method < 0 || method > 3) return fail(PCG_ERR_ARG, "Unsupported composite_method code %d", method);
    if (nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    if (method != 0 && !weight2d) return fail(PCG_ERR_ARG, "weight2d is required for this method");
    if (!volumes || !composite) return fail(PCG_ERR_ARG, "NULL pointer");
    const size_t ncol = (size_t)nx * (size_t)ny, nvox = ncol * (size_t)nz;
    if (nvox == 0) return PCG_OK;
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    const bool has_q = (method == 3) && quality;
    if ((rc = g_scratch[S_AUX4].ensure((size_t)nradar * nvox * sizeof(double)))) return rc;
    if (has_q && (rc = g_scratch[S_IDX].ensure((size_t)nradar * nvox * sizeof(double)))) return rc;
    if (weight2d && (rc = g_scratch[S_AUX3].ensure((size_t)nradar * ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(nvox * sizeof(double)))) return rc;
    if (valid_count && (rc = g_scratch[S_AUX0].ensure(nvox * sizeof(int16_t)))) return rc;
    ComposeArgs A;
    memset(&A, 0, sizeof(A));
    for (int r = 0; r < nradar; ++r) {
        if (!volumes[r] || (has_q && !quality[r])) return fail(PCG_ERR_ARG, "NULL volume pointer (radar %d)", r);
        double* dv = (double*)g_scratch[S_AUX4].ptr + (size_t)r * nvox;
        CUDA_TRY(cudaMemcpyAsync(dv, volumes[r], nvox * sizeof(double), cudaMemcpyHostToDevice, st));
        A.vol[r] = dv;
        if (has_q) {
            double* dq = (double*)g_scratch[S_IDX].ptr + (size_t)r * nvox;
            CUDA_TRY(cudaMemcpyAsync(dq, quality[r], nvox * sizeof(double), cudaMemcpyHostToDevice, st));
            A.qual[r] = dq;
        }
    }
    if (weight2d) {
        CUDA_TRY(cudaMemcpyAsync(g_scratch[S_AUX3].ptr, weight2d, (size_t)nradar * ncol * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        A.w2d = (const double*)g_scratch[S_AUX3].ptr;
    }
    A.nradar = nradar; A.method = method; A.has_quality = has_q ? 1 : 0;
    A.ncol = (long long)ncol; A.nvox = (long long)nvox; A.fill = fill;
    A.out = (double*)g_scratch[S_OUT].ptr;
    A.count = valid_count ? (int16_t*)g_scratch[S_AUX0].ptr : nullptr;
    const unsigned blocks = (unsigned)((nvox + 255) / 256 < 148 * 32 ? (nvox + 255) / 256 : 148 * 32);
    compose_kernel<<<blocks, 256, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(composite, A.out, nvox * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (valid_count) CUDA_TRY(cudaMemcpyAsync(valid_count, A.count, nvox * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

namespace {
int enqueue_columns(const double* d_vol, const double* d_levels, int nz, long long ncol, double fill, double min_dbz,
                    double max_dbz_cap, double threshold, double* d_cr, double* d_vil, double* d_et,
                    unsigned char* d_topped, cudaStream_t st, int nan_as_fill) {
    ColumnArgs A;
    A.nan_as_fill = nan_as_fill;
    A.vol = d_vol; A.levels = d_levels; A.nz = nz; A.ncol = ncol; A.fill = fill;
    A.min_dbz = min_dbz; A.max_dbz_cap = max_dbz_cap; A.threshold = threshold;
    A.cr = d_cr; A.vil = d_vil; A.et = d_et; A.topped = d_topped;
    const unsigned blocks = (unsigned)((ncol + 255) / 256 < 148 * 32 ? (ncol + 255) / 256 : 148 * 32);
    column_kernel<<<blocks, 256, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return PCG_OK;
}
}  // namespace

extern "C" int pcg_column_products_dev(const double* d_volume, const double* d_levels, int nz, int nx, int ny,
                                       double fill, double min_dbz, double max_dbz_cap, double threshold_dbz,
                                       double* d_cr, double* d_vil, double* d_et, unsigned char* d_topped,
                                       void* stream) {
    if (nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    if ((size_t)nx * ny == 0) return PCG_OK;
    return enqueue_columns(d_volume, d_levels, nz, (long long)nx * ny, fill, min_dbz, max_dbz_cap, threshold_dbz, d_cr,
                           d_vil, d_et, d_topped, (cudaStream_t)stream);
}

extern "C" int pcg_column_products(const double* volume, const double* level_heights, int nz, int nx, int ny,
                                   double fill, double min_dbz, double max_dbz_cap, double threshold_dbz, double* cr,
                                   double* vil, double* et, unsigned char* topped) {
    if (nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "grid shape must be non-negative");
    const size_t ncol = (size_t)nx * (size_t)ny, nvox = ncol * (size_t)nz;
    if (ncol == 0) return PCG_OK;
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_OUT].ensure((nvox ? nvox : 1) * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_TMP0].ensure((size_t)(nz ? nz : 1) * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_AUX3].ensure(3 * ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_AUX2].ensure(ncol))) return rc;
    if (nvox) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_OUT].ptr, volume, nvox * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nz) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_TMP0].ptr, level_heights, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, st));
    double* d3 = (double*)g_scratch[S_AUX3].ptr;
    rc = enqueue_columns((const double*)g_scratch[S_OUT].ptr, (const double*)g_scratch[S_TMP0].ptr, nz, (long long)ncol,
                         fill, min_dbz, max_dbz_cap, threshold_dbz, cr ? d3 : nullptr, vil ? d3 + ncol : nullptr,
                         et ? d3 + 2 * ncol : nullptr, topped ? (unsigned char*)g_scratch[S_AUX2].ptr : nullptr, st);
    if (rc) return rc;
    if (cr) CUDA_TRY(cudaMemcpyAsync(cr, d3, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (vil) CUDA_TRY(cudaMemcpyAsync(vil, d3 + ncol, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (et) CUDA_TRY(cudaMemcpyAsync(et, d3 + 2 * ncol, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (topped) CUDA_TRY(cudaMemcpyAsync(topped, g_scratch[S_AUX2].ptr, ncol, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

// ---- stand-alone observability mask / quality weight volume --------------------------------------

extern "C" int pcg_network_gates(int ne, const double* fix_elev, const double* first_gate, const double* last_gate,
                                 const double* beam_widths, double radar_height, double effective_earth_radius,
                                 double radar_x, double radar_y, const double* grid_x, const double* grid_y, int nx,
                                 int ny, const double* level_heights, int nz, int quality_terms, double range_decay_m,
                                 double beam_radius_ref_m, double vertical_gap_ref_deg, unsigned char* mask,
                                 double* quality) {
    if (!(effective_earth_radius > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    if (ne < 0 || nx < 0 || ny < 0 || nz < 0) return fail(PCG_ERR_ARG, "sizes must be non-negative");
    const size_t ncol = (size_t)nx * (size_t)ny, nvox = ncol * (size_t)nz;
    if (nvox == 0) return PCG_OK;
    if (ne == 0) {                                   // RadarInterp.py:555-556 / 625-626
        if (mask) memset(mask, 0, nvox);
        if (quality) for (size_t i = 0; i < nvox; ++i) quality[i] = 0.0;
        return PCG_OK;
    }
    const int npair = ne > 1 ? ne - 1 : 1;
    std::vector<GateSweep> sw(ne);
    std::vector<NetPair> pr(npair);
    std::vector<double> btan(npair), gap(npair);
    std::vector<NetLevel> lv(nz);
    for (int k = 0; k < ne; ++k) { sw[k].elev = fix_elev[k]; sw[k].r_first = first_gate[k]; sw[k].r_last = last_gate[k]; }
    memset(pr.data(), 0, pr.size() * sizeof(NetPair));
    if (ne == 1) {
        btan[0] = tan((beam_widths ? beam_widths[0] : 1.0) * kDeg2Rad * 0.5);
        gap[0] = 0.0;
    }
    for (int k = 0; k + 1 < ne; ++k) {
        pr[k].obs_first = first_gate[k] > first_gate[k + 1] ? first_gate[k] : first_gate[k + 1];
        pr[k].obs_last = last_gate[k] < last_gate[k + 1] ? last_gate[k] : last_gate[k + 1];
        const double rep = 0.5 * ((beam_widths ? beam_widths[k] : 1.0) + (beam_widths ? beam_widths[k + 1] : 1.0));
        btan[k] = tan(rep * kDeg2Rad * 0.5);         // np.tan(np.deg2rad(rep) * 0.5)
        gap[k] = 0.5 * (fix_elev[k + 1] - fix_elev[k]);
    }
    volatile double av = effective_earth_radius + radar_height;
    const double a2_np = pow((double)av, 2.0);
    volatile double twoa = 2.0 * av;
    for (int k = 0; k < nz; ++k) {
        memset(&lv[k], 0, sizeof(NetLevel));
        const double g = effective_earth_radius + level_heights[k];
        lv[k].g2_np = pow(g, 2.0);
        lv[k].a2g2_np = a2_np + lv[k].g2_np;
        volatile double tg = twoa * g;
        lv[k].twoag_np = tg;
    }
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    const size_t b_sw = sw.size() * sizeof(GateSweep), b_pr = pr.size() * sizeof(NetPair),
                 b_d = (size_t)npair * sizeof(double), b_lv = lv.size() * sizeof(NetLevel);
    if ((rc = g_scratch[S_NET].ensure(b_sw + b_pr + 2 * b_d + b_lv + 64))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double)))) return rc;
    if ((rc = g_scratch[S_GY].ensure(ncol * sizeof(double)))) return rc;
    if (mask && (rc = g_scratch[S_AUX2].ensure(nvox))) return rc;
    if (quality && (rc = g_scratch[S_AUX4].ensure(nvox * sizeof(double)))) return rc;
    char* base = (char*)g_scratch[S_NET].ptr;
    CUDA_TRY(cudaMemcpyAsync(base, sw.data(), b_sw, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + b_sw, pr.data(), b_pr, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + b_sw + b_pr, btan.data(), b_d, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + b_sw + b_pr + b_d, gap.data(), b_d, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + b_sw + b_pr + 2 * b_d, lv.data(), b_lv, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    GatesArgs A;
    memset(&A, 0, sizeof(A));
    A.sweeps = (const GateSweep*)base;
    A.pairs = (const NetPair*)(base + b_sw);
    A.beam_tan = (const double*)(base + b_sw + b_pr);
    A.gap_half = (const double*)(base + b_sw + b_pr + b_d);
    A.levels = (const NetLevel*)(base + b_sw + b_pr + 2 * b_d);
    A.ne = ne; A.nz = nz;
    A.gx = (const double*)g_scratch[S_GX].ptr; A.gy = (const double*)g_scratch[S_GY].ptr; A.ncol = (long long)ncol;
    A.radar_x = radar_x; A.radar_y = radar_y; A.Re = effective_earth_radius; A.a2_np = a2_np; A.twoa = twoa;
    double half = (beam_widths ? beam_widths[0] : 1.0) * 0.5;     // RadarInterp.py:573 / 653
    A.beam_half = half < 0.1 ? 0.1 : half;
    A.terms = quality_terms; A.range_decay = range_decay_m; A.beam_ref = beam_radius_ref_m; A.gap_ref = vertical_gap_ref_deg;
    A.mask = mask ? (unsigned char*)g_scratch[S_AUX2].ptr : nullptr;
    A.quality = quality ? (double*)g_scratch[S_AUX4].ptr : nullptr;
    gates_kernel<<<(unsigned)((ncol + 255) / 256), 256, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    if (mask) CUDA_TRY(cudaMemcpyAsync(mask, A.mask, nvox, cudaMemcpyDeviceToHost, st));
    if (quality) CUDA_TRY(cudaMemcpyAsync(quality, A.quality, nvox * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

// ---- wind-volume gridding step (retrieve_wind_volume_xy) ------------------------------------------

extern "C" int pcg_wind_gather(const double* samples, long long ns, const double* target_x, const double* target_y,
                               long long nt, double radius_m, double max_radius_m, int min_neighbors, double power,
                               double* out_f64, int32_t* out_i32, unsigned char* expanded) {
    if (ns < 0 || nt < 0) return fail(PCG_ERR_ARG, "sizes must be non-negative");
    if (!out_f64 || !out_i32 || !expanded) return fail(PCG_ERR_ARG, "NULL output pointer");
    if (nt == 0) return PCG_OK;
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    const size_t b_s = (size_t)kWindFields * (size_t)(ns ? ns : 1) * sizeof(double), b_t = (size_t)nt * sizeof(double);
    if ((rc = g_scratch[S_AUX4].ensure(b_s))) return rc;
    if ((rc = g_scratch[S_GX].ensure(b_t)) || (rc = g_scratch[S_GY].ensure(b_t))) return rc;
    if ((rc = g_scratch[S_OUT].ensure(5 * b_t)) || (rc = g_scratch[S_IDX].ensure(3 * (size_t)nt * sizeof(int))) ||
        (rc = g_scratch[S_AUX2].ensure((size_t)nt)))
        return rc;
    if (ns) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_AUX4].ptr, samples, (size_t)kWindFields * ns * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, target_x, b_t, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, target_y, b_t, cudaMemcpyHostToDevice, st));
    WindGatherArgs A;
    A.samples = (const double*)g_scratch[S_AUX4].ptr; A.ns = ns;
    A.tx = (const double*)g_scratch[S_GX].ptr; A.ty = (const double*)g_scratch[S_GY].ptr; A.nt = nt;
    A.r1sq = radius_m * radius_m;
    A.r2sq = max_radius_m * max_radius_m;
    A.min_neighbors = min_neighbors; A.power = power;
    double* o = (double*)g_scratch[S_OUT].ptr;
    A.u = o; A.v = o + nt; A.z = o + 2 * nt; A.rmse = o + 3 * nt; A.cond = o + 4 * nt;
    int* oi = (int*)g_scratch[S_IDX].ptr;
    A.valid_count = oi; A.neighbor_count = oi + nt; A.azsec = oi + 2 * nt;
    A.expanded = (unsigned char*)g_scratch[S_AUX2].ptr;
    wind_gather_kernel<<<(unsigned)((nt + kWindTile - 1) / kWindTile), kWindTile, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_f64, o, 5 * b_t, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_i32, oi, 3 * (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(expanded, A.expanded, (size_t)nt, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

extern "C" int pcg_wind_vertical(const double* sweeps, int nsweep, long long ncol, const double* level_heights, int nz,
                                 double vertical_tolerance_m, double max_vertical_gap_m, double* out_f64,
                                 int32_t* out_i32, unsigned char* quality_flag) {
    if (nsweep < 0 || nsweep > kWindMaxSweeps) return fail(PCG_ERR_ARG, "nsweep must be in [0, %d]", kWindMaxSweeps);
    if (ncol < 0 || nz < 0) return fail(PCG_ERR_ARG, "sizes must be non-negative");
    if (!out_f64 || !out_i32 || !quality_flag) return fail(PCG_ERR_ARG, "NULL output pointer");
    const size_t nvox = (size_t)nz * (size_t)ncol;
    if (nvox == 0) return PCG_OK;
    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    const size_t n_in = (size_t)7 * (size_t)(nsweep ? nsweep : 1) * (size_t)ncol;
    if ((rc = g_scratch[S_AUX4].ensure(n_in * sizeof(double))) || (rc = g_scratch[S_TMP0].ensure((size_t)nz * sizeof(double))) ||
        (rc = g_scratch[S_OUT].ensure(5 * nvox * sizeof(double))) || (rc = g_scratch[S_IDX].ensure(4 * nvox * sizeof(int))) ||
        (rc = g_scratch[S_AUX2].ensure(nvox)))
        return rc;
    if (nsweep) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_AUX4].ptr, sweeps, (size_t)7 * nsweep * ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_TMP0].ptr, level_heights, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, st));
    WindVerticalArgs A;
    const double* in = (const double*)g_scratch[S_AUX4].ptr;
    const size_t plane = (size_t)nsweep * (size_t)ncol;
    A.sz = in; A.su = in + plane; A.sv = in + 2 * plane; A.srmse = in + 3 * plane;
    A.scount = in + 4 * plane; A.sneigh = in + 5 * plane; A.sexp = in + 6 * plane;
    A.nsweep = nsweep; A.ncol = ncol;
    A.levels = (const double*)g_scratch[S_TMP0].ptr; A.nz = nz;
    A.tol = vertical_tolerance_m; A.max_gap = max_vertical_gap_m;
    double* o = (double*)g_scratch[S_OUT].ptr;
    A.u = o; A.v = o + nvox; A.rmse = o + 2 * nvox; A.spread = o + 3 * nvox; A.score = o + 4 * nvox;
    int* oi = (int*)g_scratch[S_IDX].ptr;
    A.valid_count = oi; A.source_count = oi + nvox; A.neighbor_count = oi + 2 * nvox; A.support = oi + 3 * nvox;
    A.flag = (unsigned char*)g_scratch[S_AUX2].ptr;
    wind_vertical_kernel<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_f64, o, 5 * nvox * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_i32, oi, 4 * nvox * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(quality_flag, A.flag, nvox, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

// ---- vertical sections (f3) --------------------------------------------------------------------
extern "C" int pcg_section_sample(const pcg_volume* vol, int field, int npts, const double* x, const double* y,
                                  double radar_height, double effective_earth_radius, const double* target_az,
                                  const double* target_range, int per_sweep_targets, double* out_value,
                                  double* out_az, double* out_range, double* out_z) {
    int rc;
    if ((rc = check_volume(vol))) return rc;
    if (vol->dtype != PCG_VALUE_F64)
        return fail(PCG_ERR_ARG, "section sampling needs a float64 volume (value_dtype = PCG_VALUE_F64)");
    if (field < 0 || field >= vol->nfield) return fail(PCG_ERR_ARG, "field %d out of range", field);
    if (npts < 0) return fail(PCG_ERR_ARG, "npts must be non-negative");
    if (!out_value) return fail(PCG_ERR_ARG, "NULL output pointer");
    const bool geometry = x != nullptr;
    if (geometry && (!y || !out_az || !out_range || !out_z))
        return fail(PCG_ERR_ARG, "geometry mode needs y, out_az, out_range and out_z");
    if (!geometry && (!target_az || !target_range)) return fail(PCG_ERR_ARG, "explicit mode needs target_az and target_range");
    if (geometry && !(effective_earth_radius > 0.0)) return fail(PCG_ERR_ARG, "effective_earth_radius must be positive");
    const int ne = vol->ne;
    const size_t n = (size_t)ne * (size_t)npts;
    if (n == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    const size_t n_in = geometry ? (size_t)npts : (per_sweep_targets ? n : (size_t)npts);
    if ((rc = g_scratch[S_GX].ensure(n_in * sizeof(double))) || (rc = g_scratch[S_GY].ensure(n_in * sizeof(double))) ||
        (rc = g_scratch[S_OUT].ensure(4 * n * sizeof(double))))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, geometry ? x : target_az, n_in * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, geometry ? y : target_range, n_in * sizeof(double), cudaMemcpyHostToDevice, st));
    SectionArgs A;
    A.vol = view_of(vol);
    A.field = field;
    A.npts = npts;
    A.geometry = geometry ? 1 : 0;
    A.per_sweep = per_sweep_targets ? 1 : 0;
    A.h = radar_height;
    A.Re = effective_earth_radius;
    A.x = A.t_az = (const double*)g_scratch[S_GX].ptr;
    A.y = A.t_rng = (const double*)g_scratch[S_GY].ptr;
    A.fill = vol->fill;
    double* o = (double*)g_scratch[S_OUT].ptr;
    A.out_val = o; A.out_az = o + n; A.out_rng = o + 2 * n; A.out_z = o + 3 * n;
    dim3 grid((unsigned)std::min<size_t>(((size_t)npts + 127) / 128, 4096), (unsigned)ne);
    section_kernel<<<grid, 128, 0, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_value, o, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (geometry) {
        CUDA_TRY(cudaMemcpyAsync(out_az, o + n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(out_range, o + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(out_z, o + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

// ---- local VVP (f4) ----------------------------------------------------------------------------
extern "C" int pcg_vvp_local(const double* azimuth, const double* elevation, const double* velocity, int naz,
                             int nrange, int az_num, int bin_num, double min_valid_fraction, int min_valid_count,
                             double min_azimuth_coverage_deg, int min_azimuth_sector_count,
                             double max_condition_number, double* out_f64, int32_t* out_i32) {
    if (az_num < 1 || az_num % 2 != 1) return fail(PCG_ERR_ARG, "az_num must be odd");
    if (bin_num < 1 || bin_num % 2 != 1) return fail(PCG_ERR_ARG, "bin_num must be odd");
    if (naz < 0 || nrange < 0) return fail(PCG_ERR_ARG, "sizes must be non-negative");
    if (naz > 65535) return fail(PCG_ERR_ARG, "at most 65535 rays per sweep");
    if (!out_f64 || !out_i32) return fail(PCG_ERR_ARG, "NULL output pointer");
    const size_t cells = (size_t)naz * (size_t)nrange;
    if (cells == 0) return PCG_OK;
    if (!azimuth || !elevation || !velocity) return fail(PCG_ERR_ARG, "NULL input pointer");
    const size_t nwin = (size_t)az_num * (size_t)bin_num;
    auto smem_for = [&](int warps) {
        return (4 * (size_t)az_num + (size_t)bin_num + (size_t)warps * (2 * nwin + 32 + (size_t)((az_num + 1) / 2))) * sizeof(double);
    };
    int warps = kVvpMaxWarps;
    while (warps > 1 && smem_for(warps) > 200 * 1024) warps >>= 1;
    if (smem_for(warps) > 227 * 1024) return fail(PCG_ERR_ARG, "window of %zu samples does not fit in shared memory", nwin);

    // per-ray design terms (_design_matrix :93-106), azimuth mod 360, window weight factors (:1832-1840)
    std::vector<double> host((size_t)3 * naz + az_num + bin_num);
    double* h_d0 = host.data();
    double* h_d1 = h_d0 + naz;
    double* h_azm = h_d1 + naz;
    double* h_wa = h_azm + naz;
    double* h_wb = h_wa + az_num;
    for (int i = 0; i < naz; ++i) {
        const double ar = azimuth[i] * kDeg2Rad, er = elevation[i] * kDeg2Rad;
        const double ce = cos(er);
        h_d0[i] = sin(ar) * ce;
        h_d1[i] = cos(ar) * ce;
        double m = std::isfinite(azimuth[i]) ? fmod(azimuth[i], 360.0) : NAN;
        if (m != 0.0) { if (m < 0.0) m += 360.0; } else m = 0.0;
        h_azm[i] = m;
    }
    const double sig_a = std::max(1.0, (az_num - 1) / 3.0), sig_b = std::max(1.0, (bin_num - 1) / 3.0);
    for (int i = 0; i < az_num; ++i) { const double t = ((double)i - (az_num - 1) / 2.0) / sig_a; h_wa[i] = exp(-0.5 * (t * t)); }
    for (int i = 0; i < bin_num; ++i) { const double t = ((double)i - (bin_num - 1) / 2.0) / sig_b; h_wb[i] = exp(-0.5 * (t * t)); }

    int rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_AUX4].ensure(cells * sizeof(double))) || (rc = g_scratch[S_TMP0].ensure(host.size() * sizeof(double))) ||
        (rc = g_scratch[S_OUT].ensure(9 * cells * sizeof(double))) || (rc = g_scratch[S_IDX].ensure(2 * cells * sizeof(int))))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_AUX4].ptr, velocity, cells * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_TMP0].ptr, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    VvpArgs A;
    A.vel = (const double*)g_scratch[S_AUX4].ptr;
    const double* t = (const double*)g_scratch[S_TMP0].ptr;
    A.d0 = t; A.d1 = t + naz; A.azm = t + 2 * (size_t)naz; A.w_az = t + 3 * (size_t)naz; A.w_bin = A.w_az + az_num;
    A.naz = naz; A.nrange = nrange; A.az_num = az_num; A.bin_num = bin_num;
    A.min_valid_fraction = min_valid_fraction; A.min_valid_count = min_valid_count;
    A.min_coverage = min_azimuth_coverage_deg; A.min_sector_count = min_azimuth_sector_count;
    A.max_cond = max_condition_number;
    double* o = (double*)g_scratch[S_OUT].ptr;
    A.u = o; A.v = o + cells; A.speed = o + 2 * cells; A.direction = o + 3 * cells; A.offset = o + 4 * cells;
    A.rmse = o + 5 * cells; A.cond = o + 6 * cells; A.valid_fraction = o + 7 * cells; A.coverage = o + 8 * cells;
    int* oi = (int*)g_scratch[S_IDX].ptr;
    A.sector_count = oi; A.valid_count = oi + cells;
    const size_t smem = smem_for(warps);
    CUDA_TRY(cudaFuncSetAttribute(vvp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((nrange + warps - 1) / warps), (unsigned)naz);
    vvp_kernel<<<grid, warps * 32, smem, st>>>(A);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_f64, o, 9 * cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_i32, oi, 2 * cells * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PCG_OK;
}

// ---- f1 fused: grid one radar's volume and reduce its columns without returning the volume ------
extern "C" int pcg_get_column_products(const pcg_volume* vol, double radar_height, double effective_earth_radius,
                                       const double* grid_x, const double* grid_y, int nx, int ny,
                                       const double* level_heights, int nz, double fill, int blind_code,
                                       double min_dbz, double max_dbz_cap, double threshold_dbz, double* cr,
                                       double* vil, double* et, unsigned char* et_topped) {
    int rc = check_volume(vol);
    if (rc) return rc;
    if ((rc = check_common(effective_earth_radius, nx, ny))) return rc;
    if (nz < 0) return fail(PCG_ERR_ARG, "nz must be non-negative");
    if (vol->nfield != 1) return fail(PCG_ERR_ARG, "column products take a single-moment volume");
    const size_t ncol = (size_t)nx * (size_t)ny, nvox = ncol * (size_t)nz;
    if (ncol == 0) return PCG_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaStream_t st;
    if ((rc = lib_stream(&st))) return rc;
    if ((rc = g_scratch[S_GX].ensure(ncol * sizeof(double))) || (rc = g_scratch[S_GY].ensure(ncol * sizeof(double))) ||
        (rc = g_scratch[S_OUT].ensure((nvox ? nvox : 1) * sizeof(double))) ||
        (rc = g_scratch[S_TMP0].ensure((size_t)(nz ? nz : 1) * sizeof(double))) ||
        (rc = g_scratch[S_AUX3].ensure(3 * ncol * sizeof(double))) || (rc = g_scratch[S_AUX2].ensure(ncol)) ||
        (rc = g_scratch[S_TMP2].ensure(sizeof(int))))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GX].ptr, grid_x, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g_scratch[S_GY].ptr, grid_y, ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nz) CUDA_TRY(cudaMemcpyAsync(g_scratch[S_TMP0].ptr, level_heights, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(g_scratch[S_TMP2].ptr, 0, sizeof(int), st));
    if (nz) {
        g_nonfinite_flag = (int*)g_scratch[S_TMP2].ptr;
        rc = enqueue_cappi(vol, radar_height, effective_earth_radius, (const double*)g_scratch[S_GX].ptr,
                           (const double*)g_scratch[S_GY].ptr, (long long)ncol, ny, level_heights, nz, fill, blind_code,
                           1.0, 0.0, 0.0, 0, 0, (double*)g_scratch[S_OUT].ptr, nullptr, st);
        g_nonfinite_flag = nullptr;
        if (rc) return rc;
    }
    double* d3 = (double*)g_scratch[S_AUX3].ptr;
    rc = enqueue_columns((const double*)g_scratch[S_OUT].ptr, (const double*)g_scratch[S_TMP0].ptr, nz, (long long)ncol, fill,
                         min_dbz, max_dbz_cap, threshold_dbz, cr ? d3 : nullptr, vil ? d3 + ncol : nullptr,
                         et ? d3 + 2 * ncol : nullptr, et_topped ? (unsigned char*)g_scratch[S_AUX2].ptr : nullptr, st);
    if (rc) return rc;
    if (cr) CUDA_TRY(cudaMemcpyAsync(cr, d3, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (vil) CUDA_TRY(cudaMemcpyAsync(vil, d3 + ncol, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (et) CUDA_TRY(cudaMemcpyAsync(et, d3 + 2 * ncol, ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (et_topped) CUDA_TRY(cudaMemcpyAsync(et_topped, g_scratch[S_AUX2].ptr, ncol, cudaMemcpyDeviceToHost, st));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, g_scratch[S_TMP2].ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return flag ? PCG_WARN_NONFINITE : PCG_OK;
}
