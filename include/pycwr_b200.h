/*
 * pycwr_b200.h — C ABI of the B200-native polar-volume -> Cartesian gridding path.
 *
 * This library replaces the compiled module `pycwr.core.RadarGridC` of the reference
 * (YvZheng/pycwr, `pycwr/core/RadarGridC.pyx`) behind the import seam at
 * `pycwr/core/__init__.py:10-54`, `pycwr/core/NRadar.py:17-20` and
 * `pycwr/interp/RadarInterp.py:39-42,1050-1056`.  The Python module
 * `pycwr_b200/RadarGridC.py` binds these entry points with ctypes and exports the reference's
 * eleven names with unchanged signatures (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / numpy types.
 *  - every function returns 0 on success or a negative pcg_status; pcg_last_error() gives the
 *    message of the calling thread's last failure.  There is NO CPU fallback: when no CUDA
 *    device is usable every compute entry point fails with PCG_ERR_CUDA.
 *  - `*_dev` entry points take DEVICE pointers and a cudaStream_t (passed as void*) and only
 *    enqueue work; the others take HOST pointers, copy in, launch, copy out and synchronise.
 *  - grids are row-major (nx, ny) float64 in metres, exactly what the reference passes as
 *    GridX/GridY; outputs are row-major float64 with `fill` marking "no data".
 *  - blind_code: 0 mask, 1 nearest_gate, 2 lowest_sweep, 3 hybrid, 4 nearest_valid
 *    (`RadarGridC.pyx:54-69`).
 */
#ifndef PYCWR_B200_H
#define PYCWR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PCG_WARN_NONFINITE = 1, /* pcg_get_cappi3d / pcg_get_cr on a packed (PCG_VALUE_F32) volume: the result
                               is complete, but a target coordinate was NaN / inf; the reference's
                               NaN-by-NaN behaviour is only restated by the PCG_VALUE_F64 layout    */
    PCG_OK = 0,
    PCG_ERR_ARG = -1,      /* bad argument (maps to the reference's ValueError)      */
    PCG_ERR_CUDA = -2,     /* CUDA runtime failure / no device                       */
    PCG_ERR_NOMEM = -3     /* host or device allocation failure                      */
} pcg_status;

/* index-record layout written by the *_index entry points (int32 x PCG_IDX_STRIDE per voxel):
 * [0] state (0 skip_beam 1 skip_low 2 skip_high 3 single 4 pair 5 empty volume)
 * [1] lower sweep [2] upper sweep, [3..5] iaz, ir, flags of the lower-sweep sample,
 * [6..8] the same for the upper-sweep sample, [9] reserved (-1).
 * flags: 1 azimuth wrapped, 2 range clamped to first gate, 4 beyond last gate, 8 before first
 * gate, 16 degenerate sweep (naz<2 or ngate<2). */
#define PCG_IDX_STRIDE 10

/* value storage of a packed volume on the device */
#define PCG_VALUE_F64 0   /* moments kept as float64; kernels restate the reference operation by operation
                             (values bit-compatible with the reference wherever az/r/el are)          */
#define PCG_VALUE_F32 1   /* production layout: float32 {lo,hi} radial pairs with the azimuth fill
                             substitution resolved at pack time; indices/masks still bit-exact, values
                             within a few float32 ulps (<< 1e-3 dB).  Falls back to PCG_VALUE_F64 for
                             volumes it cannot serve (fewer than 2 sweeps, unsorted/duplicate tables):
                             pcg_volume_info reports the layout actually used.                       */

typedef struct pcg_volume pcg_volume;   /* opaque: one radar volume resident in HBM */

/* ---- runtime ---------------------------------------------------------------------------- */
const char *pcg_last_error(void);
const char *pcg_version(void);
int pcg_device_count(int *count);
int pcg_set_device(int device);
/* pinned host memory for staging buffers (what the Python shim hands back as numpy arrays) */
int pcg_host_alloc(size_t bytes, void **ptr);
int pcg_host_free(void *ptr);
/* kernels launched by this library since load (the bench's `gpu_launches` claim) */
int64_t pcg_launch_count(void);

/* ---- volume packing: replaces the list-of-arrays argument convention of the reference
 *      (`vol_azimuth, vol_range, fix_elevation, vol_value`, produced by PRD.get_vol_data,
 *      `pycwr/core/NRadar.py:1931-1961`).  Inputs are borrowed host arrays: az[s] has naz[s]
 *      sorted float64 azimuths, rng[s] ng[s] ascending ranges, val[f][s] naz[s]*ng[s] row-major
 *      (radial, gate) float64 moments with missing == `fill` (the value later passed as
 *      `fillvalue` to the gridding calls; a mismatch there is PCG_ERR_ARG).  nfield >= 1
 *      moments share the azimuth/range tables (one set of indices serves all of them). */
int pcg_volume_create(int ne, const int *naz, const int *ng, const double *const *az,
                      const double *const *rng, int nfield, const double *const *const *val,
                      const double *fix_elev, double fill, int value_dtype, pcg_volume **out);

/* The packer itself on the device (PRD.get_vol_data / _extract_sweep_volume,
 * `pycwr/core/NRadar.py:665-712,1931-1961`): same result as pcg_volume_create, but the inputs are the sweeps
 * as the file reader left them — az[s] and the rows of val[f][s] in SCAN order, NaN for missing.  The host
 * supplies ray_order[s] = argsort(az[s]) (naz[s] ints, `:705-707`) and val_ld[s] = elements between
 * consecutive rays of val[f][s] (>= ng[s]; NULL = ng[s]): ng[s] may be smaller than the stored ray length,
 * which is the max_range_km clip of `:655-663` (a prefix of every ray).  The row gather, the clip and the
 * NaN -> fill substitution (`:1956`) run inside the pack kernels; sweeps are passed in fixed-angle order. */
int pcg_volume_create_raw(int ne, const int *naz, const int *ng, const double *const *az,
                          const double *const *rng, int nfield, const double *const *const *val,
                          const long long *val_ld, const int *const *ray_order, const double *fix_elev,
                          double fill, int value_dtype, pcg_volume **out);
int pcg_volume_destroy(pcg_volume *vol);
int pcg_volume_info(const pcg_volume *vol, int *ne, int *nfield, int *value_dtype,
                    size_t *device_bytes);

/* ---- host-buffer entry points (drop-in for the reference functions) ------------------- */

/* get_CAPPI_3d (`RadarGridC.pyx:485-521`) / get_CAPPI_xy (`:356-477`, nz == 1 with its
 * beam_width_deg; get_CAPPI_3d callers pass 1.0 because the reference does not forward it).
 * out: [nfield][nz][nx*ny].  index_out (optional, may be NULL): [nz][nx*ny][PCG_IDX_STRIDE]. */
int pcg_get_cappi3d(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                    const double *grid_x, const double *grid_y, int nx, int ny,
                    const double *level_heights, int nz, double fill, int blind_code,
                    double beam_width_deg, double *out, int32_t *index_out);

/* get_CR_xy (`RadarGridC.pyx:596-624`): max over sweeps of ppi_to_grid planes at the fixed
 * elevations, blind "mask".  Uses field 0.  out: [nx*ny]. */
int pcg_get_cr(const pcg_volume *vol, double radar_height, double effective_earth_radius,
               const double *grid_x, const double *grid_y, int nx, int ny, double fill,
               double *out);

/* Same calls with the missing-value marker of the RESULT chosen by the caller: cells the gridding leaves at
 * `fill` are stored as `out_missing` (typically NaN) before the result is copied back — the
 * `np.where(GridV == fillvalue, np.nan, GridV)` pass of the product wrappers (`pycwr/core/NRadar.py:1339,
 * 1374-1392,1420-1437`) without a host sweep over the volume.  out_missing == fill is the plain call. */
int pcg_get_cappi3d_ex(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                       const double *grid_x, const double *grid_y, int nx, int ny,
                       const double *level_heights, int nz, double fill, int blind_code,
                       double beam_width_deg, double out_missing, double *out, int32_t *index_out);
int pcg_get_cr_ex(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                  const double *grid_x, const double *grid_y, int nx, int ny, double fill, double out_missing,
                  double *out);

/* ppi_to_grid (`RadarGridC.pyx:303-352`): sweep `sweep` of the packed volume gridded at
 * elevation `elevation_deg`.  out: [nx*ny]; index_out optional [nx*ny][PCG_IDX_STRIDE]. */
int pcg_ppi_to_grid(const pcg_volume *vol, int sweep, double elevation_deg, double radar_height,
                    double effective_earth_radius, const double *grid_x, const double *grid_y,
                    int nx, int ny, double fill, int blind_code, double *out, int32_t *index_out);

/* get_mosaic_CAPPI_3d (`RadarGridC.pyx:525-593`): per radar local grid = grid - (x, y),
 * get_CAPPI_3d with blind "mask", fill-aware running max in radar order.  out: [nz][nx*ny]. */
int pcg_get_mosaic_cappi3d(int nradar, const pcg_volume *const *vols, const double *radar_x,
                           const double *radar_y, const double *radar_height,
                           const double *effective_earth_radius, const double *grid_x,
                           const double *grid_y, int nx, int ny, const double *level_heights,
                           int nz, double fill, double *out);

/* ---- device-resident entry points (what bench.py times with inputs already in HBM) ---- */
int pcg_get_cappi3d_dev(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                        const double *d_grid_x, const double *d_grid_y, int nx, int ny,
                        const double *level_heights /* host */, int nz, double fill,
                        int blind_code, double beam_width_deg, double radar_x, double radar_y,
                        int merge_max, double *d_out, int32_t *d_index_out, void *stream);
int pcg_get_cr_dev(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                   const double *d_grid_x, const double *d_grid_y, int nx, int ny, double fill,
                   double radar_x, double radar_y, double *d_out, void *stream);

/* ---- network compositing: the multi-radar merge of run_radar_network_3d, one fused pass per
 *      field (`pycwr/interp/RadarInterp.py`): per radar grid_single_radar_to_latlon_3d (:963-1023,
 *      get_CAPPI_3d on grid - (radar_x, radar_y) with `blind_code`, then the range-sanity filter),
 *      _compute_quality_weight_volume (:603-696), _compute_observability_mask (:540-600),
 *      _build_composite_weight_cache (:1379-1396) and compose_network_volume (:1399-1487), plus the
 *      coverage_count / blind_mask (:1938-1945) and the CR max-merge of per-radar get_CR_xy
 *      (:1026-1068, 2047-2063).  Every volume must be packed (PCG_VALUE_F32 in effect) and uses
 *      field 0.  beam_widths: NULL (1 degree everywhere) or per radar NULL / [ne] degrees
 *      (`_get_sweep_beam_widths`, :527-537).  Outputs: composite [nz][nx*ny] (fill where no data);
 *      optional valid_count int16 [nz][nx*ny] (second result of compose_network_volume),
 *      coverage_count int16 and blind_mask int8 [nz][nx*ny], cr [nx*ny] (NaN where no echo),
 *      quality [nz][nx*ny] = geometric quality weight of the LAST radar (the per-radar function
 *      when nradar == 1), tie_voxels = voxels whose observability gate was decided by evaluating
 *      the reference's NumPy-twin geometry (within 1e-12 of a gate or inside the elevation guard
 *      band).  Sharding: pass a row slab (pointer offset into GridX/GridY, smaller nx). */
#define PCG_NET_MAX 0               /* composite_method "max"              */
#define PCG_NET_NEAREST 1           /* "nearest"                           */
#define PCG_NET_EXP_WEIGHTED 2      /* "exp_weighted"                      */
#define PCG_NET_QUALITY_WEIGHTED 3  /* "quality_weighted" (the default, `pycwr/configure/config.py:26`) */
#define PCG_TERM_RANGE 1
#define PCG_TERM_BEAM 2
#define PCG_TERM_VERTICAL 4
typedef struct {
    int method;                 /* PCG_NET_*                                                     */
    int blind_code;             /* as for pcg_get_cappi3d (network default: 3 hybrid, :1759)     */
    int quality_terms;          /* PCG_TERM_* bits (`quality_weight_terms`)                      */
    int sanity_filter;          /* 1: range-sanity filter of grid_single_radar_to_latlon_3d      */
    double influence_radius_m;  /* 300000 (:1683)                                                */
    double range_decay_m;       /* 230000 (:1769-1774)                                           */
    double beam_radius_ref_m;   /* 3000                                                          */
    double vertical_gap_ref_deg;/* 0.5                                                           */
} pcg_network_params;

int pcg_network_compose(int nradar, const pcg_volume *const *vols, const double *radar_x,
                        const double *radar_y, const double *radar_height,
                        const double *effective_earth_radius, const double *const *beam_widths,
                        const pcg_network_params *params, const double *grid_x, const double *grid_y,
                        int nx, int ny, const double *level_heights, int nz, double fill,
                        double *composite, int16_t *valid_count, int16_t *coverage_count,
                        int8_t *blind_mask, double *cr, double *quality, uint64_t *tie_voxels);
/* pcg_network_compose with (a) the missing-value marker of the composite chosen by the caller (`out_missing`,
 * typically NaN: the `np.where(values == fillvalue, np.nan, values)` of build_network_dataset, `RadarInterp.py:1539`)
 * and (b) the column products of the composite computed while it is still on the device: column_params =
 * {vil_min_dbz, vil_max_dbz_cap, et_threshold_dbz} (NULL: none) -> vil, et [nx*ny] (NaN where undefined) and
 * et_topped [nx*ny] uint8 (`RadarInterp.py:2064-2119`, derive_vil / derive_et `pycwr/core/RadarProduct.py:33-85`);
 * composite may be NULL when only the column products are wanted (the product-level pass of `:1971-2000`). */
int pcg_network_compose_ex(int nradar, const pcg_volume *const *vols, const double *radar_x,
                           const double *radar_y, const double *radar_height,
                           const double *effective_earth_radius, const double *const *beam_widths,
                           const pcg_network_params *params, const double *grid_x, const double *grid_y,
                           int nx, int ny, const double *level_heights, int nz, double fill, double out_missing,
                           double *composite, int16_t *valid_count, int16_t *coverage_count,
                           int8_t *blind_mask, double *cr, double *quality, uint64_t *tie_voxels,
                           const double *column_params, double *vil, double *et, unsigned char *et_topped);
/* pcg_network_compose_ex for rows [row_offset, row_offset + nrows) of a grid of full_rows rows: grid_x / grid_y and
 * every output are the FULL host arrays ([full_rows][ny] planes, [nz][full_rows][ny] volumes); only the slab's rows
 * are uploaded, composited and written.  One host thread per GPU (pcg_set_device is per thread; all library state is
 * per device) calls it for its slab with the radars that reach the slab, so the GPUs of a box fill one product over
 * their own PCIe links — the device form of the reference's fork pool over radars
 * (`pycwr/interp/RadarInterp.py:1816-1861`; SURVEY 8e "D2H-copying each slab from its own GPU").  The slab's rows of
 * the result are bit-identical to the same rows of the unsharded call. */
int pcg_network_compose_rows(int nradar, const pcg_volume *const *vols, const double *radar_x,
                             const double *radar_y, const double *radar_height,
                             const double *effective_earth_radius, const double *const *beam_widths,
                             const pcg_network_params *params, const double *grid_x, const double *grid_y,
                             long long full_rows, long long row_offset, int nrows, int ny,
                             const double *level_heights, int nz, double fill, double out_missing,
                             double *composite, int16_t *valid_count, int16_t *coverage_count,
                             int8_t *blind_mask, double *cr, uint64_t *tie_voxels,
                             const double *column_params, double *vil, double *et, unsigned char *et_topped);
int pcg_network_compose_dev(int nradar, const pcg_volume *const *vols, const double *radar_x,
                            const double *radar_y, const double *radar_height,
                            const double *effective_earth_radius, const double *const *beam_widths,
                            const pcg_network_params *params, const double *d_grid_x,
                            const double *d_grid_y, int nx, int ny, const double *level_heights /* host */,
                            int nz, double fill, double *d_composite, int16_t *d_valid_count,
                            int16_t *d_coverage_count, int8_t *d_blind_mask, double *d_cr,
                            double *d_quality, void *stream);
/* Multi-GPU network composite (row slabs of the output grid, one process per GPU).  The finished slab
 * of every rank has to reach every rank (the all-gather that follows the merge); here the finalize kernel itself
 * stores the slab into the FULL composite grid of each peer through NVLink peer mappings, so producing the slab
 * and gathering it are one kernel.  pcg_peer_alloc: device buffer + its 64-byte CUDA IPC handle (exchange the
 * handles with any host-side collective); pcg_peer_open / pcg_peer_close: map / unmap a peer's buffer.
 * pcg_network_compose_gather_dev: pcg_network_compose_dev for rows [row_offset, row_offset + nx) of a grid of
 * full_rows rows whose composite is also written to peer_full[0..npeer) (each [nz][full_rows][ny]; list the own
 * buffer too); d_composite (the slab alone) may be NULL.  The caller synchronises the ranks afterwards. */
int pcg_peer_alloc(size_t bytes, void **dptr, unsigned char *handle64);
int pcg_peer_open(const unsigned char *handle64, void **dptr);
int pcg_peer_close(void *dptr);
int pcg_peer_free(void *dptr);
int pcg_network_compose_gather_dev(int nradar, const pcg_volume *const *vols, const double *radar_x,
                                   const double *radar_y, const double *radar_height,
                                   const double *effective_earth_radius, const double *const *beam_widths,
                                   const pcg_network_params *params, const double *d_grid_x,
                                   const double *d_grid_y, int nx, int ny, const double *level_heights, int nz,
                                   double fill, double *d_composite, int16_t *d_valid_count,
                                   int16_t *d_coverage_count, int8_t *d_blind_mask, double *d_cr,
                                   int npeer, double *const *peer_full, long long full_rows,
                                   long long row_offset, void *stream);
/* min / max / count of the valid (finite, != fill) source values of one moment: the window of the
 * range-sanity filter (`pycwr/interp/RadarInterp.py:1003-1007`), reduced on the device at pack time */
int pcg_volume_value_range(const pcg_volume *vol, int field, double *vmin, double *vmax, uint64_t *count);

/* _compute_observability_mask (`pycwr/interp/RadarInterp.py:540-600`) and
 * _compute_quality_weight_volume (:603-696) as stand-alone functions for any sweep count (the
 * NumPy-twin geometry of `pycwr/core/transforms.py:201-222` restated in FP64 for every voxel).
 * first_gate / last_gate: vol_range[s][0] / vol_range[s][-1]; beam_widths [ne] degrees or NULL (1.0).
 * mask: uint8 [nz][nx*ny] or NULL; quality: float64 [nz][nx*ny] or NULL. */
int pcg_network_gates(int ne, const double *fix_elev, const double *first_gate, const double *last_gate,
                      const double *beam_widths, double radar_height, double effective_earth_radius,
                      double radar_x, double radar_y, const double *grid_x, const double *grid_y, int nx,
                      int ny, const double *level_heights, int nz, int quality_terms,
                      double range_decay_m, double beam_radius_ref_m, double vertical_gap_ref_deg,
                      unsigned char *mask, double *quality);

/* compose_network_volume (`pycwr/interp/RadarInterp.py:1399-1487`) on already gridded per-radar
 * volumes (host arrays [nz][nx*ny] each), restated in FP64 with the radar axis accumulated in order.
 * quality: per-radar weight volumes or NULL; weight2d: [nradar][nx*ny] the 2-D cache of
 * _build_composite_weight_cache (:1379-1396; squared distances for "nearest"), NULL for "max". */
int pcg_compose_volumes(int nradar, const double *const *volumes, const double *const *quality,
                        const double *weight2d, int method, int nz, int nx, int ny, double fill,
                        double *composite, int16_t *valid_count);

/* derive_cr / derive_vil / derive_et (`pycwr/core/RadarProduct.py:26-85`) over a gridded volume
 * [nz][nx*ny]; every output is optional (NULL skips it): cr, vil, et [nx*ny] float64 (NaN where
 * undefined), topped [nx*ny] uint8. */
int pcg_column_products(const double *volume, const double *level_heights, int nz, int nx, int ny,
                        double fill, double min_dbz, double max_dbz_cap, double threshold_dbz,
                        double *cr, double *vil, double *et, unsigned char *topped);
/* Grid one radar's volume (pcg_get_cappi3d, one moment) and reduce its columns in the same call: only the
 * (nx, ny) planes come back — the flow of PRD.add_product_VIL_xy / add_product_ET_xy
 * (`pycwr/core/NRadar.py:1459-1572`: _grid_reflectivity_volume_xy then derive_vil / derive_et) without the
 * volume's round trip over PCIe.  Any of cr / vil / et / et_topped may be NULL. */
int pcg_get_column_products(const pcg_volume *vol, double radar_height, double effective_earth_radius,
                            const double *grid_x, const double *grid_y, int nx, int ny,
                            const double *level_heights, int nz, double fill, int blind_code,
                            double min_dbz, double max_dbz_cap, double threshold_dbz, double *cr,
                            double *vil, double *et, unsigned char *et_topped);
int pcg_column_products_dev(const double *d_volume, const double *d_levels, int nz, int nx, int ny,
                            double fill, double min_dbz, double max_dbz_cap, double threshold_dbz,
                            double *d_cr, double *d_vil, double *d_et, unsigned char *d_topped,
                            void *stream);

/* ---- local VVP wind retrieval on one sweep -----------------------------
 * pcg_vvp_local: _run_local_vvp (`pycwr/retrieve/WindField.py:1843-1941`): for every (ray, gate) a
 *   moving window of az_num rays (wrapping) x bin_num gates (both odd), its valid-sample statistics
 *   (`:58-90`), and — where the window passes the thresholds — the weighted least-squares fit of
 *   (u, v, radial offset) with Gaussian window weights (`:1832-1840`) and up to five Huber re-weightings
 *   (`_solve_weighted_least_squares`, `:115-200`).  velocity: [naz][nrange] float64, NaN = missing (the
 *   caller has already mapped its fill value to NaN, `:50-55`); azimuth, elevation: [naz] degrees.
 *   out_f64: [9][naz][nrange] in the order u, v, wind_speed, wind_direction, radial_offset, fit_rmse,
 *   condition_number, valid_fraction, azimuth_coverage_deg (NaN where undefined); out_i32:
 *   [2][naz][nrange] azimuth_sector_count, valid_count. */
int pcg_vvp_local(const double *azimuth, const double *elevation, const double *velocity, int naz, int nrange,
                  int az_num, int bin_num, double min_valid_fraction, int min_valid_count,
                  double min_azimuth_coverage_deg, int min_azimuth_sector_count, double max_condition_number,
                  double *out_f64, int32_t *out_i32);

/* ---- vertical sections and pseudo-RHI cuts -------------------------------
 * pcg_section_sample: for every sweep of a float64 volume (value_dtype = PCG_VALUE_F64; missing moments
 *   NaN or the volume's fill value) sample `npts` line points.
 *   Geometry mode (x, y != NULL): the per-sweep loop of PRD.extract_section (`pycwr/core/NRadar.py:1048-1060`):
 *   cartesian_xy_elevation_to_range_z (`pycwr/core/transforms.py:224-250`) at the sweep's fixed elevation
 *   gives (azimuth, slant range, beam height), written to out_az / out_range / out_z [ne][npts], and the
 *   moment is interpolated there.  Explicit mode (x == NULL): targets come from target_az / target_range,
 *   [ne][npts] when per_sweep_targets != 0 (the RHI cut of PRD._extract_ppi_rhi, `NRadar.py:1144-1155`,
 *   rows padded with NaN) else one [npts] row used for every sweep; out_az / out_range / out_z may be NULL.
 *   Interpolation: PRD._interpolate_ppi_strip_arrays (`NRadar.py:833-913`); azimuth tables must already be
 *   de-duplicated and reduced mod 360 (`:845-850`, host side).  out_value [ne][npts], NaN = no data. */
int pcg_section_sample(const pcg_volume *vol, int field, int npts, const double *x, const double *y,
                       double radar_height, double effective_earth_radius, const double *target_az,
                       const double *target_range, int per_sweep_targets, double *out_value,
                       double *out_az, double *out_range, double *out_z);

/* ---- the gridding step of retrieve_wind_volume_xy (`pycwr/retrieve/WindField.py`) ----------
 * pcg_wind_gather: _grid_vvp_retrieval_xy (:629-689) with _aggregate_neighbors (:597-626) for one sweep.
 *   samples: [10][ns] float64 structure of arrays in the order x, y, z, u, v, fit_rmse, valid_count,
 *   valid_fraction, condition_number, azimuth_sector_count (already reduced to the finite samples,
 *   _flatten_valid_vvp_samples :561-594); targets [nt]; max_radius_m <= radius_m disables the expansion.
 *   out_f64: [5][nt] u, v, z, fit_rmse, condition_number (NaN where undefined); out_i32: [3][nt]
 *   valid_count, neighbor_count, azimuth_sector_count; expanded: [nt] uint8.
 * pcg_wind_vertical: _vertical_profile_from_sweeps (:692-810) + _compute_voxel_quality (:813-851) for
 *   every column (_reconstruct_wind_volume_chunk :1183-1229).  sweeps: [7][nsweep][ncol] float64 in the
 *   order z, u, v, fit_rmse, valid_count, neighbor_count, expanded_radius; out_f64: [5][nz][ncol] u, v,
 *   fit_rmse, sweep_spread_m, quality_score; out_i32: [4][nz][ncol] valid_count, source_sweep_count,
 *   neighbor_count, effective_vertical_support; quality_flag: [nz][ncol] uint8. */
int pcg_wind_gather(const double *samples, long long ns, const double *target_x, const double *target_y,
                    long long nt, double radius_m, double max_radius_m, int min_neighbors, double power,
                    double *out_f64, int32_t *out_i32, unsigned char *expanded);
int pcg_wind_vertical(const double *sweeps, int nsweep, long long ncol, const double *level_heights, int nz,
                      double vertical_tolerance_m, double max_vertical_gap_m, double *out_f64,
                      int32_t *out_i32, unsigned char *quality_flag);

/* ---- diagnostics: the device's geometry inversion for a list of points (one level), used by
 *      the tests to measure bit-agreement of the device math with glibc. ------------------- */
int pcg_debug_cartesian_to_antenna(const double *x, const double *y, size_t n, double z, double h,
                                   double effective_earth_radius, double *az, double *r,
                                   double *el);
/* the kernels' fused sqrt + reciprocal-root against the IEEE square root: number of inputs whose
 * root differs (must be 0) and the worst |y*sqrt(a) - 1| of the reciprocal by-product */
int pcg_debug_sqrt_check(const double *a, size_t n, uint64_t *mismatches, double *worst_rinv_err);
int pcg_debug_xye_to_antenna(const double *x, const double *y, size_t n, double elevation_deg,
                             double h, double effective_earth_radius, double *az, double *r);

#ifdef __cplusplus
}
#endif
#endif /* PYCWR_B200_H */
