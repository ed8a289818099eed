// od_bookkeep.cuh -- per-element housekeeping of the run loop and of OceanDrift.update() that is not a move.
//
//   buoyancy_one      OceanDrift.vertical_buoyancy (opendrift/models/oceandrift.py:352-357):
//                       z[z < 0] = np.minimum(0, z + terminal_velocity * dt)
//                     and the 'lift_to_seafloor' / 'deactivate' branches of OpenDriftSimulation.interact_with_seafloor
//                     (opendrift/models/basemodel/__init__.py:748-783) that the same method reaches through
//                     bottom_interaction (:359-367): z < -(sea_floor_depth + sea_surface_height) -> z = -(...).
//   bookkeep_one      one pass over the active elements for what the run loop does between get_environment and update()
//                     (basemodel/__init__.py:2249-2270):
//                       deactivate_outside       (:2358-2386)  lon / lat against the validity domain -> status 'outside'
//                       state_to_buffer          (:2384-2403)  lon / lat / z / status of the element into column `col`
//                                                               of the [trajectory, time] output block (optional)
//                       increase_age_and_retire  (:2345-2356)  age_seconds += dt; age >= max_age -> status 'retired'
//                     deactivate_elements (:1774-1795) semantics: the status of an element that is already deactivated is
//                     kept, `moving` becomes 0.  The reference numbers a status category when it first occurs
//                     (status_categories.append at the first deactivate_elements call that selects an element), so the
//                     kernel writes the codes the host hands it -- provisional ones until the category has a number --
//                     and counts, per launch, the elements it put into 'outside' / 'retired' and those whose status is
//                     non-zero after the pass (one atomic per warp and counter): the host numbers new categories in the
//                     reference's order and skips the compaction when nothing left.
//   coast_one         OpenDriftSimulation.interact_with_coastline (basemodel/__init__.py:671-746) with
//                     general:coastline_approximation_precision = None (the bisection towards the coastline, coastline_crossing
//                     :81-134, queries the GSHHG landmask of the roaring_landmask package: IO-backed, not on this path):
//                       'stranding': deactivate_elements((land_binary_mask == 1) & (z <= 0), 'stranded')
//                       'previous':  elements of age 0 on land -> 'seeded_on_land' (while elements are being released), then every
//                                    element on land goes back to its position of the previous step
//                     and the 'previous' branch of interact_with_seafloor (:775-783): elements below the sea floor go back likewise
//                     and, for both, elements the mask reader does not cover -> 'missing_data' (report_missing_variables,
//                     :2501-2515: land_binary_mask has no fallback value).
//                     The previous positions are float32: the reference keeps them in a copy of its float32 result block
//                     (:2164-2165, default_dtype :2094), so an element that is moved back lands on float32-rounded coordinates.
//   store_previous    update_previous_state (:642-669) for lon / lat: previous[ID] = present position of every active element.
// NumPy dtype rules are kept: the dtypes of z / terminal_velocity / age_seconds are whatever the reference's arrays have
// (float32 as seeded arrays, float64 once a scalar property was broadcast on release, elements/elements.py:213-216).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define OD_BK_HD __host__ __device__ __forceinline__
#else
#define OD_BK_HD static inline
#endif
// NumPy rounds every product and every sum: no fused multiply-add (nvcc contracts a * b + c by default)
#if defined(__CUDA_ARCH__)
#define OD_BK_MUL_F(a, b) __fmul_rn((a), (b))
#define OD_BK_ADD_F(a, b) __fadd_rn((a), (b))
#define OD_BK_MUL_D(a, b) __dmul_rn((a), (b))
#define OD_BK_ADD_D(a, b) __dadd_rn((a), (b))
#else
#define OD_BK_MUL_F(a, b) ((a) * (b))
#define OD_BK_ADD_F(a, b) ((a) + (b))
#define OD_BK_MUL_D(a, b) ((a) * (b))
#define OD_BK_ADD_D(a, b) ((a) + (b))
#endif

namespace od {

struct BuoyancyParams {
    int64_t n;
    const void* z_in;            // float32 / float64 (z_f64)
    void* z_out;                 // same dtype as z_in; may alias it
    const void* tv;              // terminal_velocity float32 / float64 (tv_f64); NULL = no buoyancy move (sea floor only)
    const float* sea_floor;      // sea_floor_depth_below_sea_level at the elements, or NULL
    int32_t* status;             // with seafloor_action 'deactivate': status / moving of the elements that hit the floor
    int32_t* moving;
    unsigned* counter;           // += elements deactivated here
    double dt;
    float ssh;                   // sea_surface_height (its fallback: a reader for it is refused at run())
    int32_t z_f64, tv_f64;
    int32_t seafloor_code;       // status code of 'seafloor' when seafloor_action == 'deactivate', else 0
};

OD_BK_HD bool buoyancy_one(const BuoyancyParams& p, int64_t i) {
    bool deact = false;
    if (p.z_f64) {
        double z = ((const double*)p.z_in)[i];
        if (p.tv && z < 0.0) {
            // float32 * Python float stays float32 (weak scalar); the sum with a float64 z is float64
            // (the product is rounded on its own -- float32 for a float32 terminal velocity -- before the float64 sum)
            const double d = p.tv_f64 ? OD_BK_MUL_D(((const double*)p.tv)[i], p.dt) : (double)OD_BK_MUL_F(((const float*)p.tv)[i], (float)p.dt);
            z = fmin(0.0, OD_BK_ADD_D(z, d));
        }
        if (p.sea_floor) {
            const float zmin = -(p.sea_floor[i] + p.ssh);       // float32 environment arithmetic
            if (z < (double)zmin) {
                z = (double)zmin;
                deact = p.seafloor_code != 0;
            }
        }
        ((double*)p.z_out)[i] = z;
    } else {
        float z = ((const float*)p.z_in)[i];
        if (p.tv && z < 0.0f) {
            if (p.tv_f64) z = (float)fmin(0.0, OD_BK_ADD_D((double)z, OD_BK_MUL_D(((const double*)p.tv)[i], p.dt)));    // float64 sum, stored into the float32 array
            else z = fminf(0.0f, OD_BK_ADD_F(z, OD_BK_MUL_F(((const float*)p.tv)[i], (float)p.dt)));
        }
        if (p.sea_floor) {
            const float zmin = -(p.sea_floor[i] + p.ssh);
            if (z < zmin) {
                z = zmin;
                deact = p.seafloor_code != 0;
            }
        }
        ((float*)p.z_out)[i] = z;
    }
    if (deact && p.status) {
        if (p.status[i] == 0) p.status[i] = p.seafloor_code;
        if (p.moving) p.moving[i] = 0;
    }
    return deact;
}

struct BookkeepParams {
    int64_t n;
    const double* lon;
    const double* lat;
    const void* z;               // float32 / float64 (z_f64); only read for the output block
    void* age;                   // age_seconds float32 / float64 (age_f64), updated in place
    int32_t* status;
    int32_t* moving;
    const int32_t* ids;          // element IDs (rows of the output block)
    unsigned* counters;          // [0] += newly 'outside', [1] += newly 'retired', [2] += status != 0 after the pass
    double dt_age;               // seconds added to age_seconds
    double max_age;              // drift:max_age_seconds, or NaN = no retirement
    double west, east, south, north;    // validity domain, NaN = no limit (east > 180: longitudes < 0 are compared as lon + 360, :2362-2376)
    int32_t outside_code, retired_code;
    int32_t z_f64, age_f64;
    int32_t pos_f32;             // lon / lat still carry float32 values: NumPy compares them with the limits in float32
    int32_t only_deactivated;    // not an output time: only elements with status != 0 are written, into the NEXT output column (:2390-2396, 'backfill')
    // output block [n_total][ncols], or blon == NULL when this is not an output step
    int64_t n_total;
    int32_t col, ncols;
    float* blon;
    float* blat;
    float* bz;
    int32_t* bstatus;
};

// returns bit 0: newly 'outside', bit 1: newly 'retired', bit 2: status non-zero after the pass
OD_BK_HD int bookkeep_one(const BookkeepParams& p, int64_t i) {
    int st = p.status[i];
    int flags = 0;
    bool off = false;
    const double lon = p.lon[i], lat = p.lat[i];
    // deactivate_outside: four separate deactivate_elements calls (west, east, south, north), all with the same reason
    bool out;
    if (p.pos_f32) {         // float32 array against a Python float: the scalar is cast to float32 (NumPy weak scalars)
        const float lf = (float)lon, af = (float)lat;
        const float lc = (p.east == p.east && p.east > 180.0 && lf < 0.0f) ? lf + 360.0f : lf;
        out = (p.west == p.west && lc < (float)p.west) || (p.east == p.east && lc > (float)p.east) ||
              (p.south == p.south && af < (float)p.south) || (p.north == p.north && af > (float)p.north);
    } else {
        const double lonc = (p.east == p.east && p.east > 180.0 && lon < 0.0) ? lon + 360.0 : lon;
        out = (p.west == p.west && lonc < p.west) || (p.east == p.east && lonc > p.east) ||
              (p.south == p.south && lat < p.south) || (p.north == p.north && lat > p.north);
    }
    if (out) {
        if (st == 0) { st = p.outside_code; flags |= 1; }
        off = true;
    }
    if (p.blon && (!p.only_deactivated || st != 0)) {
        const int64_t id = p.ids[i];
        if (id >= 0 && id < p.n_total) {
            const int64_t o = id * p.ncols + p.col;
            p.blon[o] = (float)lon;
            p.blat[o] = (float)lat;
            p.bz[o] = p.z_f64 ? (float)((const double*)p.z)[i] : ((const float*)p.z)[i];
            p.bstatus[o] = st;
        }
    }
    // age_seconds += time_step.total_seconds()  (in place: the array keeps its dtype)
    bool old;
    if (p.age_f64) {
        const double a = ((double*)p.age)[i] + p.dt_age;
        ((double*)p.age)[i] = a;
        old = a >= p.max_age;
    } else {
        const float a = ((float*)p.age)[i] + (float)p.dt_age;
        ((float*)p.age)[i] = a;
        old = a >= (float)p.max_age;
    }
    if (p.max_age == p.max_age && old) {
        if (st == 0) { st = p.retired_code; flags |= 2; }
        off = true;
    }
    if (off) {
        p.status[i] = st;
        p.moving[i] = 0;
    }
    return flags | (st != 0 ? 4 : 0);
}

struct CoastParams {
    int64_t n;
    const float* mask;           // land_binary_mask at the elements (float32 environment value; NaN = not covered)
    double* lon;
    double* lat;
    const void* z;               // float32 / float64 (z_f64)
    const void* age;             // age_seconds float32 / float64 (age_f64)
    int32_t* status;
    int32_t* moving;
    const int32_t* ids;
    float* prev_lon;             // [n_total] keyed by ID - id_base
    float* prev_lat;
    unsigned* counters;          // [0] += stranded, [1] += seeded_on_land, [2] += missing_data, [3] += moved back
    int64_t n_total;
    int32_t id_base;
    int32_t action;              // 1 'stranding', 2 'previous'; 3: general:seafloor_action = 'previous' (mask = sea floor depth)
    float ssh;                   // action 3: sea_surface_height (the water column is sea floor depth + sea surface height)
    int32_t stranded_code, seeded_code, missing_code;
    int32_t check_seeded;        // elements were released this step (newly_seeded_IDs is not None)
    int32_t z_f64, age_f64;
};

// returns bit 0: newly 'stranded', bit 1: newly 'seeded_on_land', bit 2: newly 'missing_data', bit 3: moved back
OD_BK_HD int coast_one(const CoastParams& p, int64_t i) {
    const float m = p.mask[i];
    int flags = 0;
    int st = p.status[i];
    bool off = false;
    if (p.action == 3) {
        // interact_with_seafloor, 'previous' (:775-783): an element below the sea floor goes back to the horizontal position of the
        // previous step; its depth stays.  -(sea_floor_depth + sea_surface_height) is float32 environment arithmetic.
        const float zmin = -(m + p.ssh);
        const bool below = p.z_f64 ? ((const double*)p.z)[i] < (double)zmin : ((const float*)p.z)[i] < zmin;
        if (below) {
            const int64_t k = (int64_t)p.ids[i] - p.id_base;
            if (k >= 0 && k < p.n_total) {
                p.lon[i] = (double)p.prev_lon[k];
                p.lat[i] = (double)p.prev_lat[k];
                flags |= 8;
            }
        }
        return flags;
    }
    if (!(m == m) || !(fabsf(m) <= 3.4028234663852886e38f)) {        // report_missing_variables comes first in the loop (:2247)
        if (p.missing_code) {
            if (st == 0) { st = p.missing_code; flags |= 4; }
            off = true;
        }
    } else if (m == 1.0f) {
        if (p.action == 1) {
            const double z = p.z ? (p.z_f64 ? ((const double*)p.z)[i] : (double)((const float*)p.z)[i]) : 0.0;
            if (z <= 0.0) {
                if (st == 0) { st = p.stranded_code; flags |= 1; }
                off = true;
            }
        } else if (p.action == 2) {
            if (p.check_seeded) {
                const bool age0 = p.age_f64 ? ((const double*)p.age)[i] == 0.0 : ((const float*)p.age)[i] == 0.0f;
                if (age0) {
                    if (st == 0) { st = p.seeded_code; flags |= 2; }
                    off = true;
                }
            }
            const int64_t k = (int64_t)p.ids[i] - p.id_base;
            if (k >= 0 && k < p.n_total) {
                p.lon[i] = (double)p.prev_lon[k];
                p.lat[i] = (double)p.prev_lat[k];
                flags |= 8;
            }
        }
    }
    if (off) {
        p.status[i] = st;
        p.moving[i] = 0;
    }
    return flags;
}

OD_BK_HD void store_previous_one(int64_t i, const double* lon, const double* lat, const int32_t* ids, int32_t id_base, int64_t n_total,
                                 float* prev_lon, float* prev_lat) {
    const int64_t k = (int64_t)ids[i] - id_base;
    if (k >= 0 && k < n_total) {
        prev_lon[k] = (float)lon[i];
        prev_lat[k] = (float)lat[i];
    }
}

}  // namespace od
