// od_analytic.cuh -- analytical (continuous) readers on a projected plane, evaluated per particle inside the step.
//
// BASELINE configs[0] (examples/example_double_gyre_advection_schemes.py) drives OceanDrift with
// opendrift/readers/reader_double_gyre.py: a closed-form current on a 2 m x 1 m box of a spherical stereographic plane.
// For every get_environment call (4 per RK4 step) the reference does, per particle
//   modulate_longitude (readers/basereader/variables.py:259-280) -> pyproj.Proj forward (lonlat2xy, :129-143) ->
//   covers_positions_xy (:229-257) -> Reader.get_variables (reader_double_gyre.py:57-82, float64) ->
//   rotate_vectors (:59-109: inverse projection of (x, y) and (x, y + 10 m), Geod.inv azimuth of that line, rotation
//   by minus the azimuth) -> NaN when uncovered (:841-853) -> float32 (environment.py:695-696) -> fallback (:782-791).
// Here that chain is one device function, and the Euler / RK2 / RK4 stage loop around it is one kernel launch per step
// (analytic_step_particle), with the same arithmetic policies as the gridded step (od_advect.cuh).
//
// Projection: spherical stereographic, the four aspects PROJ's stere.cpp distinguishes (Snyder 1987, eqs. 21-2..21-4,
// 20-14, 20-15, 20-18, 21-15), wrapped in PROJ's generic steps (lam = lon - lon_0 reduced to [-pi, pi], x = a x' + x_0).
// Geod.inv: only the forward azimuth of a short line is needed; it is obtained by inverting the direct solution
// (mid-latitude first guess, one correction with the miss of the direct series/Karney move) -- the miss shrinks by
// (s/a)^2 ~ 2e-12 per pass for the 10 m line.
#pragma once
#include <string.h>
#include "../../include/odcuda.h"
#include "od_advect.cuh"
#include "od_proj.cuh"

namespace od {

// reader_double_gyre.py:66-73
struct GyreField {
    double A, epsilon, omega;
};

OD_HD void gyre_uv(const GyreField& G, double t, double x, double y, double& u, double& v) {
    const double so = sin(G.omega * t);
    const double a = G.epsilon * so;
    const double b = 1.0 - 2.0 * G.epsilon * so;
    const double f = a * x * x + b * x;
    const double dfdx = 2.0 * a * x + b;
    double sf, cf, sy, cy;
    sincos(kPi * f, &sf, &cf);
    sincos(kPi * y, &sy, &cy);
    u = -kPi * G.A * sf * cy;
    v = kPi * G.A * cf * sy * dfdx;
}

struct AnalyticReader {
    ProjStere proj;
    GyreField gyre;
    double xmin, xmax, ymin, ymax;      // coverage in the reader's plane
    double rot_delta;                   // length of the line that defines the y-axis azimuth (10 m, variables.py:79-82)
    int lon_mode;                       // 0: np.mod(lon, 360); 1: np.mod(lon + 180, 360) - 180
    float fallback[2];                  // environment:fallback values, NaN = none
};

// od_analytic_desc (include/odcuda.h) -> the per-launch constants; the aspect and scale constant are chosen as PROJ's
// stere setup does for a sphere.  Returns 0, or 1 unknown reader kind, 2 unknown projection, 3 bad radius / scale.
static inline int analytic_from_desc(const od_analytic_desc* r, AnalyticReader* R) {
    if (r->kind != OD_ANALYTIC_DOUBLE_GYRE) return 1;
    if (r->proj.kind != OD_PROJ_STERE_SPHERE) return 2;
    if (!(r->proj.a > 0.0) || !(r->proj.k_0 > 0.0)) return 3;
    memset(R, 0, sizeof(*R));
    if (proj_from_desc(&r->proj, &R->proj)) return 3;
    R->gyre.A = r->par[0];
    R->gyre.epsilon = r->par[1];
    R->gyre.omega = r->par[2];
    R->xmin = r->xmin; R->xmax = r->xmax; R->ymin = r->ymin; R->ymax = r->ymax;
    R->rot_delta = r->rot_delta;
    R->lon_mode = r->lon_mode;
    R->fallback[0] = r->fallback[0];
    R->fallback[1] = r->fallback[1];
    return 0;
}

// The reader chain for one particle: float32 velocity, NaN where the reader does not cover the position
OD_HD void analytic_sample_raw(const AnalyticReader& R, double t, double lon, double lat, bool pos_f32, float& u, float& v) {
    double x;
    if (pos_f32) {          // first step after seeding: the element arrays are float32 and so is np.mod's result
        const float xf = (R.lon_mode == 0) ? np_mod360f((float)lon) : OD_FADD(np_mod360f(OD_FADD((float)lon, 180.0f)), -180.0f);
        x = (double)xf;
    } else {
        x = (R.lon_mode == 0) ? np_mod360(lon) : OD_DSUB(np_mod360(OD_DADD(lon, 180.0)), 180.0);
    }
    double px, py;
    u = NAN;
    v = NAN;
    if (!stere_forward(R.proj, x, lat, px, py)) return;
    if (!(px >= R.xmin && px <= R.xmax && py >= R.ymin && py <= R.ymax)) return;
    double fu, fv;
    gyre_uv(R.gyre, t, px, py, fu, fv);
    double lon1, lat1, lon2, lat2;
    stere_inverse(R.proj, px, py, lon1, lat1);
    stere_inverse(R.proj, px, py + R.rot_delta, lon2, lat2);
    const double rot = -inverse_azimuth_short(lon1, lat1, lon2, lat2);
    double sr, cr;
    sincos(rot, &sr, &cr);
    u = (float)(fu * cr - fv * sr);
    v = (float)(fu * sr + fv * cr);
}

OD_HD void analytic_sample(const AnalyticReader& R, double t, double lon, double lat, bool pos_f32, float& u, float& v) {
    analytic_sample_raw(R, t, lon, lat, pos_f32, u, v);
    if (!finite_f(u)) u = R.fallback[0];
    if (!finite_f(v)) v = R.fallback[1];
}

struct AnalyticStepParams {
    AnalyticReader R;
    double t_start, t_mid, t_end;       // seconds since the reader's initial_time: t, t + dt/2, t + dt
    double dt;
    float dt32;
    int32_t has_k1, pos_f32;
    int64_t n;
    double* lon;
    double* lat;
    const void* factor;
    const int32_t* moving;
    const float* k1u;
    const float* k1v;
    float* env_u;
    float* env_v;
};

// PhysicsMethods.advect_ocean_current (physics_methods.py:611-691) for one particle with the analytical reader:
// the same stage sequence, quirks and dtype flow as step_particle / rk_velocity in od_advect.cuh.
template <int SCHEME, bool F64, class MATH>
OD_HD void analytic_step_particle(const AnalyticStepParams& p, int64_t i) {
    const double lon0 = p.lon[i], lat0 = p.lat[i];
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const typename MATH::Start gs = MATH::start(lat0);
    float k1u, k1v;
    if (p.has_k1) {
        k1u = p.k1u[i];
        k1v = p.k1v[i];
    } else {
        analytic_sample(p.R, p.t_start, lon0, lat0, p.pos_f32 != 0, k1u, k1v);
    }
    if (p.env_u) p.env_u[i] = k1u;
    if (p.env_v) p.env_v[i] = k1v;
    float ru = k1u, rv = k1v;
    if (SCHEME != 0) {
        float ku = k1u, kv = k1v, su = k1u, sv = k1v;
        const int last = SCHEME == 1 ? 1 : 3;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int st = 1; st <= last; ++st) {
            double mlon, mlat;
            MATH::midpoint(gs, lon0, lat0, ku, kv, p.dt32, mlon, mlat);
            analytic_sample(p.R, st == 3 ? p.t_end : p.t_mid, mlon, mlat, false, ku, kv);
            if (st < 3) {
                su = OD_FADD(su, OD_FMUL(2.0f, ku));
                sv = OD_FADD(sv, OD_FMUL(2.0f, kv));
            } else {
                su = OD_FADD(su, ku);
                sv = OD_FADD(sv, kv);
            }
        }
        if (SCHEME == 1) {
            ru = ku;
            rv = kv;
        } else {
            ru = su / 6.0f;
            rv = sv / 6.0f;
        }
    }
    double lon1, lat1;
    if (F64) {
        const double f = p.factor ? ((const double*)p.factor)[i] : 1.0;
        MATH::move64(gs, lon0, lat0, OD_DMUL((double)ru, f), OD_DMUL((double)rv, f), mv, p.dt, lon1, lat1);
    } else {
        const float f = p.factor ? ((const float*)p.factor)[i] : 1.0f;
        MATH::move32(gs, lon0, lat0, OD_FMUL(ru, f), OD_FMUL(rv, f), mv, p.dt, lon1, lat1);
    }
    p.lon[i] = lon1;
    p.lat[i] = lat1;
}

}  // namespace od
