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

namespace od {

enum { PROJ_EQUIT = 0, PROJ_OBLIQ = 1, PROJ_N_POLE = 2, PROJ_S_POLE = 3 };

struct ProjStere {
    int mode;
    double a, ra, akm1, sinX1, cosX1, phi0, lam0, x0, y0;
};

constexpr double kPi = 3.14159265358979323846;
constexpr double kHalfPi = 1.57079632679489661923;

// PROJ adjlon: reduce to [-pi, pi]; values already inside are left untouched
OD_HD double adjlon(double lam) {
    if (fabs(lam) <= kPi) return lam;
    double t = lam + kPi;
    t = t - 2.0 * kPi * floor(t / (2.0 * kPi));
    return t - kPi;
}

// lon, lat in degrees -> x, y in metres; returns false where the projection is undefined (antipode)
OD_HD bool stere_forward(const ProjStere& P, double lon, double lat, double& x, double& y) {
    const double lam = adjlon(lon * kDeg - P.lam0);
    double phi = lat * kDeg;
    double sinphi, cosphi, sinlam, coslam;
    sincos(phi, &sinphi, &cosphi);
    sincos(lam, &sinlam, &coslam);
    double px, py;
    if (P.mode == PROJ_EQUIT || P.mode == PROJ_OBLIQ) {
        const double d = P.mode == PROJ_EQUIT ? 1.0 + cosphi * coslam : 1.0 + P.sinX1 * sinphi + P.cosX1 * cosphi * coslam;
        if (!(d > 1e-10)) return false;
        const double k = P.akm1 / d;
        px = k * cosphi * sinlam;
        py = P.mode == PROJ_EQUIT ? k * sinphi : k * (P.cosX1 * sinphi - P.sinX1 * cosphi * coslam);
    } else {
        if (P.mode == PROJ_N_POLE) {
            coslam = -coslam;
            phi = -phi;
        }
        if (fabs(phi - kHalfPi) < 1e-8) return false;
        py = P.akm1 * tan(kPio4 + 0.5 * phi);
        px = sinlam * py;
        py = py * coslam;
    }
    x = P.a * px + P.x0;
    y = P.a * py + P.y0;
    return true;
}

// x, y in metres -> lon, lat in degrees
OD_HD void stere_inverse(const ProjStere& P, double x, double y, double& lon, double& lat) {
    x = (x - P.x0) * P.ra;
    y = (y - P.y0) * P.ra;
    const double rh = hypot(x, y);
    const double c = 2.0 * atan(rh / P.akm1);
    double sinc, cosc;
    sincos(c, &sinc, &cosc);
    const bool small = fabs(rh) <= 1e-10;
    double phi, lam = 0.0;
    if (P.mode == PROJ_EQUIT) {
        phi = small ? 0.0 : asin(fmin(1.0, fmax(-1.0, y * sinc / rh)));
        if (cosc != 0.0 || x != 0.0) lam = atan2(x * sinc, cosc * rh);
    } else if (P.mode == PROJ_OBLIQ) {
        phi = small ? P.phi0 : asin(fmin(1.0, fmax(-1.0, cosc * P.sinX1 + y * sinc * P.cosX1 / rh)));
        const double cc = cosc - P.sinX1 * sin(phi);
        if (cc != 0.0 || x != 0.0) lam = atan2(x * sinc * P.cosX1, cc * rh);
    } else {
        if (P.mode == PROJ_N_POLE) y = -y;
        phi = small ? P.phi0 : asin(P.mode == PROJ_S_POLE ? -cosc : cosc);
        lam = (x == 0.0 && y == 0.0) ? 0.0 : atan2(x, y);
    }
    lon = adjlon(lam + P.lam0) * kRad2Deg;
    lat = phi * kRad2Deg;
}

OD_HD double wrap180(double d) { return d - 360.0 * rint(d * (1.0 / 360.0)); }

// Forward azimuth (radians) at point 1 of the short WGS84 geodesic to point 2 (what rotate_vectors takes from Geod.inv)
OD_HD double inverse_azimuth_short(double lon1, double lat1, double lon2, double lat2) {
    double sm, cm;
    sincos(0.5 * (lat1 + lat2) * kDeg, &sm, &cm);
    const double w2 = 1.0 - Wgs84::e2 * sm * sm;
    const double w = sqrt(w2);
    const double M = Wgs84::a * (1.0 - Wgs84::e2) / (w2 * w);        // meridional radius of curvature
    const double Nc = Wgs84::a / w * cm;                              // radius of the parallel
    const double dlon = wrap180(lon2 - lon1);
    double north = M * (lat2 - lat1) * kDeg;
    double east = Nc * dlon * kDeg;
    // azimuth at point 1 = azimuth at the mid-point minus half the meridian convergence
    const double az = atan2(east, north) - 0.5 * dlon * kDeg * sm;
    const double s = hypot(east, north);
    double sa, ca;
    sincos(az, &sa, &ca);
    north = s * ca;
    east = s * sa;
    // one correction with the miss of the direct solution
    const SeriesStart ss = series_start(lat1);
    double lo, la;
    geod_move_ne(ss, lon1, north, east, lo, la);
    north += M * (lat2 - la) * kDeg;
    east += Nc * wrap180(lon2 - lo) * kDeg;
    return atan2(east, north);
}

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
    ProjStere& P = R->proj;
    P.a = r->proj.a;
    P.ra = 1.0 / r->proj.a;
    P.phi0 = r->proj.lat_0 * kDeg;
    P.lam0 = r->proj.lon_0 * kDeg;
    P.x0 = r->proj.x_0;
    P.y0 = r->proj.y_0;
    const double t = fabs(P.phi0);
    if (fabs(t - kHalfPi) < 1e-10) P.mode = P.phi0 < 0 ? PROJ_S_POLE : PROJ_N_POLE;
    else P.mode = t > 1e-10 ? PROJ_OBLIQ : PROJ_EQUIT;
    P.sinX1 = sin(P.phi0);
    P.cosX1 = cos(P.phi0);
    const double phits = fabs(r->proj.has_lat_ts ? r->proj.lat_ts * kDeg : kHalfPi);
    if (P.mode == PROJ_OBLIQ || P.mode == PROJ_EQUIT) P.akm1 = 2.0 * r->proj.k_0;
    else P.akm1 = fabs(phits - kHalfPi) >= 1e-10 ? cos(phits) / tan(kPio4 - 0.5 * phits) : 2.0 * r->proj.k_0;
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
