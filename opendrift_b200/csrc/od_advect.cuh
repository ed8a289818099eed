// od_advect.cuh -- one particle, one time step of ocean-current advection (Euler / RK2 / RK4).
//
// Restates PhysicsMethods.advect_ocean_current (opendrift/models/physics_methods.py:611-691) and
// OpenDriftSimulation.update_positions (opendrift/models/basemodel/__init__.py:4630-4669) for a GPU
// thread, keeping the reference's dtype flow:
//   * stage velocities are float32 (Environment.get_environment casts, environment.py:695-696);
//   * mid-point azimuth / speed / distance are float32: az = degrees(arctan2(u, v)), speed = sqrt(u*u + v*v),
//     dist = speed * dt * .5 (physics_methods.py:629-631), promoted to float64 inside the geodesic;
//   * RK4 stage 4 samples at x0 (+) 0.5*dt*k3 but at time t + dt (physics_methods.py:660-670);
//   * u4 = (k1 + 2 k2 + 2 k3 + k4) / 6 in float32 (:674-675);
//   * the final move multiplies by factor*current_drift_factor, which is float64 when the element property
//     was a scalar (LagrangianArray.move_elements promotes, elements/elements.py:213-216) and float32 when
//     the user seeded an array; update_positions then works in that dtype.
#pragma once
#include <stdint.h>
#include "od_interp.cuh"

namespace od {

// np.degrees for float32 multiplies by (float)(180.0f / NPY_PIf) = 0x42652EE0
OD_HD float rad2deg_f32() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x42652EE0);
#else
    union { unsigned u; float f; } c;
    c.u = 0x42652EE0u;
    return c.f;
#endif
}

OD_HD float az_f32(float xv, float yv) { return OD_FMUL(atan2f(xv, yv), rad2deg_f32()); }
OD_HD float speed_f32(float xv, float yv) { return sqrtf(OD_FADD(OD_FMUL(xv, xv), OD_FMUL(yv, yv))); }

// x0 (+) 0.5*dt*k with the reference's float32 azimuth / distance
OD_HD void rk_midpoint(const GeodStart& gs, double lon0, float ku, float kv, float dt32,
                       double& mlon, double& mlat) {
    const float az = az_f32(ku, kv);
    const float dist = OD_FMUL(OD_FMUL(speed_f32(ku, kv), dt32), 0.5f);
    geod_move(gs, lon0, (double)az, (double)dist, mlon, mlat);
}

// update_positions with float32 velocities (float32 azimuth and speed, float64 distance)
OD_HD void final_move_f32(const GeodStart& gs, double lon0, float xv, float yv, double moving, double dt,
                          double& lon1, double& lat1) {
    const float az = az_f32(xv, yv);
    const double vel = OD_DMUL((double)speed_f32(xv, yv), moving);
    geod_move(gs, lon0, (double)az, OD_DMUL(vel, dt), lon1, lat1);
}

// update_positions with float64 velocities
OD_HD void final_move_f64(const GeodStart& gs, double lon0, double xv, double yv, double moving, double dt,
                          double& lon1, double& lat1) {
    const double az = OD_DMUL(atan2(xv, yv), kRad2Deg);
    const double vel = OD_DMUL(sqrt(OD_DADD(OD_DMUL(xv, xv), OD_DMUL(yv, yv))), moving);
    geod_move(gs, lon0, az, OD_DMUL(vel, dt), lon1, lat1);
}

struct CurrentStages {
    GroupGeom g;
    PairRef t_start, t_mid, t_end;
};

// Returns the RK-combined velocity (float32) that the final move uses; k1 is sampled here unless given.
template <int SCHEME>
OD_HD void rk_velocity(const CurrentStages& cs, const VertW& vw, const GeodStart& gs, double lon0, double lat0,
                       float dt32, float k1u, float k1v, float& ou, float& ov) {
    if (SCHEME == 0) {
        ou = k1u;
        ov = k1v;
        return;
    }
    double mlon, mlat;
    rk_midpoint(gs, lon0, k1u, k1v, dt32, mlon, mlat);
    float k2u, k2v;
    sample2(cs.g, cs.t_mid, vw, mlon, mlat, k2u, k2v);
    if (SCHEME == 1) {
        ou = k2u;
        ov = k2v;
        return;
    }
    rk_midpoint(gs, lon0, k2u, k2v, dt32, mlon, mlat);
    float k3u, k3v;
    sample2(cs.g, cs.t_mid, vw, mlon, mlat, k3u, k3v);
    rk_midpoint(gs, lon0, k3u, k3v, dt32, mlon, mlat);     // half step (reference quirk) ...
    float k4u, k4v;
    sample2(cs.g, cs.t_end, vw, mlon, mlat, k4u, k4v);     // ... at time t + dt
    // (x_vel + 2*x_vel2 + 2*x_vel3 + x_vel4)/6.0, float32, left to right
    ou = OD_FADD(OD_FADD(OD_FADD(k1u, OD_FMUL(2.0f, k2u)), OD_FMUL(2.0f, k3u)), k4u) / 6.0f;
    ov = OD_FADD(OD_FADD(OD_FADD(k1v, OD_FMUL(2.0f, k2v)), OD_FMUL(2.0f, k3v)), k4v) / 6.0f;
}

struct StepParams {
    CurrentStages cs;
    double dt;
    float dt32;
    int32_t has_k1;
    int32_t pos_f32, z_f64;
    int64_t n;
    double* lon;
    double* lat;
    const void* z;                // float32 or float64 (z_f64)
    const void* factor;
    const int32_t* moving;
    const float* k1u;
    const float* k1v;
    float* env_u;
    float* env_v;
    double truncate_below;
    // extras (od_step_oceandrift)
    int32_t wind_on, wdf_f64, w_on, w_at_surface, diff_on, zio_f64;   // zio_f64: dtype of z_inout
    GroupGeom gwind;
    PairRef pwind;
    const void* wdf;
    double wind_drift_depth;
    GroupGeom gw;
    PairRef pw;
    void* z_inout;                // dtype as z; may be a different buffer than z (after vertical mixing)
    const double* rand_x;
    const double* rand_y;
    const float* diffusivity;
    float diffusivity_const;
    float adt32;
};

// One particle, one step (the body of step_kernel; also compiled for the host by tests/hostshim).
// zs/zy and zsw/zyw are the level tables of the current and the vertical-velocity group.
template <int SCHEME, bool F64, bool EXTRAS>
OD_HD void step_particle(const StepParams& p, int64_t i, const double* zs, const double* zy,
                         const double* zsw, const double* zyw) {
    const GroupGeom& g = p.cs.g;
    const double lon0 = p.lon[i], lat0 = p.lat[i];
    const bool zf32 = p.z_f64 == 0;
    const double z0 = p.z ? (zf32 ? (double)((const float*)p.z)[i] : ((const double*)p.z)[i]) : 0.0;
    double zt = z0;                // drift:truncate_ocean_model_below_m (environment.py:554-562)
    if (p.truncate_below > 0.0 && zt < -p.truncate_below) zt = zf32 ? (double)(float)(-p.truncate_below) : -p.truncate_below;
    const VertW vw = vert_weights(g, zs, zy, zt, zf32);
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const GeodStart gs = geod_start(lat0);

    // stage 1: the start-of-step environment
    float k1u, k1v;
    if (p.has_k1) {
        k1u = p.k1u[i];
        k1v = p.k1v[i];
    } else {
        sample2(g, p.cs.t_start, vw, lon0, lat0, k1u, k1v, p.pos_f32 != 0);
    }
    if (p.env_u) p.env_u[i] = k1u;
    if (p.env_v) p.env_v[i] = k1v;

    float ru, rv;
    rk_velocity<SCHEME>(p.cs, vw, gs, lon0, lat0, p.dt32, k1u, k1v, ru, rv);

    double lon1, lat1;
    if (F64) {
        const double f = p.factor ? ((const double*)p.factor)[i] : 1.0;
        final_move_f64(gs, lon0, OD_DMUL((double)ru, f), OD_DMUL((double)rv, f), mv, p.dt, lon1, lat1);
    } else {
        const float f = p.factor ? ((const float*)p.factor)[i] : 1.0f;
        final_move_f32(gs, lon0, OD_FMUL(ru, f), OD_FMUL(rv, f), mv, p.dt, lon1, lat1);
    }

    if (EXTRAS) {
        // ---- advect_wind (physics_methods.py:712-791): wind sampled at the start-of-step position
        if (p.wind_on) {
            const VertW v0 = {0, 0, 1.0};
            float xw, yw;
            sample2(p.gwind, p.pwind, v0, lon0, lat0, xw, yw, p.pos_f32 != 0);
            const double wdd = fabs(p.wind_drift_depth);
            const bool surface = z0 >= -wdd;
            if (p.wdf_f64 || wdd != 0.0) {
                double wdf = p.wdf_f64 ? ((const double*)p.wdf)[i] : (double)((const float*)p.wdf)[i];
                if (wdd != 0.0) {
                    const double air = wdf;
                    wdf = OD_DMUL(wdf, OD_DADD(wdd, z0)) / wdd;
                    if (z0 > 0.0) wdf = air;
                }
                if (!surface) wdf = 0.0;
                const double xv = OD_DMUL((double)xw, wdf), yv = OD_DMUL((double)yw, wdf);
                if (xv != 0.0 || yv != 0.0) {
                    const GeodStart g1 = geod_start(lat1);
                    final_move_f64(g1, lon1, xv, yv, mv, p.dt, lon1, lat1);
                }
            } else {
                float wdf = ((const float*)p.wdf)[i];
                if (!surface) wdf = 0.0f;
                const float xv = OD_FMUL(xw, wdf), yv = OD_FMUL(yw, wdf);
                if (xv != 0.0f || yv != 0.0f) {
                    const GeodStart g1 = geod_start(lat1);
                    final_move_f32(g1, lon1, xv, yv, mv, p.dt, lon1, lat1);
                }
            }
        }
        // ---- vertical_advection (oceandrift.py:315-350): z = min(0, z + moving*w*dt); w sampled at the
        // start-of-step depth, applied to the current depth (which vertical mixing may already have changed)
        if (p.w_on) {
            const bool zio32 = p.zio_f64 == 0;
            const double zc = zio32 ? (double)((const float*)p.z_inout)[i] : ((const double*)p.z_inout)[i];
            const bool applicable = p.w_at_surface ? (zc <= 0.0) : (zc < 0.0);
            if (applicable) {
                const VertW vww = vert_weights(p.gw, zsw, zyw, zt, zf32);
                const float w = sample1(p.gw, p.pw, vww, lon0, lat0, p.pos_f32 != 0);
                const double zn = fmin(0.0, OD_DADD(zc, OD_DMUL(OD_DMUL(mv, (double)w), p.dt)));
                if (zio32) ((float*)p.z_inout)[i] = (float)zn;
                else ((double*)p.z_inout)[i] = zn;
            }
        }
        // ---- horizontal_diffusion (basemodel/__init__.py:1746-1772)
        if (p.diff_on) {
            const float D = p.diffusivity ? p.diffusivity[i] : p.diffusivity_const;
            const float s = sqrtf(OD_FMUL(2.0f, D) / p.adt32);
            const double sd = OD_DMUL(mv, (double)s);
            const double xv = OD_DMUL(sd, p.rand_x[i]), yv = OD_DMUL(sd, p.rand_y[i]);
            if (xv != 0.0 || yv != 0.0) {
                const GeodStart g1 = geod_start(lat1);
                final_move_f64(g1, lon1, xv, yv, mv, p.dt, lon1, lat1);
            }
        }
    }
    p.lon[i] = lon1;
    p.lat[i] = lat1;
}

}  // namespace od
