// od_stokes.cuh -- Stokes drift velocity of one particle and its geodesic move.
//
// Restates PhysicsMethods.stokes_drift (opendrift/models/physics_methods.py:793-848) with the depth profiles
// stokes_drift_profile_{monochromatic,exponential,phillips} (:332-416), Hs from the environment or from wind
// (significant_wave_height, :893-906: 0.0246 |U|^2) and the wave period from wind (wave_period /
// _wave_frequency, :908-943: omega = 0.877 g / (1.17 |U|), 5 rad/s without wind), keeping NumPy's dtype flow
// (float32 environment; float32 wind speed, Hs and omega arithmetic; everything downstream float64).
// The collective decisions of the reference (is there any Stokes drift / any Hs / any wind at all) are taken by
// the host from od_minmax_f32 reductions and arrive here as hs_mode.
#pragma once
#include "od_advect.cuh"

namespace od {

struct StokesParams {
    int64_t n;
    double* lon;
    double* lat;
    const void* z;
    const float* us;
    const float* vs;
    const float* hs;
    const float* xwind;
    const float* ywind;
    const int32_t* moving;
    double dt;
    int32_t z_f64, hs_mode /* 0 env, 1 from wind, 2 constant 1 */, profile /* 0 mono, 1 exp, 2 Phillips */, pad_;
};

OD_HD double erfc_(double x) { return erfc(x); }

OD_HD void stokes_particle(const StokesParams& p, int64_t i) {
    const float us = p.us[i], vs = p.vs[i];
    const float speed = sqrtf(OD_FADD(OD_FMUL(us, us), OD_FMUL(vs, vs)));
    if (speed == 0.0f) return;                       // zeromask: zero velocity, nothing moves
    const double z = p.z_f64 ? ((const double*)p.z)[i] : (double)((const float*)p.z)[i];
    const float xw = p.xwind ? p.xwind[i] : 0.0f, yw = p.ywind ? p.ywind[i] : 0.0f;
    const float ws = sqrtf(OD_FADD(OD_FMUL(xw, xw), OD_FMUL(yw, yw)));
    // wave period from wind (float32 quotient stored into a float64 array)
    const double omega = ws > 0.0f ? (double)((float)(0.877 * 9.81) / OD_FMUL((float)1.17, ws)) : 5.0;
    const double T = (2 * 3.141592653589793) / omega;
    const double freq = 2. * 3.141592653589793 / T;
    double hs2;                                      // np.power(Hs, 2)
    if (p.hs_mode == 0) hs2 = (double)OD_FMUL(p.hs[i], p.hs[i]);
    else if (p.hs_mode == 1) { const float h = OD_FMUL((float)0.0246, OD_FMUL(ws, ws)); hs2 = (double)OD_FMUL(h, h); }
    else hs2 = 1.0;
    const double transport = OD_DMUL(freq, hs2) / 16;
    double unit;
    if (p.profile == 0) {
        const double km = (double)speed / OD_DMUL(2.0, transport);
        unit = exp(OD_DMUL(OD_DMUL(2.0, km), z));
    } else if (p.profile == 1) {
        const double km = (double)speed / OD_DMUL(2.0, transport);
        const double ke = km / 3;
        unit = exp(OD_DMUL(OD_DMUL(2.0, ke), z)) / OD_DSUB(1.0, OD_DMUL(OD_DMUL(8.0, ke), z));
    } else {
        const double km = (double)OD_FMUL(speed, (float)(1 - 2 * 1 / 3.0)) / OD_DMUL(2.0, transport);
        const double az = fabs(z);
        unit = OD_DSUB(exp(OD_DMUL(OD_DMUL(2.0, km), z)),
                       OD_DMUL(sqrt(OD_DMUL(OD_DMUL(2 * 3.141592653589793, km), az)),
                               erfc_(sqrt(OD_DMUL(OD_DMUL(2.0, km), az)))));
    }
    const double su = OD_DMUL((double)us, unit), sv = OD_DMUL((double)vs, unit);
    if (su == 0.0 && sv == 0.0) return;
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const GeodStart gs = geod_start(p.lat[i]);
    double lo, la;
    final_move_f64(gs, p.lon[i], su, sv, mv, p.dt, lo, la);
    p.lon[i] = lo;
    p.lat[i] = la;
}

}  // namespace od
