// od_leeway.cuh -- one Leeway time step of one drifting object.
//
// Restates Leeway.update (opendrift/models/leeway.py:430-494): optional capsizing (:438-454), wind speed and direction from the
// float32 environment, down- and cross-wind leeway from the per-element coefficients (float32, as NumPy computes
// them), the leeway move and then the current move (two sequential update_positions, basemodel/__init__.py:4630),
// both sampled at the START-of-step position, and the random jibing (crosswind_slope -> -crosswind_slope,
// orientation -> 1 - orientation).  Elements whose wind or current sample is missing (fallback None in
// Leeway.required_variables) are flagged in `status` and do not move (report_missing_variables,
// basemodel/__init__.py:2501-2515).
#pragma once
#include "od_mix.cuh"

namespace od {

struct LeewayParams {
    GroupGeom gwind, gcur;
    PairRef pwind, pcur;
    int64_t n;
    double* lon;
    double* lat;
    const float* dw_slope; const float* dw_offset; const float* dw_eps;
    float* cw_slope; const float* cw_offset; const float* cw_eps;
    uint8_t* orientation;
    uint8_t* capsized;                // NULL = none capsized; toggled when processes:capsizing is on
    const void* jibe_probability;     // float32 or float64 (jp_f64)
    const int32_t* moving;
    int32_t* status;                  // may be NULL
    const int32_t* ids;
    const double* rand;               // np.random.random(n) of this step, or NULL -> Philox
    double dt;
    unsigned long long seed;
    float capsize_fraction;
    int32_t jp_f64, pos_f32, step_index, missing_code;
    // processes:capsizing (:438-454): an eligible element (capsized == capsize_from: 0 in forward runs, 1 in backward runs)
    // flips with probability (0.5 + 0.5 tanh((wind - threshold) / sigma)) |dt| / 3600
    int32_t capsize_on, capsize_from;
    float wind_threshold, wind_sigma;
    const double* rand_capsize;       // [n] np.random.rand draws laid out per element (entries of ineligible elements unused), or NULL -> Philox
    const double* noise_cur;          // uncertainty draws of the step (environment.py:869-891), see od_leeway_args
    const double* noise_wind;
    int32_t noise_kinds, pad2_;
};

template <bool PROJ = false>
OD_HD void leeway_particle(const LeewayParams& p, int64_t i) {
    const double lon0 = p.lon[i], lat0 = p.lat[i];
    const VertW v0 = {0, 0, 1.0};
    float xw, yw, cu, cv;
    if (PROJ) {
        sample2_any(p.gwind, p.pwind, v0, lon0, lat0, xw, yw, p.pos_f32 != 0);
        sample2_any(p.gcur, p.pcur, v0, lon0, lat0, cu, cv, p.pos_f32 != 0);
    } else {
        sample2(p.gwind, p.pwind, v0, lon0, lat0, xw, yw, p.pos_f32 != 0);
        sample2(p.gcur, p.pcur, v0, lon0, lat0, cu, cv, p.pos_f32 != 0);
    }
    if (p.noise_cur) {                                   // env[var] += draw on float32 arrays: normal first, then uniform
        for (int kind = 0; kind < 2; ++kind) {
            if (!(p.noise_kinds & (1 << kind))) continue;
            const double* base = p.noise_cur + (int64_t)(kind * 2) * p.n;
            cu = (float)OD_DADD((double)cu, base[i]);
            cv = (float)OD_DADD((double)cv, base[p.n + i]);
        }
    }
    if (p.noise_wind) {
        xw = (float)OD_DADD((double)xw, p.noise_wind[i]);
        yw = (float)OD_DADD((double)yw, p.noise_wind[p.n + i]);
    }
    if (!(finite_f(xw) && finite_f(yw) && finite_f(cu) && finite_f(cv))) {
        if (p.status && p.status[i] == 0) p.status[i] = p.missing_code;
        return;
    }
    if (p.status && p.status[i] != 0) return;          // already scheduled for removal
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const float ws = sqrtf(OD_FADD(OD_FMUL(xw, xw), OD_FMUL(yw, yw)));
    const float wd = atan2f(xw, yw);
    const float dwe = p.dw_eps[i], cwe = p.cw_eps[i];
    const float dl = OD_FMUL(OD_FADD(OD_FADD(OD_FMUL(OD_FADD(p.dw_slope[i], dwe / 20.0f), ws), p.dw_offset[i]), dwe / 2.0f), 0.01f);
    const float cl = OD_FMUL(OD_FADD(OD_FADD(OD_FMUL(OD_FADD(p.cw_slope[i], cwe / 20.0f), ws), p.cw_offset[i]), cwe / 2.0f), 0.01f);
    const float sinth = sinf(wd), costh = cosf(wd);
    float yl = OD_FADD(OD_FMUL(dl, costh), OD_FMUL(cl, sinth));
    float xl = OD_FADD(OD_FMUL(-dl, sinth), OD_FMUL(cl, costh));
    if (p.capsize_on && p.capsized && p.capsized[i] == (uint8_t)p.capsize_from) {
        // float32 probability per hour, promoted to float64 by np.abs(dt) (a NumPy float64 scalar)
        const float ph = OD_FADD(0.5f, OD_FMUL(0.5f, tanhf(OD_FADD(ws, -p.wind_threshold) / p.wind_sigma)));
        const double prob = OD_DMUL((double)ph, fabs(p.dt)) / 3600;
        double Uc, spare_c;
        if (p.rand_capsize) Uc = p.rand_capsize[i];
        else philox_uniform2(p.seed, p.ids ? (unsigned)p.ids[i] : (unsigned)i, (unsigned)p.step_index, 0x43415053u, Uc, spare_c);
        if (Uc < prob) p.capsized[i] = (uint8_t)(1 - p.capsized[i]);
    }
    if (p.capsized && p.capsized[i] == 1) {
        xl = OD_FMUL(xl, p.capsize_fraction);
        yl = OD_FMUL(yl, p.capsize_fraction);
    }
    double lon1, lat1;
    final_move_f32(geod_start(lat0), lon0, -xl, yl, mv, p.dt, lon1, lat1);          // update_positions(-x_leeway, y_leeway)
    final_move_f32(geod_start(lat1), lon1, cu, cv, mv, p.dt, lon1, lat1);           // update_positions(current)
    p.lon[i] = lon1;
    p.lat[i] = lat1;
    // jibing (:478-488)
    const double jp = p.jp_f64 ? ((const double*)p.jibe_probability)[i] : (double)((const float*)p.jibe_probability)[i];
    const double rate = -log(OD_DSUB(1.0, jp)) / 3600;
    const double pstep = OD_DSUB(1.0, exp(OD_DMUL(-rate, fabs(p.dt))));
    double U, spare;
    if (p.rand) U = p.rand[i];
    else philox_uniform2(p.seed, p.ids ? (unsigned)p.ids[i] : (unsigned)i, (unsigned)p.step_index, 0x4a494245u, U, spare);
    if (pstep > U) {
        p.cw_slope[i] = -p.cw_slope[i];
        p.orientation[i] = (uint8_t)(1 - p.orientation[i]);
    }
}

}  // namespace od
