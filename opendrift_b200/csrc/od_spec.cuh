// od_spec.cuh -- the Runge-Kutta step of od_advect.cuh specialised for the launch the benchmarked configurations make
// (BASELINE.json configs[1] / [3]): a geographic, non-periodic 3-D current group sampled between two reader times, float64
// positions, default (series) arithmetic.  Same reference operations as the general step (advect_ocean_current,
// opendrift/models/physics_methods.py:611-691; the reader chain of od_interp.cuh), same roundings, same results bit for bit --
// what changes is what a thread has to decide per sample:
//   * everything that is the same for all particles of the launch is decided by the host (spec_eligible): projection, periodic
//     / global longitude, 2-D blocks, float32 positions, a given k1, uncertainty draws, reader priority lists, the reciprocal
//     division, a reader that does not cover the time;
//   * everything that is rare per particle is not solved where it occurs but FLAGGED, as SeriesHot already does for moves
//     outside the short-arc series (od_advect.cuh): a sample in the last row / column of the block or clamped onto it, a
//     longitude outside [-360, 360) or a mid-point that crosses the antimeridian, a numerator of the index division too small
//     for the reciprocal shortcut.  A flagged particle writes nothing and is redone from its start state by the general
//     step (step_particle_redo, out of line at the end of the thread).
// The sampler then is straight-line code: the four corners of a cell are p, p + 1, p + nx, p + nx + 1 (one address per layer,
// immediate offsets for the rest), the second layer is the first plus a stride kept from the start of the step, both reader
// times of both components arrive in one 16-byte load per corner, and there is no clamp, no wrap and no coverage branch
// (an uncovered sample reads texel 0 and is replaced by NaN -> fallback afterwards).
#pragma once
#include "od_advect.cuh"

namespace od {

static inline bool spec_pair_ok(const PairRef& r) { return r.mode >= 0 && r.mode <= 2 && r.tex != nullptr; }

// what the host checks before launching the specialised kernel (also used by tests/hostshim)
static inline bool spec_eligible(const StepParams& p, int scheme) {
    const GroupGeom& g = p.cs.g;
    const double ax = g.xspan < 0 ? -g.xspan : g.xspan, ay = g.yspan < 0 ? -g.yspan : g.yspan;
    return scheme == 2 && g.proj_kind == 0 && g.wrap == 0 && g.glob == 0 && g.nz > 1 && g.ncomp == 2 && g.nx >= 2 && g.ny >= 2 &&
           g.rxspan != 0.0 && g.ryspan != 0.0 && ax >= 1e-9 && ax <= 1e9 && ay >= 1e-9 && ay <= 1e9 &&
           !p.has_k1 && !p.pos_f32 && !p.noise_cur && p.n_chain == 0 &&
           spec_pair_ok(p.cs.t_start) && spec_pair_ok(p.cs.t_mid) && spec_pair_ok(p.cs.t_end);
}

static inline bool spec_all_lerp(const StepParams& p) { return p.cs.t_start.mode == 0 && p.cs.t_mid.mode == 0 && p.cs.t_end.mode == 0; }

// np_mod360 (od_interp.cuh) on [-360, 360): fmod(x, 360) is x itself there, negative values get + 360 (-360 gives +0 either
// way); anything else (and NaN) is flagged
OD_HD double mod360_spec(double x, bool& bad) {
    if (!(x >= -360.0 && x < 360.0)) bad = true;
    return x < 0.0 ? OD_DADD(x, 360.0) : x;
}

// div_rn (od_interp.cuh) without its branches: the reciprocal is valid (spec_eligible), a non-zero numerator below 2^-830
// (the general code takes the true division below 1e-250) is flagged by its exponent field
OD_HD double div_rn_spec(double d, double s, double r, bool& bad) {
#if defined(__CUDA_ARCH__)
    const unsigned e2 = ((unsigned)__double2hiint(d)) << 1;              // exponent + top of the mantissa, sign shifted out
    if (e2 - 1u < (193u << 21) - 1u) bad = true;                         // 0 < |d| < 2^-830 (a zero high word passes: 0 / s = 0 either way)
    const double q0 = __dmul_rn(d, r);
    const double e = __fma_rn(-q0, s, d);
    return __fma_rn(e, r, q0);
#else
    (void)r; (void)bad;
    return d / s;
#endif
}

// horiz_weights (od_interp.cuh) for an interior cell.  covered: the position passes the reader's coverage test; h.valid: covered
// and the cell is interior (0 <= ix < nx - 1, 0 <= iy < ny - 1), so that ix1 = ix + 1, iy1 = iy + 1 and no index is clamped;
// covered but not interior -> flagged.  h.i00 is 0 unless h.valid (the loads stay in bounds whatever the position was).
OD_HD HorizW horiz_spec(const GroupGeom& g, double lon, double lat, bool& covered, bool& bad) {
    HorizW h;
    const double x = (g.lon_mode == 0) ? mod360_spec(lon, bad) : OD_DSUB(mod360_spec(OD_DADD(lon, 180.0), bad), 180.0);
    const double xi = OD_DMUL(div_rn_spec(OD_DSUB(x, g.x0), g.xspan, g.rxspan, bad), g.nxm1);
    const double yi = OD_DMUL(div_rn_spec(OD_DSUB(lat, g.y0), g.yspan, g.ryspan, bad), g.nym1);
    covered = (x >= g.xmin) && (x <= g.xmax) && (lat >= g.ymin) && (lat <= g.ymax);
#if defined(__CUDA_ARCH__) && defined(OD_SPEC_MAGIC_FLOOR)
    // floor and integer part without the conversion unit: xi + 1.5 * 2^52 holds the integer nearest to xi in its low word
    // (|xi| < 2^31; a covered position is within a cell of [0, nx - 1], an uncovered one is not used)
    const double kM = 6755399441055744.0;
    const double tx = __dadd_rn(xi, kM), ty = __dadd_rn(yi, kM);
    double fx = __dsub_rn(tx, kM), fy = __dsub_rn(ty, kM);
    int ix = __double2loint(tx), iy = __double2loint(ty);
    if (fx > xi) { fx = __dsub_rn(fx, 1.0); ix -= 1; }
    if (fy > yi) { fy = __dsub_rn(fy, 1.0); iy -= 1; }
#else
    const double fx = floor(xi), fy = floor(yi);
    const int ix = (int)fx, iy = (int)fy;
#endif
    const bool interior = (unsigned)ix < (unsigned)(g.nx - 1) && (unsigned)iy < (unsigned)(g.ny - 1);
    if (covered && !interior) bad = true;
    h.valid = covered && interior;
    h.wx0 = OD_DSUB(1.0, OD_DSUB(xi, fx));
    h.wx1 = OD_DSUB(1.0, h.wx0);
    h.wy0 = OD_DSUB(1.0, OD_DSUB(yi, fy));
    h.wy1 = OD_DSUB(1.0, h.wy0);
    h.ix = ix; h.iy = iy; h.ix1 = ix + 1; h.iy1 = iy + 1;
    h.i00 = h.valid ? iy * g.nx + ix : 0;
    h.i01 = h.i00 + 1;
    h.i10 = h.i00 + g.nx;
    h.i11 = h.i10 + 1;
    return h;
}

// one pair texel.  OD_SPEC_LD128: a single 16-byte load (the compiler splits a plain float4 load in two 8-byte loads because
// either half is only used in one of the time-mode branches -- measured faster than the 16-byte load, so the default)
OD_HD Tex4 fetch4_spec(const float* p) {
#if defined(__CUDA_ARCH__) && defined(OD_SPEC_LD128)
    Tex4 r;
    asm("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return fetch4(p, 0);
#endif
}

// the two components at both reader times of one layer: sample2_h's layer_bilin with the corners at fixed offsets
OD_HD LayerVals layer_spec(const HorizW& h, const float* p00, int row, int mode) {
    const Tex4 a00 = fetch4_spec(p00), a01 = fetch4_spec(p00 + 4), a10 = fetch4_spec(p00 + row), a11 = fetch4_spec(p00 + row + 4);
    LayerVals r = {0.f, 0.f, 0.f, 0.f};
    if (mode != 2) {
        r.uA = bilin(h, a00.x, a01.x, a10.x, a11.x);
        r.vA = bilin(h, a00.y, a01.y, a10.y, a11.y);
    }
    if (mode != 1) {
        r.uB = bilin(h, a00.z, a01.z, a10.z, a11.z);
        r.vB = bilin(h, a00.w, a01.w, a10.w, a11.w);
    }
    return r;
}

// combine (od_interp.cuh) for a 3-D block
OD_HD float combine_spec(const PairRef& pr, const VertW& vw, int mode, float haA, float hbA, float haB, float hbB) {
    const double omw = OD_DSUB(1.0, vw.wa);
    double vA = 0.0, vB = 0.0;
    if (mode != 2) vA = OD_DADD(OD_DMUL((double)haA, vw.wa), OD_DMUL((double)hbA, omw));
    if (mode == 1) return (float)vA;
    vB = OD_DADD(OD_DMUL((double)haB, vw.wa), OD_DMUL((double)hbB, omw));
    if (mode == 2) return (float)vB;
    return (float)OD_DADD(OD_DMUL(vA, OD_DSUB(1.0, pr.w)), OD_DMUL(vB, pr.w));
}

// sample2_h for the specialised step.  lay_a: float offset of layer vw.ia in the pair texels, dlay: float offset from it to
// layer vw.ib (both fixed for the step: the depth does not change between the stages).
// LERP: every time sample of the launch lies between two reader times (mode 0 known at compile time: no time-mode branch at all)
template <bool LERP>
OD_HD void sample2_spec(const GroupGeom& g, const PairRef& pr, const VertW& vw, long long lay_a, long long dlay, const HorizW& h,
                        bool covered, float& u, float& v) {
    const int mode = LERP ? 0 : pr.mode;
    const float* pa = pr.tex + (lay_a + 4ll * h.i00);
    const int row = 4 * g.nx;
    const LayerVals A = layer_spec(h, pa, row, mode);
    const LayerVals B = layer_spec(h, pa + dlay, row, mode);
    float ru = combine_spec(pr, vw, mode, A.uA, B.uA, A.uB, B.uB);
    float rv = combine_spec(pr, vw, mode, A.vA, B.vA, A.vB, B.vB);
    if (!covered) ru = rv = NAN;
    if (!finite_f(ru)) ru = g.fallback[0];
    if (!finite_f(rv)) rv = g.fallback[1];
    u = ru;
    v = rv;
}

// One particle, one step.  Returns 0: done; 1: flagged before anything was written (redo the whole step); 2: flagged after the
// depth update of vertical advection (redo the moves only).
template <int SCHEME, bool F64, int EXTRAS, bool LERP>
OD_HD int step_particle_spec(const StepParams& p, int64_t i, const double* zs, const double* zy, const double* zsw, const double* zyw) {
    typedef SeriesHot MATH;
    const GroupGeom& g = p.cs.g;
    const double lon0 = p.lon[i], lat0 = p.lat[i];
    const bool zf32 = p.z_f64 == 0;
    const double z0 = p.z ? (zf32 ? (double)((const float*)p.z)[i] : ((const double*)p.z)[i]) : 0.0;
    double zt = z0;                // drift:truncate_ocean_model_below_m (environment.py:554-562)
    if (p.truncate_below > 0.0 && zt < -p.truncate_below) zt = zf32 ? (double)(float)(-p.truncate_below) : -p.truncate_below;
    const VertW vw = vert_weights(g, zs, zy, zt, zf32);
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const MATH::Start gs = MATH::start(lat0);
    const long long plane = 4ll * g.nx * g.ny;
    const long long lay_a = vw.ia * plane, dlay = (vw.ib - vw.ia) * plane;

    // stage 1: the start-of-step environment
    bool bad = false, cov0;
    const HorizW h0 = horiz_spec(g, lon0, lat0, cov0, bad);
    if (bad) return 1;
    float k1u, k1v;
    sample2_spec<LERP>(g, p.cs.t_start, vw, lay_a, dlay, h0, cov0, k1u, k1v);
    if (p.env_u) p.env_u[i] = k1u;
    if (p.env_v) p.env_v[i] = k1v;

    if (EXTRAS) {
        // ---- vertical_advection (oceandrift.py:315-350), as in step_particle
        if (p.w_on) {
            const bool zio32 = p.zio_f64 == 0;
            const double zc = zio32 ? (double)((const float*)p.z_inout)[i] : ((const double*)p.z_inout)[i];
            const bool applicable = p.w_at_surface ? (zc <= 0.0) : (zc < 0.0);
            if (applicable) {
                float w;
                if (p.w_same_grid) {
                    w = sample1_h(p.gw, p.pw, vw, h0);
                } else {
                    const VertW vww = vert_weights(p.gw, zsw, zyw, zt, zf32);
                    w = MATH::sample_s(p.gw, p.pw, vww, lon0, lat0, false);
                }
                const double zn = fmin(0.0, OD_DADD(zc, OD_DMUL(OD_DMUL(mv, (double)w), p.dt)));
                if (zio32) ((float*)p.z_inout)[i] = (float)zn;
                else ((double*)p.z_inout)[i] = zn;
            }
        }
    }

    // stages 2..4 (rk_velocity of od_advect.cuh): mid-points by the third-order series from the normalised start longitude
    const double lon0n = ang_normalize(lon0);
    const double hdt = OD_DMUL((double)p.dt32, 0.5);
    float ku = k1u, kv = k1v, su = k1u, sv = k1v;
    const int last = SCHEME == 1 ? 1 : 3;
    // (measured on B200, 10 M particles: the all-lerp instantiation is fastest with the three passes unrolled -- the time pair of
    // each pass is then a compile-time choice, 0.897 -> 0.888 ms; the instantiation with time-mode branches is not, 0.927 -> 0.931)
#if defined(__CUDA_ARCH__)
#pragma unroll(LERP ? 3 : 1)
#endif
    for (int st = 1; st <= last; ++st) {
        double mlon, mlat;
        if (!series_move3_raw(gs, lon0n, OD_DMUL((double)kv, hdt), OD_DMUL((double)ku, hdt), mlon, mlat)) bad = true;
        if (!(fabs(mlon) < 180.0)) bad = true;                 // ang_normalize would wrap (or NaN)
        const PairRef& pr = st == 3 ? p.cs.t_end : p.cs.t_mid;
        bool cov;
        const HorizW h = horiz_spec(g, mlon, mlat, cov, bad);
        sample2_spec<LERP>(g, pr, vw, lay_a, dlay, h, cov, ku, kv);
        if (st < 3) {
            su = OD_FADD(su, OD_FMUL(2.0f, ku));
            sv = OD_FADD(sv, OD_FMUL(2.0f, kv));
        } else {
            su = OD_FADD(su, ku);
            sv = OD_FADD(sv, kv);
        }
    }
    float ru, rv;
    if (SCHEME == 1) {
        ru = ku;
        rv = kv;
    } else {
        ru = su / 6.0f;
        rv = sv / 6.0f;
    }

    double lon1, lat1;
    if (F64) {
        const double f = p.factor ? ((const double*)p.factor)[i] : 1.0;
        do_move64<MATH>(gs, lon0, lat0, OD_DMUL((double)ru, f), OD_DMUL((double)rv, f), mv, p.dt, lon1, lat1, bad);
    } else {
        const float f = p.factor ? ((const float*)p.factor)[i] : 1.0f;
        do_move32<MATH>(gs, lon0, lat0, OD_FMUL(ru, f), OD_FMUL(rv, f), mv, p.dt, lon1, lat1, bad);
    }
    if (EXTRAS == 1) {
        if (p.wind_on) extras_wind<MATH, false>(p, i, z0, mv, lon0, lat0, &lon1, &lat1, bad);
        if (p.diff_on) extras_diffusion<MATH>(p, i, mv, &lon1, &lat1, bad);
    }
    if (bad) return 2;
    p.lon[i] = lon1;
    p.lat[i] = lat1;
    return 0;
}

}  // namespace od
