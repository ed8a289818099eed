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

// ---------------------------------------------------------------------------------------------------------
// Arithmetic policies of the step.  ExactMath is the restatement of the reference (bit-exact field sampling,
// Karney geodesic); SeriesMath (the default) keeps the bit-exact sampling and evaluates the moves with the short-arc
// series; FastMath (opt-in) is SeriesMath with the field sampled in float32 (below).
// ---------------------------------------------------------------------------------------------------------
struct ExactMath {
    typedef GeodStart Start;
    static constexpr bool kExactSampler = true;      // sample_uv == sample2: the horizontal weights can be shared
    static constexpr bool kDefer = false;            // moves outside the short-arc series' range are solved in place
    OD_HDS Start start(double lat0) { return geod_start(lat0); }
    OD_HDS void midpoint(const Start& s, double lon0, double lat0, float ku, float kv, float dt32, double& mlon, double& mlat) {
        rk_midpoint(s, lon0, ku, kv, dt32, mlon, mlat);
    }
    OD_HDS void move32(const Start& s, double lon0, double lat0, float xv, float yv, double mv, double dt, double& lon1, double& lat1) {
        final_move_f32(s, lon0, xv, yv, mv, dt, lon1, lat1);
    }
    OD_HDS void move64(const Start& s, double lon0, double lat0, double xv, double yv, double mv, double dt, double& lon1, double& lat1) {
        final_move_f64(s, lon0, xv, yv, mv, dt, lon1, lat1);
    }
    OD_HDS void sample_uv(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, float& u, float& v, bool pos_f32,
                          const TileView& tv = TileView()) {
        sample2(g, pr, vw, lon, lat, u, v, pos_f32, tv);
    }
    OD_HDS float sample_s(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, bool pos_f32) {
        return sample1(g, pr, vw, lon, lat, pos_f32);
    }
};

// SeriesMath: the reference's field sampling bit for bit (ExactMath's sampler) with every move evaluated by the
// fifth-order short-arc series of the direct geodesic (od_geod.cuh: series_move) -- the same positions as the full
// solution to <= 1e-13 deg, at a tenth of its FP64 instructions.  Differences from ExactMath:
//   * RK mid-points take the displacement components (0.5 dt ku, 0.5 dt kv) directly instead of the reference's
//     float32 azimuth / float32 distance (physics_methods.py:629-631); the rounding this skips moves a mid-point by
//     <= 2e-5 m, which changes a sampled float32 velocity in the last bit at most (positions: ~1e-10 deg, below the
//     ~1e-9 deg/step that NumPy's own non-reproducible float32 arctan2 puts on the reference);
//   * float64 final moves (the default dtype flow) take (dt xv, dt yv) directly: closer to the reference's
//     float64 arctan2 -> geodesic chain than any re-implementation of that chain;
//   * float32 final moves keep the float32 azimuth and speed of update_positions (basemodel/__init__.py:4643-4650).
struct SeriesMath : ExactMath {
    typedef SeriesStart Start;
    OD_HDS Start start(double lat0) { return series_start(lat0); }
    OD_HDS void midpoint(const Start& s, double lon0, double lat0, float ku, float kv, float dt32, double& mlon, double& mlat) {
        const double h = OD_DMUL((double)dt32, 0.5);
        const double xn = OD_DMUL((double)kv, h), ye = OD_DMUL((double)ku, h);
        if (series_move3(s, lon0, xn, ye, mlon, mlat)) return;      // (same arithmetic as SeriesHot: a redone particle repeats its mid-points to the bit)
        geod_move_ne(s, lon0, xn, ye, mlon, mlat);
    }
    OD_HDS void move32(const Start& s, double lon0, double lat0, float xv, float yv, double mv, double dt, double& lon1, double& lat1) {
        const float az = az_f32(xv, yv);
        const double dist = OD_DMUL(OD_DMUL((double)speed_f32(xv, yv), mv), dt);
        double sa, ca;
        sincosd(ang_round(ang_normalize((double)az)), sa, ca);
        geod_move_ne(s, lon0, OD_DMUL(dist, ca), OD_DMUL(dist, sa), lon1, lat1);
    }
    OD_HDS void move64(const Start& s, double lon0, double lat0, double xv, double yv, double mv, double dt, double& lon1, double& lat1) {
        const double k = OD_DMUL(mv, dt);
        geod_move_ne(s, lon0, OD_DMUL(yv, k), OD_DMUL(xv, k), lon1, lat1);
    }
};

// SeriesHot: SeriesMath for the hot path of the step kernels.  A move outside the range of the short-arc series (long steps,
// the polar caps, NaN) is not solved where it occurs -- the out-of-line call to the full solution there made the compiler
// save and restore the live registers of the Runge-Kutta loop around it on EVERY pass (ABI call inside the loop: 170 local
// loads and 90 local stores per particle-step, two thirds of the kernel's L1 traffic) -- but flagged: the particle finishes the
// step with the move skipped, writes nothing, and is then redone from its untouched start state by the full SeriesMath path
// in one out-of-line call at the very end of the thread (step_kernel), where nothing is live any more.
struct SeriesHot : SeriesMath {
    static constexpr bool kDefer = true;
    OD_HDS void midpoint_d(const Start& s, double lon0, double lat0, float ku, float kv, float dt32, double& mlon, double& mlat, bool& bad) {
        const double h = OD_DMUL((double)dt32, 0.5);
        if (!series_move3(s, lon0, OD_DMUL((double)kv, h), OD_DMUL((double)ku, h), mlon, mlat)) { bad = true; mlon = lon0; mlat = lat0; }
    }
    OD_HDS void move32_d(const Start& s, double lon0, double lat0, float xv, float yv, double mv, double dt, double& lon1, double& lat1, bool& bad) {
        const float az = az_f32(xv, yv);
        const double dist = OD_DMUL(OD_DMUL((double)speed_f32(xv, yv), mv), dt);
        double sa, ca;
        sincosd(ang_round(ang_normalize((double)az)), sa, ca);
        if (!series_move(s, lon0, OD_DMUL(dist, ca), OD_DMUL(dist, sa), lon1, lat1)) { bad = true; lon1 = lon0; lat1 = lat0; }
    }
    OD_HDS void move64_d(const Start& s, double lon0, double lat0, double xv, double yv, double mv, double dt, double& lon1, double& lat1, bool& bad) {
        const double k = OD_DMUL(mv, dt);
        if (!series_move(s, lon0, OD_DMUL(yv, k), OD_DMUL(xv, k), lon1, lat1)) { bad = true; lon1 = lon0; lat1 = lat0; }
    }
};

// the three moves of a step through the policy, deferring variant when the policy has one
template <class MATH>
OD_HD void do_midpoint(const typename MATH::Start& s, double lon0, double lat0, float ku, float kv, float dt32, double& mlon, double& mlat, bool& bad) {
    if constexpr (MATH::kDefer) MATH::midpoint_d(s, lon0, lat0, ku, kv, dt32, mlon, mlat, bad);
    else MATH::midpoint(s, lon0, lat0, ku, kv, dt32, mlon, mlat);
}
template <class MATH>
OD_HD void do_move32(const typename MATH::Start& s, double lon0, double lat0, float xv, float yv, double mv, double dt, double& lon1, double& lat1, bool& bad) {
    if constexpr (MATH::kDefer) MATH::move32_d(s, lon0, lat0, xv, yv, mv, dt, lon1, lat1, bad);
    else MATH::move32(s, lon0, lat0, xv, yv, mv, dt, lon1, lat1);
}
template <class MATH>
OD_HD void do_move64(const typename MATH::Start& s, double lon0, double lat0, double xv, double yv, double mv, double dt, double& lon1, double& lat1, bool& bad) {
    if constexpr (MATH::kDefer) MATH::move64_d(s, lon0, lat0, xv, yv, mv, dt, lon1, lat1, bad);
    else MATH::move64(s, lon0, lat0, xv, yv, mv, dt, lon1, lat1);
}

// FastMath (opt-in, od_advect_args.fast = OD_MATH_FAST): SeriesMath's moves -- float64 positions, every mid-point and every kept
// move the round-off-accurate short-arc series -- with the field sampled in float32: the time lerp of the eight corner texels, then
// the trilinear interpolation, with float32 FMAs (the fractional cell index is still formed in float64 from the float64
// position).  What it gives up against the reference is the float64 accumulation order of scipy's map_coordinates: a sampled
// velocity differs in its last float32 bits (relative 1e-7), positions by ~1e-9 deg per step.  (Round 1's FastMath also took
// first-order mid-points; they neglected the convergence of the meridians -- 5e-6 deg at 80-86 N in the randomised replays --
// and are gone.)
OD_HD float lerp_f32(float a, float b, float t) { return fmaf(t, b - a, a); }

// time lerp of the corners, then trilinear, float32 FMAs
OD_HD void fast_sample_uv(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, float& u, float& v, bool pos_f32,
                          const TileView& tv) {
    if (pos_f32) {       // first step after seeding: the reference forms the cell index in float32 (ulp 3e-5 at lon 359);
        sample2(g, pr, vw, lon, lat, u, v, true, tv);     // in a strong gradient that is visible, so replay it exactly
        return;
    }
    float ru = NAN, rv = NAN;
    double x = (g.lon_mode == 0) ? np_mod360(lon) : np_mod360(lon + 180.0) - 180.0;
    double xi = (x - g.x0) * g.inv_dx, yi = (lat - g.y0) * g.inv_dy;
    if (pr.mode != 3 && (g.glob != 0 || (x >= g.xmin && x <= g.xmax)) && lat >= g.ymin && lat <= g.ymax &&
        xi == xi && yi == yi) {
        xi = xi < 0.0 ? 0.0 : (xi > g.nxm1 ? g.nxm1 : xi);       // covered: edge value (see horiz_weights)
        yi = yi < 0.0 ? 0.0 : (yi > g.nym1 ? g.nym1 : yi);
        const double fx = floor(xi), fy = floor(yi);
        int ix = (int)fx;
        const int iy = (int)fy;
        const int nxv = g.nx + g.wrap;
        int ix1 = ix + 1 < nxv ? ix + 1 : nxv - 1;
        if (ix >= g.nx) ix -= g.nx;
        if (ix1 >= g.nx) ix1 -= g.nx;
        const int iy1 = iy + 1 < g.ny ? iy + 1 : g.ny - 1;
        const float tx = (float)(xi - fx), ty = (float)(yi - fy);
        const float tw = pr.mode == 0 ? (float)pr.w : (pr.mode == 1 ? 0.0f : 1.0f);
        const TexelSource ts = texel_source(pr.tex, tv, g.nx, g.ny, ix, ix1, iy, iy1, vw.ia, g.nz > 1 ? vw.ib : vw.ia);
        float lu[2], lvv[2];
        const int nl = g.nz > 1 ? 2 : 1;
        for (int l = 0; l < nl; ++l) {
            const int lay = l == 0 ? vw.ia : vw.ib;
            const float* lp = layer_ptr(ts, lay);
            const int r0 = 4 * iy * ts.lx, r1 = 4 * iy1 * ts.lx;
            const Tex4 a00 = fetch4(lp, r0 + 4 * ix), a01 = fetch4(lp, r0 + 4 * ix1);
            const Tex4 a10 = fetch4(lp, r1 + 4 * ix), a11 = fetch4(lp, r1 + 4 * ix1);
            const float u0 = lerp_f32(lerp_f32(a00.x, a00.z, tw), lerp_f32(a01.x, a01.z, tw), tx);
            const float u1 = lerp_f32(lerp_f32(a10.x, a10.z, tw), lerp_f32(a11.x, a11.z, tw), tx);
            const float v0 = lerp_f32(lerp_f32(a00.y, a00.w, tw), lerp_f32(a01.y, a01.w, tw), tx);
            const float v1 = lerp_f32(lerp_f32(a10.y, a10.w, tw), lerp_f32(a11.y, a11.w, tw), tx);
            lu[l] = lerp_f32(u0, u1, ty);
            lvv[l] = lerp_f32(v0, v1, ty);
        }
        if (nl == 2) {
            const float wb = 1.0f - (float)vw.wa;
            ru = lerp_f32(lu[0], lu[1], wb);
            rv = lerp_f32(lvv[0], lvv[1], wb);
        } else {
            ru = lu[0];
            rv = lvv[0];
        }
    }
    if (!finite_f(ru)) ru = g.fallback[0];
    if (!finite_f(rv)) rv = g.fallback[1];
    u = ru;
    v = rv;
}

struct FastMath : SeriesMath {
    static constexpr bool kExactSampler = false;
    OD_HDS void sample_uv(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, float& u, float& v, bool pos_f32,
                          const TileView& tv = TileView()) {
        fast_sample_uv(g, pr, vw, lon, lat, u, v, pos_f32, tv);
    }
};

// FastMath for the hot path of the step kernels: SeriesHot's deferring moves (see there)
struct FastHot : SeriesHot {
    static constexpr bool kExactSampler = false;
    OD_HDS void sample_uv(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, float& u, float& v, bool pos_f32,
                          const TileView& tv = TileView()) {
        fast_sample_uv(g, pr, vw, lon, lat, u, v, pos_f32, tv);
    }
};

#ifndef OD_MAX_CHAIN
#define OD_MAX_CHAIN 2
#endif

struct CurrentStages {
    GroupGeom g;
    PairRef t_start, t_mid, t_end;
};

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
    // drift:current_uncertainty[_uniform] / drift:wind_uncertainty (environment.py:869-891): per get_environment
    // call the reference adds N(0, std) and then U(-std, std) draws to the float32 current (and N(0, std) to the
    // wind).  noise_cur: [stage 0..3][kind 0 normal, 1 uniform][component][n] float64 draws of the legacy generator
    // (already scaled), or NULL; kinds present are flagged in noise_kinds (bit 0 normal, bit 1 uniform).
    const double* noise_cur;
    const double* noise_wind;     // [component][n] normal draws for the wind, or NULL
    int32_t noise_kinds;
    int32_t w_same_grid;          // the vertical-velocity group has the geometry and levels of the current group: share index arithmetic
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
    // Reader priority list for the current (environment.py:613-780): where the first group gives NaN the next one is
    // sampled, and so on; the environment fallback applies after the last.  Only read by the CHAIN instantiations.
    int32_t n_chain, pad_chain_;
    GroupGeom cg[OD_MAX_CHAIN];           // their fallback fields are NaN
    PairRef ct[OD_MAX_CHAIN][3];          // pairs at t, t + dt/2, t + dt
    float chain_fallback[2];
};

// The current sample of a stage.  GENERAL (the reader chain family of kernels) also serves groups on a projected plane
// (od_interp.cuh: sample2_any, exact sampler + vector rotation); the default kernels are only launched for geographic groups.
template <class MATH, bool GENERAL>
OD_HD void sample_cur(const GroupGeom& g, const PairRef& pr, const VertW& vw, double lon, double lat, float& u, float& v, bool pos_f32,
                      const TileView& tv) {
    if (GENERAL && g.proj_kind) {
        sample2_any(g, pr, vw, lon, lat, u, v, pos_f32);
        return;
    }
    MATH::sample_uv(g, pr, vw, lon, lat, u, v, pos_f32, tv);
}

// fill what the groups so far left missing from the next readers of the priority list
template <class MATH>
OD_HD void chain_fill(const StepParams& p, int which, double zt, bool zf32, double lon, double lat, bool pos_f32, float& u, float& v) {
    for (int k = 0; k < p.n_chain; ++k) {
        if (finite_f(u) && finite_f(v)) break;
        const GroupGeom& g = p.cg[k];
        const VertW vw = vert_weights(g, g.zs, g.zy, zt, zf32);
        float a, b;
        sample_cur<MATH, true>(g, p.ct[k][which], vw, lon, lat, a, b, pos_f32, TileView());
        if (!finite_f(u)) u = a;
        if (!finite_f(v)) v = b;
    }
    if (!finite_f(u)) u = p.chain_fallback[0];
    if (!finite_f(v)) v = p.chain_fallback[1];
}

// env[var] += draws  on a float32 array: float32(float64(k) + draw), normal first, then uniform
OD_HD void add_current_noise(const StepParams& p, int stage, int64_t i, float& u, float& v) {
    if (!p.noise_cur) return;
    for (int kind = 0; kind < 2; ++kind) {
        if (!(p.noise_kinds & (1 << kind))) continue;
        const double* base = p.noise_cur + ((int64_t)(stage * 2 + kind) * 2) * p.n;
        u = (float)OD_DADD((double)u, base[i]);
        v = (float)OD_DADD((double)v, base[p.n + i]);
    }
}

// Returns the RK-combined velocity (float32) that the final move uses; k1 is sampled here unless given.
// The stages run as one rolled loop (a single copy of the move and of the sampler in the instruction stream: the
// unrolled form exceeded the SM's instruction cache and spent a fifth of its issue slots waiting for fetches).
//   stage 1..3: position x0 (+) 0.5*dt*k_{stage}; pair t_mid, t_mid, t_end (the reference's stage-4 quirk: half
//   step, end time); RK2 stops after stage 1 and returns k2; RK4 returns (k1 + 2 k2 + 2 k3 + k4) / 6 in float32,
//   accumulated left to right as the reference writes it.
template <int SCHEME, class MATH, bool CHAIN = false>
OD_HD void rk_velocity(const StepParams& p, int64_t i, const VertW& vw, const typename MATH::Start& gs, double lon0, double lat0,
                       float dt32, float k1u, float k1v, float& ou, float& ov, const TileView& tv, bool& bad, double zt = 0.0, bool zf32 = true) {
    const CurrentStages& cs = p.cs;
    if (SCHEME == 0) {
        ou = k1u;
        ov = k1v;
        return;
    }
    float ku = k1u, kv = k1v;           // velocity of the previous stage
    float su = k1u, sv = k1v;           // running RK4 sum
    const int last = SCHEME == 1 ? 1 : 3;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int st = 1; st <= last; ++st) {
        double mlon, mlat;
        do_midpoint<MATH>(gs, lon0, lat0, ku, kv, dt32, mlon, mlat, bad);
        const PairRef& pr = st == 3 ? cs.t_end : cs.t_mid;
        sample_cur<MATH, CHAIN>(cs.g, pr, vw, mlon, mlat, ku, kv, false, tv);
        if (CHAIN) chain_fill<MATH>(p, st == 3 ? 2 : 1, zt, zf32, mlon, mlat, false, ku, kv);
        add_current_noise(p, st, i, ku, kv);
        if (st < 3) {
            su = OD_FADD(su, OD_FMUL(2.0f, ku));
            sv = OD_FADD(sv, OD_FMUL(2.0f, kv));
        } else {
            su = OD_FADD(su, ku);
            sv = OD_FADD(sv, kv);
        }
    }
    if (SCHEME == 1) {
        ou = ku;
        ov = kv;
        return;
    }
    // (x_vel + 2*x_vel2 + 2*x_vel3 + x_vel4)/6.0, float32, left to right
    ou = su / 6.0f;
    ov = sv / 6.0f;
}

// The wind move and the diffusion move of the fused step.  (Measured: making them __noinline__ to keep their code out of
// the hot loop's instruction stream costs far more than it saves -- the calls give the kernel a 1.2 KB stack frame and
// double its run time -- so they are inlined.)
#if defined(__CUDACC__)
#define OD_COLD static __host__ __device__ __forceinline__
#else
#define OD_COLD static
#endif

template <class MATH, bool GENERAL = false>
OD_COLD void extras_wind(const StepParams& p, int64_t i, double z0, double mv, double lon0, double lat0,
                         double* plon1, double* plat1, bool& bad) {
    double lon1 = *plon1, lat1 = *plat1;
    // ---- advect_wind (physics_methods.py:712-791): wind sampled at the start-of-step position
    {
        const VertW v0 = {0, 0, 1.0};
        float xw, yw;
        sample_cur<MATH, GENERAL>(p.gwind, p.pwind, v0, lon0, lat0, xw, yw, p.pos_f32 != 0, TileView());
        if (p.noise_wind) {
            xw = (float)OD_DADD((double)xw, p.noise_wind[i]);
            yw = (float)OD_DADD((double)yw, p.noise_wind[p.n + i]);
        }
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
                const typename MATH::Start g1 = MATH::start(lat1);
                do_move64<MATH>(g1, lon1, lat1, xv, yv, mv, p.dt, lon1, lat1, bad);
            }
        } else {
            float wdf = ((const float*)p.wdf)[i];
            if (!surface) wdf = 0.0f;
            const float xv = OD_FMUL(xw, wdf), yv = OD_FMUL(yw, wdf);
            if (xv != 0.0f || yv != 0.0f) {
                const typename MATH::Start g1 = MATH::start(lat1);
                do_move32<MATH>(g1, lon1, lat1, xv, yv, mv, p.dt, lon1, lat1, bad);
            }
        }
    }
    *plon1 = lon1;
    *plat1 = lat1;
}

template <class MATH>
OD_COLD void extras_diffusion(const StepParams& p, int64_t i, double mv, double* plon1, double* plat1, bool& bad) {
    double lon1 = *plon1, lat1 = *plat1;
    // ---- horizontal_diffusion (basemodel/__init__.py:1746-1772)
    {
        const float D = p.diffusivity ? p.diffusivity[i] : p.diffusivity_const;
        const float s = sqrtf(OD_FMUL(2.0f, D) / p.adt32);
        const double sd = OD_DMUL(mv, (double)s);
        const double xv = OD_DMUL(sd, p.rand_x[i]), yv = OD_DMUL(sd, p.rand_y[i]);
        if (xv != 0.0 || yv != 0.0) {
            const typename MATH::Start g1 = MATH::start(lat1);
            do_move64<MATH>(g1, lon1, lat1, xv, yv, mv, p.dt, lon1, lat1, bad);
        }
    }
    *plon1 = lon1;
    *plat1 = lat1;
}

// One particle, one step (the body of step_kernel; also compiled for the host by tests/hostshim).
// zs/zy and zsw/zyw are the level tables of the current and the vertical-velocity group.
// EXTRAS: 0 current advection only; 1 all extras (wind move, vertical advection, diffusion move); 2 vertical advection only
// Returns true when the particle is done.  false (deferring policies only, MATH::kDefer): one of its moves was outside the
// range of the hot path's geodesic; lon / lat have not been written and the caller redoes the particle with the full policy
// and redo = true (the depth update of vertical advection, which does not depend on the moves, is not applied twice).
template <int SCHEME, bool F64, int EXTRAS, class MATH = ExactMath, bool CHAIN = false>
OD_HD bool step_particle(const StepParams& p, int64_t i, const double* zs, const double* zy,
                         const double* zsw, const double* zyw, const TileView& tv = TileView(), bool redo = false) {
    const GroupGeom& g = p.cs.g;
    const double lon0 = p.lon[i], lat0 = p.lat[i];
    const bool zf32 = p.z_f64 == 0;
    const double z0 = p.z ? (zf32 ? (double)((const float*)p.z)[i] : ((const double*)p.z)[i]) : 0.0;
    double zt = z0;                // drift:truncate_ocean_model_below_m (environment.py:554-562)
    if (p.truncate_below > 0.0 && zt < -p.truncate_below) zt = zf32 ? (double)(float)(-p.truncate_below) : -p.truncate_below;
    const VertW vw = vert_weights(g, zs, zy, zt, zf32);
    const double mv = p.moving ? (double)p.moving[i] : 1.0;
    const typename MATH::Start gs = MATH::start(lat0);

    // stage 1: the start-of-step environment
    float k1u, k1v;
    HorizW h0;                     // horizontal cell / weights of the start-of-step position (exact samplers)
    const bool projected = CHAIN && g.proj_kind != 0;          // (general kernels only)
    const bool share_h0 = EXTRAS != 0 && MATH::kExactSampler && p.w_on && p.w_same_grid && !projected;
    if (share_h0 || (MATH::kExactSampler && !p.has_k1 && !projected)) h0 = horiz_weights(g, lon0, lat0, p.pos_f32 != 0);
    if (p.has_k1) {
        k1u = p.k1u[i];
        k1v = p.k1v[i];
    } else {
        if (projected) sample2_any(g, p.cs.t_start, vw, lon0, lat0, k1u, k1v, p.pos_f32 != 0);
        else if (MATH::kExactSampler) sample2_h(g, p.cs.t_start, vw, h0, k1u, k1v, tv);
        else MATH::sample_uv(g, p.cs.t_start, vw, lon0, lat0, k1u, k1v, p.pos_f32 != 0, tv);
        if (CHAIN) chain_fill<MATH>(p, 0, zt, zf32, lon0, lat0, p.pos_f32 != 0, k1u, k1v);
        add_current_noise(p, 0, i, k1u, k1v);
    }
    if (p.env_u) p.env_u[i] = k1u;
    if (p.env_v) p.env_v[i] = k1v;

    if (EXTRAS) {
        // (done here, next to the stage-1 sample whose cell and weights it can share; it only touches z)
        // ---- vertical_advection (oceandrift.py:315-350): z = min(0, z + moving*w*dt); w sampled at the
        // start-of-step depth, applied to the current depth (which vertical mixing may already have changed)
        if (p.w_on && !redo) {
            const bool zio32 = p.zio_f64 == 0;
            const double zc = zio32 ? (double)((const float*)p.z_inout)[i] : ((const double*)p.z_inout)[i];
            const bool applicable = p.w_at_surface ? (zc <= 0.0) : (zc < 0.0);
            if (applicable) {
                float w;
                if (share_h0) {
                    w = sample1_h(p.gw, p.pw, vw, h0);
                } else {
                    const VertW vww = vert_weights(p.gw, zsw, zyw, zt, zf32);
                    w = CHAIN ? sample1_any(p.gw, p.pw, vww, lon0, lat0, p.pos_f32 != 0)
                              : MATH::sample_s(p.gw, p.pw, vww, lon0, lat0, p.pos_f32 != 0);
                }
                const double zn = fmin(0.0, OD_DADD(zc, OD_DMUL(OD_DMUL(mv, (double)w), p.dt)));
                if (zio32) ((float*)p.z_inout)[i] = (float)zn;
                else ((double*)p.z_inout)[i] = zn;
            }
        }
    }

    float ru, rv;
    bool bad = false;
    rk_velocity<SCHEME, MATH, CHAIN>(p, i, vw, gs, lon0, lat0, p.dt32, k1u, k1v, ru, rv, tv, bad, zt, zf32);

    double lon1, lat1;
    if (F64) {
        const double f = p.factor ? ((const double*)p.factor)[i] : 1.0;
        do_move64<MATH>(gs, lon0, lat0, OD_DMUL((double)ru, f), OD_DMUL((double)rv, f), mv, p.dt, lon1, lat1, bad);
    } else {
        const float f = p.factor ? ((const float*)p.factor)[i] : 1.0f;
        do_move32<MATH>(gs, lon0, lat0, OD_FMUL(ru, f), OD_FMUL(rv, f), mv, p.dt, lon1, lat1, bad);
    }

    if (EXTRAS == 1) {
        if (p.wind_on) extras_wind<MATH, CHAIN>(p, i, z0, mv, lon0, lat0, &lon1, &lat1, bad);
        if (p.diff_on) extras_diffusion<MATH>(p, i, mv, &lon1, &lat1, bad);
    }
    if (MATH::kDefer && bad) return false;
    p.lon[i] = lon1;
    p.lat[i] = lat1;
    return true;
}

// The step as the kernels run it: the hot policy first (SeriesHot for SeriesMath, the policy itself otherwise); a particle
// whose step contained a move the hot path does not solve is redone with the full policy, out of line.
template <class MATH> struct HotPolicy { typedef MATH type; };
template <> struct HotPolicy<SeriesMath> { typedef SeriesHot type; };
template <> struct HotPolicy<FastMath> { typedef FastHot type; };

#if defined(__CUDACC__)
#define OD_NOINLINE static __host__ __device__ __noinline__
#else
#define OD_NOINLINE static
#endif

template <int SCHEME, bool F64, int EXTRAS, class MATH, bool CHAIN>
OD_NOINLINE void step_particle_redo(const StepParams* p, int64_t i, const double* zs, const double* zy, const double* zsw, const double* zyw,
                                    bool depth_done = true) {
    step_particle<SCHEME, F64, EXTRAS, MATH, CHAIN>(*p, i, zs, zy, zsw, zyw, TileView(), depth_done);
}

template <int SCHEME, bool F64, int EXTRAS, class MATH = ExactMath, bool CHAIN = false>
OD_HD void step_particle_full(const StepParams& p, int64_t i, const double* zs, const double* zy,
                              const double* zsw, const double* zyw, const TileView& tv = TileView()) {
    typedef typename HotPolicy<MATH>::type HOT;
    if (!step_particle<SCHEME, F64, EXTRAS, HOT, CHAIN>(p, i, zs, zy, zsw, zyw, tv))
        step_particle_redo<SCHEME, F64, EXTRAS, MATH, CHAIN>(&p, i, zs, zy, zsw, zyw);
}

}  // namespace od
