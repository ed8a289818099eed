// hostshim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the per-particle device math of
// opendrift_b200/csrc/*.cuh for the host (g++, -ffp-contract=off) so that the CPU test-suite can check
// the arithmetic of the CUDA kernels against the oracle without a GPU.  Never loaded by the product
// package (opendrift_b200/_lib.py only ever loads libodcuda.so and fails without it).
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../opendrift_b200/csrc/od_advect.cuh"
#include "../../opendrift_b200/csrc/od_spec.cuh"
#include "../../opendrift_b200/csrc/od_mix.cuh"
#include "../../opendrift_b200/csrc/od_stokes.cuh"
#include "../../opendrift_b200/csrc/od_leeway.cuh"
#include "../../opendrift_b200/csrc/od_analytic.cuh"
#include "../../opendrift_b200/csrc/od_history.cuh"
#include "../../opendrift_b200/csrc/od_bookkeep.cuh"

using namespace od;

extern "C" {

struct hs_group {
    int32_t ncomp, nx, ny, nz, lon_mode, wrap_x, global_x, pad_;
    double x0, xspan, y0, yspan, xmin, xmax, ymin, ymax;
    float fallback[2];
    const double* z_levels;   // as the reader gives them
    od_proj_desc proj;        // kind 0: geographic
    int32_t rotate_vectors, pad2_;
};

struct hs_pair {
    const float* tex;
    int32_t mode, pad_;
    double w;
};

struct hs_levels {
    std::vector<double> zs, zy;
    double zmin = 0, zmax = 0;
};

static GroupGeom make_geom(const hs_group& d, hs_levels& lv) {
    GroupGeom q;
    memset(&q, 0, sizeof(q));
    q.nx = d.nx; q.ny = d.ny; q.nz = d.nz; q.ncomp = d.ncomp; q.lon_mode = d.lon_mode; q.wrap = d.wrap_x ? 1 : 0; q.glob = (d.global_x || d.wrap_x) ? 1 : 0;
    q.x0 = d.x0; q.xspan = d.xspan; q.y0 = d.y0; q.yspan = d.yspan;
    q.xmin = d.xmin; q.xmax = d.xmax; q.ymin = d.ymin; q.ymax = d.ymax;
    q.nxm1 = (double)(d.nx - 1 + q.wrap); q.nym1 = (double)(d.ny - 1);
    q.inv_dx = q.nxm1 / q.xspan; q.inv_dy = q.nym1 / q.yspan;
    q.rxspan = div_rn_reciprocal(q.xspan); q.ryspan = div_rn_reciprocal(q.yspan);
    q.fallback[0] = d.fallback[0]; q.fallback[1] = d.fallback[1];
    if (d.nz > 1) {
        lv.zs.resize(d.nz); lv.zy.resize(d.nz);
        bool inc = d.z_levels[1] > d.z_levels[0];
        for (int i = 0; i < d.nz; ++i) {
            int src = inc ? i : d.nz - 1 - i;
            lv.zs[i] = d.z_levels[src];
            lv.zy[i] = (double)src;
        }
        q.zmin = lv.zs[0]; q.zmax = lv.zs[d.nz - 1];
        q.zs = lv.zs.data(); q.zy = lv.zy.data();
    }
    if (d.proj.kind != 0 && proj_from_desc(&d.proj, &q.proj) == 0) {
        q.proj_kind = d.proj.kind;
        q.rotate = d.rotate_vectors ? 1 : 0;
        q.rot_delta = 10.0;
    }
    return q;
}

static PairRef make_pair(const hs_pair& p) {
    PairRef r;
    r.tex = p.tex; r.mode = p.mode; r.pad_ = 0; r.w = p.w;
    return r;
}

void hs_geod_direct(int64_t n, const double* lon, const double* lat, const double* az, const double* dist,
                    double* lon2, double* lat2) {
    for (int64_t i = 0; i < n; ++i) geod_direct(lon[i], lat[i], az[i], dist[i], lon2[i], lat2[i]);
}

// series_move on (north, east) displacements; used[i] = 1 where the series applied (else the full solution ran)
void hs_geod_series(int64_t n, const double* lon, const double* lat, const double* xn, const double* ye,
                    double* lon2, double* lat2, int32_t* used) {
    for (int64_t i = 0; i < n; ++i) {
        const SeriesStart st = series_start(lat[i]);
        used[i] = series_move(st, lon[i], xn[i], ye[i], lon2[i], lat2[i]) ? 1 : 0;
        if (!used[i]) geod_move_ne(st, lon[i], xn[i], ye[i], lon2[i], lat2[i]);
    }
}

void hs_interp(const hs_group* g, const hs_pair* pr, int64_t n, const double* lon, const double* lat, const float* z,
               int pos_f32, float* out0, float* out1) {
    hs_levels lv;
    GroupGeom q = make_geom(*g, lv);
    PairRef p = make_pair(*pr);
    for (int64_t i = 0; i < n; ++i) {
        VertW vw = vert_weights(q, q.zs, q.zy, (z && q.nz > 1) ? (double)z[i] : 0.0);
        if (q.ncomp == 2) {
            float u, v;
            sample2(q, p, vw, lon[i], lat[i], u, v, pos_f32 != 0);
            out0[i] = u; out1[i] = v;
        } else {
            out0[i] = sample1(q, p, vw, lon[i], lat[i], pos_f32 != 0);
        }
    }
}

struct hs_step_args {
    int32_t scheme, factor_f64, pos_f32, z_f64;
    hs_group g_uv;
    hs_pair t_start, t_mid, t_end;
    double dt;
    int64_t n;
    double* lon; double* lat; const void* z;
    const void* factor; const int32_t* moving;
    double truncate_below;
    // extras
    int32_t wind_on, wdf_f64, w_on, w_at_surface;
    hs_group g_wind; hs_pair t_wind; const void* wdf; double wind_drift_depth;
    hs_group g_w; hs_pair t_w; void* z_inout;
    const double* rand_x; const double* rand_y; const float* diffusivity; float diffusivity_const; int32_t z_inout_f64;
    int32_t fast, noise_kinds;
    const double* noise_cur; const double* noise_wind;
};

}  // extern "C"

template <int S, bool F>
static void run(const StepParams& p, const GroupGeom& gw, int fast = 0) {
    if (fast == 2) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, 1, SeriesMath>(p, i, p.cs.g.zs, p.cs.g.zy, gw.zs, gw.zy);
    else if (fast) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, 1, FastMath>(p, i, p.cs.g.zs, p.cs.g.zy, gw.zs, gw.zy);
    else for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, 1, ExactMath>(p, i, p.cs.g.zs, p.cs.g.zy, gw.zs, gw.zy);
}

extern "C" {

int hs_step(const hs_step_args* a) {
    hs_levels l1, l2, l3;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.cs.g = make_geom(a->g_uv, l1);
    p.cs.t_start = make_pair(a->t_start); p.cs.t_mid = make_pair(a->t_mid); p.cs.t_end = make_pair(a->t_end);
    p.dt = a->dt; p.dt32 = (float)a->dt; p.adt32 = (float)fabs(a->dt);
    p.n = a->n; p.lon = a->lon; p.lat = a->lat; p.z = a->z; p.factor = a->factor; p.moving = a->moving;
    p.truncate_below = a->truncate_below;
    p.pos_f32 = a->pos_f32;
    p.z_f64 = a->z_f64;
    p.noise_cur = a->noise_kinds ? a->noise_cur : nullptr; p.noise_kinds = a->noise_kinds; p.noise_wind = a->noise_wind;
    if (a->wind_on) {
        p.wind_on = 1; p.wdf_f64 = a->wdf_f64; p.gwind = make_geom(a->g_wind, l2); p.pwind = make_pair(a->t_wind);
        p.wdf = a->wdf; p.wind_drift_depth = a->wind_drift_depth;
    }
    if (a->w_on) {
        p.w_on = 1; p.w_at_surface = a->w_at_surface; p.gw = make_geom(a->g_w, l3); p.pw = make_pair(a->t_w);
        p.z_inout = a->z_inout;
        p.zio_f64 = a->z_inout_f64;
        const hs_group &du = a->g_uv, &dw = a->g_w;
        bool same = du.nx == dw.nx && du.ny == dw.ny && du.nz == dw.nz && du.lon_mode == dw.lon_mode && du.wrap_x == dw.wrap_x && du.global_x == dw.global_x &&
                    du.x0 == dw.x0 && du.xspan == dw.xspan && du.y0 == dw.y0 && du.yspan == dw.yspan && du.xmin == dw.xmin &&
                    du.xmax == dw.xmax && du.ymin == dw.ymin && du.ymax == dw.ymax;
        for (int k = 0; same && du.nz > 1 && k < du.nz; ++k) same = du.z_levels[k] == dw.z_levels[k];
        p.w_same_grid = same ? 1 : 0;
    }
    if (a->rand_x) {
        p.diff_on = 1; p.rand_x = a->rand_x; p.rand_y = a->rand_y; p.diffusivity = a->diffusivity;
        p.diffusivity_const = a->diffusivity_const;
    }
    const bool f = a->factor_f64 != 0;
    switch (a->scheme) {
        case 0: f ? run<0, true>(p, p.gw, a->fast) : run<0, false>(p, p.gw, a->fast); break;
        case 1: f ? run<1, true>(p, p.gw, a->fast) : run<1, false>(p, p.gw, a->fast); break;
        case 2: f ? run<2, true>(p, p.gw, a->fast) : run<2, false>(p, p.gw, a->fast); break;
        default: return -2;
    }
    return 0;
}

struct hs_mix_args {
    hs_group g; hs_pair t_k;
    int64_t n; const double* lon; const double* lat; const void* z_in; double* z_out;
    const int32_t* moving; const double* rand; double dt_mix; double sea_floor_const;
    unsigned long long seed; int32_t ntimes, z_in_f64, mix_at_surface, pos_f32, step_index, model;
    int32_t nlev, pad_; const float* wind_speed; double mld_const, background, k_const;
};

int hs_mix(const hs_mix_args* a) {
    hs_levels lv;
    MixParams p;
    memset(&p, 0, sizeof(p));
    std::vector<double> xs, xy, zl;
    if (a->model == 0) {
        p.g = make_geom(a->g, lv);
        p.pr = make_pair(a->t_k);
        const int nz = a->g.nz;
        xs.resize(nz); xy.resize(nz); zl.assign(a->g.z_levels, a->g.z_levels + nz);
        const bool inc = zl[1] > zl[0];
        for (int i = 0; i < nz; ++i) { int src = inc ? nz - 1 - i : i; xs[i] = -zl[src]; xy[i] = (double)src; }
        p.zl = zl.data(); p.xs = xs.data(); p.xy = xy.data();
        p.uniform_dz = 1; p.dz0 = zl[1] - zl[0];
        for (int k = 1; k + 1 < nz; ++k) if (zl[k + 1] - zl[k] != p.dz0) p.uniform_dz = 0;
    } else {
        p.g.nz = a->nlev;
        p.uniform_dz = 1; p.dz0 = -1.0;
        p.wind_speed = a->wind_speed; p.mld_const = (float)a->mld_const; p.background = a->background; p.k_const = a->k_const;
    }
    p.model = a->model;
    p.n = a->n; p.lon = a->lon; p.lat = a->lat; p.z_in = a->z_in; p.z_out = a->z_out; p.moving = a->moving;
    p.rand = a->rand; p.dt_mix = a->dt_mix; p.zmin_const = -(double)(float)a->sea_floor_const; p.seed = a->seed;
    p.ntimes = a->ntimes; p.z_in_f64 = a->z_in_f64; p.mix_at_surface = a->mix_at_surface; p.pos_f32 = a->pos_f32;
    p.step_index = a->step_index;
    for (int64_t i = 0; i < a->n; ++i) mix_particle(p, i, p.xs, p.xy);
    return 0;
}

// the diffusivity column a particle sees (k_level of od_mix.cuh), all levels: out[n][nz]
void hs_kcolumn(const hs_group* g, const hs_pair* pr, int64_t n, const double* lon, const double* lat, int pos_f32, double* out) {
    hs_levels lv;
    MixParams p;
    memset(&p, 0, sizeof(p));
    p.g = make_geom(*g, lv);
    p.pr = make_pair(*pr);
    for (int64_t i = 0; i < n; ++i) {
        const HorizW h = horiz_weights(p.g, lon[i], lat[i], pos_f32 != 0);
        for (int l = 0; l < p.g.nz; ++l) out[i * p.g.nz + l] = k_level(p, h, l);
    }
}

struct hs_stokes_args {
    int64_t n; double* lon; double* lat; const void* z; const float* us; const float* vs; const float* hs;
    const float* xwind; const float* ywind; const int32_t* moving; double dt; int32_t z_f64, hs_mode, profile, pad_;
    double factor; const void* d_factor; int32_t factor_f64, pad2_;
    const float* sw_dir; const float* sw_period; const float* sw_hs; const float* ws_dir; const float* ws_period; const float* ws_hs;
};

int hs_stokes(const hs_stokes_args* a) {
    StokesParams p;
    memset(&p, 0, sizeof(p));
    p.n = a->n; p.lon = a->lon; p.lat = a->lat; p.z = a->z; p.us = a->us; p.vs = a->vs; p.hs = a->hs;
    p.xwind = a->xwind; p.ywind = a->ywind; p.moving = a->moving; p.dt = a->dt; p.z_f64 = a->z_f64;
    p.hs_mode = a->hs_mode; p.profile = a->profile;
    p.factor = a->factor; p.factor_arr = a->d_factor; p.factor_f64 = a->factor_f64;
    p.sw_dir = a->sw_dir; p.sw_period = a->sw_period; p.sw_hs = a->sw_hs; p.ws_dir = a->ws_dir; p.ws_period = a->ws_period; p.ws_hs = a->ws_hs;
    for (int64_t i = 0; i < a->n; ++i) stokes_particle(p, i);
    return 0;
}

struct hs_leeway_args {
    hs_group g_wind, g_cur; hs_pair t_wind, t_cur;
    int64_t n; double* lon; double* lat; const float* dw_slope; const float* dw_offset; const float* dw_eps;
    float* cw_slope; const float* cw_offset; const float* cw_eps; uint8_t* orientation; const double* jibe_probability;
    const int32_t* moving; const double* rand; double dt; float capsize_fraction; int32_t pos_f32;
    uint8_t* capsized; const double* rand_capsize; int32_t capsize_on, capsize_from; float wind_threshold, wind_sigma;
};

int hs_leeway(const hs_leeway_args* a) {
    hs_levels l1, l2;
    LeewayParams p;
    memset(&p, 0, sizeof(p));
    p.gwind = make_geom(a->g_wind, l1); p.gcur = make_geom(a->g_cur, l2);
    p.pwind = make_pair(a->t_wind); p.pcur = make_pair(a->t_cur);
    p.n = a->n; p.lon = a->lon; p.lat = a->lat; p.dw_slope = a->dw_slope; p.dw_offset = a->dw_offset; p.dw_eps = a->dw_eps;
    p.cw_slope = a->cw_slope; p.cw_offset = a->cw_offset; p.cw_eps = a->cw_eps; p.orientation = a->orientation;
    p.jibe_probability = a->jibe_probability; p.jp_f64 = 1; p.moving = a->moving; p.rand = a->rand; p.dt = a->dt;
    p.capsize_fraction = a->capsize_fraction; p.pos_f32 = a->pos_f32;
    p.capsized = a->capsized; p.rand_capsize = a->rand_capsize; p.capsize_on = a->capsize_on; p.capsize_from = a->capsize_from;
    p.wind_threshold = a->wind_threshold; p.wind_sigma = a->wind_sigma;
    for (int64_t i = 0; i < a->n; ++i) leeway_particle(p, i);
    return 0;
}

// analytical reader on a projected plane (od_analytic.cuh), same descriptor and argument structs as the C-ABI
int hs_analytic_interp(const od_analytic_desc* r, double t, int64_t n, const double* lon, const double* lat, int flags,
                       float* u, float* v) {
    AnalyticReader R;
    int rc = analytic_from_desc(r, &R);
    if (rc) return -rc;
    for (int64_t i = 0; i < n; ++i) analytic_sample_raw(R, t, lon[i], lat[i], (flags & 1) != 0, u[i], v[i]);
    return 0;
}

// any od_proj_desc: forward (degrees -> metres; INFINITY where undefined) or inverse (metres -> degrees); returns proj_from_desc's status
int hs_proj(const od_proj_desc* d, int inverse, int64_t n, const double* a, const double* b, double* oa, double* ob) {
    ProjStere P;
    const int rc = proj_from_desc(d, &P);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (inverse) stere_inverse(P, a[i], b[i], oa[i], ob[i]);
        else if (!stere_forward(P, a[i], b[i], oa[i], ob[i])) oa[i] = ob[i] = INFINITY;
    }
    return 0;
}

void hs_stere(const od_analytic_desc* r, int inverse, int64_t n, const double* a, const double* b, double* oa, double* ob) {
    AnalyticReader R;
    if (analytic_from_desc(r, &R)) return;
    for (int64_t i = 0; i < n; ++i) {
        if (inverse) stere_inverse(R.proj, a[i], b[i], oa[i], ob[i]);
        else if (!stere_forward(R.proj, a[i], b[i], oa[i], ob[i])) oa[i] = ob[i] = INFINITY;
    }
}

// od_update_positions (update_positions_kernel of od_kernels.cu): the full Karney move with float32 or float64 velocities
void hs_update_positions(int64_t n, double* lon, double* lat, const void* xv, const void* yv, int vel_f64, const int32_t* moving,
                         double dt) {
    for (int64_t i = 0; i < n; ++i) {
        const double lon0 = lon[i], lat0 = lat[i];
        const double mv = moving ? (double)moving[i] : 1.0;
        const GeodStart gs = geod_start(lat0);
        double lo, la;
        if (vel_f64) final_move_f64(gs, lon0, ((const double*)xv)[i], ((const double*)yv)[i], mv, dt, lo, la);
        else final_move_f32(gs, lon0, ((const float*)xv)[i], ((const float*)yv)[i], mv, dt, lo, la);
        lon[i] = lo;
        lat[i] = la;
    }
}

void hs_minmax_f32(int64_t n, const float* a, const float* b, float* lo, float* hi) {
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        const float v = b ? a[i] + b[i] : a[i];
        if (v != v) continue;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    if (mn <= mx) {         // untouched when every value is NaN, like od_minmax_f32
        *lo = mn;
        *hi = mx;
    }
}

int hs_history_scatter(const od_history_args* a) {
    HistoryParams p;
    p.n = a->n; p.n_total = a->n_total; p.col = a->col; p.ncols = a->ncols; p.z_f64 = a->z_f64; p.pad_ = 0;
    p.ids = a->d_ids; p.lon = a->d_lon; p.lat = a->d_lat; p.z = a->d_z; p.status = a->d_status;
    p.blon = a->d_buf_lon; p.blat = a->d_buf_lat; p.bz = a->d_buf_z; p.bstatus = a->d_buf_status;
    if (a->ncols <= 0 || a->col < 0 || a->col >= a->ncols) return -1;
    for (int64_t i = 0; i < p.n; ++i) history_scatter_one(p, i);
    return 0;
}

int hs_vertical_buoyancy(const od_buoyancy_args* a) {
    BuoyancyParams p;
    p.n = a->n; p.z_in = a->d_z_in; p.z_out = a->d_z_out; p.tv = a->d_terminal_velocity; p.sea_floor = a->d_sea_floor;
    p.status = a->d_status; p.moving = a->d_moving; p.counter = nullptr; p.dt = a->dt; p.ssh = a->sea_surface_height;
    p.z_f64 = a->z_f64; p.tv_f64 = a->tv_f64; p.seafloor_code = a->seafloor_code;
    int64_t c = 0;
    for (int64_t i = 0; i < p.n; ++i) c += buoyancy_one(p, i) ? 1 : 0;
    if (a->h_n_deactivated) *a->h_n_deactivated = c;
    return 0;
}

int hs_bookkeeping(const od_bookkeep_args* a) {
    BookkeepParams p;
    p.n = a->n; p.lon = a->d_lon; p.lat = a->d_lat; p.z = a->d_z; p.age = a->d_age; p.status = a->d_status; p.moving = a->d_moving;
    p.ids = a->d_ids; p.counters = nullptr; p.dt_age = a->dt_age; p.max_age = a->max_age;
    p.west = a->west; p.east = a->east; p.south = a->south; p.north = a->north;
    p.outside_code = a->outside_code; p.retired_code = a->retired_code; p.z_f64 = a->z_f64; p.age_f64 = a->age_f64;
    p.pos_f32 = a->pos_f32; p.only_deactivated = a->only_deactivated;
    p.n_total = a->n_total; p.col = a->col; p.ncols = a->ncols;
    p.blon = a->d_buf_lon; p.blat = a->d_buf_lat; p.bz = a->d_buf_z; p.bstatus = a->d_buf_status;
    int64_t c[3] = {0, 0, 0};
    for (int64_t i = 0; i < p.n; ++i) {
        const int f = bookkeep_one(p, i);
        c[0] += f & 1; c[1] += (f >> 1) & 1; c[2] += (f >> 2) & 1;
    }
    if (a->h_counts) for (int k = 0; k < 3; ++k) a->h_counts[k] = c[k];
    return 0;
}

void hs_inverse_azimuth(int64_t n, const double* lon1, const double* lat1, const double* lon2, const double* lat2, double* az) {
    for (int64_t i = 0; i < n; ++i) az[i] = inverse_azimuth_short(lon1[i], lat1[i], lon2[i], lat2[i]);
}

}  // extern "C"

template <class MATH>
static void run_analytic(const AnalyticStepParams& p, int scheme, bool f64) {
    for (int64_t i = 0; i < p.n; ++i) {
        if (scheme == 0) { if (f64) analytic_step_particle<0, true, MATH>(p, i); else analytic_step_particle<0, false, MATH>(p, i); }
        else if (scheme == 1) { if (f64) analytic_step_particle<1, true, MATH>(p, i); else analytic_step_particle<1, false, MATH>(p, i); }
        else { if (f64) analytic_step_particle<2, true, MATH>(p, i); else analytic_step_particle<2, false, MATH>(p, i); }
    }
}

extern "C" {

int hs_analytic_advect(const od_analytic_desc* r, const od_analytic_advect_args* a) {
    AnalyticStepParams p;
    memset(&p, 0, sizeof(p));
    int rc = analytic_from_desc(r, &p.R);
    if (rc) return -rc;
    p.t_start = a->t_start; p.t_mid = a->t_mid; p.t_end = a->t_end;
    p.dt = a->dt; p.dt32 = (float)a->dt;
    p.has_k1 = a->d_k1_u != nullptr; p.pos_f32 = a->pos_f32; p.n = a->n;
    p.lon = a->d_lon; p.lat = a->d_lat; p.factor = a->d_factor; p.moving = a->d_moving;
    p.k1u = a->d_k1_u; p.k1v = a->d_k1_v; p.env_u = a->d_env_u; p.env_v = a->d_env_v;
    if (a->math == OD_MATH_FAST) run_analytic<FastMath>(p, a->scheme, a->factor_f64 != 0);
    else if (a->math == OD_MATH_SERIES) run_analytic<SeriesMath>(p, a->scheme, a->factor_f64 != 0);
    else run_analytic<ExactMath>(p, a->scheme, a->factor_f64 != 0);
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// hs2_*: the library's own argument structs (include/odcuda.h) + the field groups / time pairs they name, resolved by
// the caller (tests/hostengine.py keeps the slab ring and builds the pair texels).  The parameter blocks are filled the
// way od_kernels.cu fills them (fill_current / fill_step / od_vertical_mixing / od_leeway_step), so that the drop-in
// classes can run end to end on the host build.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int hs2_interp(const hs_group* g, const hs_pair* pr, int64_t n, const double* lon, const double* lat, const void* z, int flags,
               void* out0, void* out1) {
    hs_levels lv;
    GroupGeom q = make_geom(*g, lv);
    const bool f64 = (flags & OD_INTERP_OUT_F64) != 0, nearest = (flags & OD_INTERP_NEAREST) != 0;
    if (f64 && !nearest && !(flags & OD_INTERP_NO_FALLBACK)) return -7;
    if (nearest && (q.ncomp != 1 || q.nz > 1 || q.proj_kind != 0 || q.wrap != 0 || !(q.xspan > 0.0) || !(q.yspan > 0.0))) return -8;
    if (flags & OD_INTERP_NO_FALLBACK) q.fallback[0] = q.fallback[1] = NAN;
    if (flags & OD_INTERP_NO_ROTATE) q.rotate = 0;
    const PairRef p = make_pair(*pr);
    const bool z64 = (flags & OD_INTERP_Z_F64) != 0, p32 = (flags & OD_INTERP_POS_F32) != 0;
    for (int64_t i = 0; i < n; ++i) {
        const double zz = (z && q.nz > 1) ? (z64 ? ((const double*)z)[i] : (double)((const float*)z)[i]) : 0.0;
        const VertW vw = vert_weights(q, q.zs, q.zy, zz, !z64);
        if (nearest) {
            const float r = sample1_nearest(q, p, lon[i], lat[i], p32);
            if (out0) { if (f64) ((double*)out0)[i] = (double)r; else ((float*)out0)[i] = r; }
        } else if (f64) {
            if (q.ncomp == 2) {
                double u, v;
                sample2_any_d(q, p, vw, lon[i], lat[i], u, v, p32);
                if (out0) ((double*)out0)[i] = u;
                if (out1) ((double*)out1)[i] = v;
            } else if (out0) {
                ((double*)out0)[i] = sample1_any_d(q, p, vw, lon[i], lat[i], p32);
            }
        } else if (q.ncomp == 2) {
            float u, v;
            sample2_any(q, p, vw, lon[i], lat[i], u, v, p32);
            if (out0) ((float*)out0)[i] = u;
            if (out1) ((float*)out1)[i] = v;
        } else {
            const float r = sample1_any(q, p, vw, lon[i], lat[i], p32);
            if (out0) ((float*)out0)[i] = r;
        }
    }
    return 0;
}

int hs2_coastline(const od_coast_args* a) {
    if (a->action < 1 || a->action > 3) return -2;
    unsigned c[4] = {0, 0, 0, 0};
    CoastParams p;
    p.n = a->n; p.mask = a->d_mask; p.lon = a->d_lon; p.lat = a->d_lat; p.z = a->d_z; p.age = a->d_age; p.status = a->d_status;
    p.moving = a->d_moving; p.ids = a->d_ids; p.prev_lon = a->d_prev_lon; p.prev_lat = a->d_prev_lat; p.counters = c;
    p.n_total = a->n_total; p.id_base = a->id_base; p.action = a->action; p.ssh = a->ssh; p.stranded_code = a->stranded_code;
    p.seeded_code = a->seeded_code; p.missing_code = a->missing_code; p.check_seeded = a->check_seeded; p.z_f64 = a->z_f64; p.age_f64 = a->age_f64;
    for (int64_t i = 0; i < a->n; ++i) {
        const int f = coast_one(p, i);
        for (int b = 0; b < 4; ++b) c[b] += (f >> b) & 1;
    }
    if (a->h_counts) for (int b = 0; b < 4; ++b) a->h_counts[b] = c[b];
    return 0;
}

int hs2_store_previous(int64_t n, const double* lon, const double* lat, const int32_t* ids, int32_t id_base, int64_t n_total,
                       float* prev_lon, float* prev_lat) {
    for (int64_t i = 0; i < n; ++i) store_previous_one(i, lon, lat, ids, id_base, n_total, prev_lon, prev_lat);
    return 0;
}

}  // extern "C"

static bool same_grid(const hs_group& du, const hs_group& dw) {
    bool same = du.nx == dw.nx && du.ny == dw.ny && du.nz == dw.nz && du.lon_mode == dw.lon_mode && du.wrap_x == dw.wrap_x &&
                du.global_x == dw.global_x && du.x0 == dw.x0 && du.xspan == dw.xspan && du.y0 == dw.y0 && du.yspan == dw.yspan &&
                du.xmin == dw.xmin && du.xmax == dw.xmax && du.ymin == dw.ymin && du.ymax == dw.ymax;
    for (int k = 0; same && du.nz > 1 && k < du.nz; ++k) same = du.z_levels[k] == dw.z_levels[k];
    return same;
}

struct hs_chain {             // the chained current groups of od_advect_args (resolved by the caller)
    const hs_group* g[OD_MAX_CHAIN];
    hs_pair t[OD_MAX_CHAIN][3];
};

static void fill_chain2(const od_advect_args* a, const hs_chain* c, StepParams* p, hs_levels* lv) {
    p->n_chain = a->n_chain;
    for (int k = 0; k < a->n_chain; ++k) {
        p->cg[k] = make_geom(*c->g[k], lv[k]);
        p->cg[k].fallback[0] = p->cg[k].fallback[1] = NAN;
        for (int w = 0; w < 3; ++w) p->ct[k][w] = make_pair(c->t[k][w]);
    }
    if (a->n_chain > 0) {
        p->chain_fallback[0] = p->cs.g.fallback[0];
        p->chain_fallback[1] = p->cs.g.fallback[1];
        p->cs.g.fallback[0] = p->cs.g.fallback[1] = NAN;
    }
}

static int fill_current2(const od_advect_args* a, const hs_group* g, const hs_pair* t3, StepParams* p, hs_levels& lv) {
    if (a->scheme < 0 || a->scheme > 2) return -2;
    if (g->nz > 1 && !a->d_z) return -3;
    memset(p, 0, sizeof(*p));
    p->cs.g = make_geom(*g, lv);
    p->has_k1 = a->d_k1_u != nullptr;
    if (!p->has_k1) p->cs.t_start = make_pair(t3[0]);
    if (a->scheme != 0) p->cs.t_mid = make_pair(t3[1]);
    if (a->scheme == 2) p->cs.t_end = make_pair(t3[2]);
    p->dt = a->dt; p->dt32 = (float)a->dt; p->adt32 = (float)fabs(a->dt);
    p->n = a->n; p->lon = a->d_lon; p->lat = a->d_lat; p->z = a->d_z;
    p->factor = a->d_factor; p->moving = a->d_moving;
    p->k1u = a->d_k1_u; p->k1v = a->d_k1_v; p->env_u = a->d_env_u; p->env_v = a->d_env_v;
    p->truncate_below = a->truncate_below; p->pos_f32 = a->pos_f32; p->z_f64 = a->z_f64;
    p->noise_cur = a->noise_kinds ? a->d_noise_cur : nullptr;
    p->noise_kinds = a->d_noise_cur ? a->noise_kinds : 0;
    return 0;
}

template <int S, bool F, int E>
static void run2c(const StepParams& p, int mode) {          // step_chain_kernel
    const double* zs = p.cs.g.zs; const double* zy = p.cs.g.zy;
    if (mode == OD_MATH_SERIES) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, SeriesMath, true>(p, i, zs, zy, p.gw.zs, p.gw.zy);
    else if (mode == OD_MATH_FAST) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, FastMath, true>(p, i, zs, zy, p.gw.zs, p.gw.zy);
    else for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, ExactMath, true>(p, i, zs, zy, p.gw.zs, p.gw.zy);
}

// the specialised step (od_spec.cuh) as launch_step of od_kernels.cu selects it; hs_spec_mode(0) keeps the general step,
// hs_spec_counts() tells how many particles took it and how many of those were flagged and redone
static int g_spec_on = 1;
static int64_t g_spec_n = 0, g_spec_redo = 0;

template <int S, bool F, int E>
static void run2(const StepParams& p, int mode) {
    const bool general = p.n_chain > 0 || p.cs.g.proj_kind != 0 || (E != 0 && ((p.wind_on && p.gwind.proj_kind != 0) || (p.w_on && p.gw.proj_kind != 0)));
    if (general) { run2c<S, F, (E == 0 ? 0 : 1)>(p, mode); return; }
    const double* zs = p.cs.g.zs; const double* zy = p.cs.g.zy;
    if (S == 2 && F && mode == OD_MATH_SERIES && g_spec_on && spec_eligible(p, S)) {
        for (int64_t i = 0; i < p.n; ++i) {
            const int rc = spec_all_lerp(p) ? step_particle_spec<2, true, E, true>(p, i, zs, zy, p.gw.zs, p.gw.zy)
                                            : step_particle_spec<2, true, E, false>(p, i, zs, zy, p.gw.zs, p.gw.zy);
            if (rc) { step_particle_redo<2, true, E, SeriesMath, false>(&p, i, zs, zy, p.gw.zs, p.gw.zy, rc == 2); ++g_spec_redo; }
        }
        g_spec_n += p.n;
        return;
    }
    if (mode == OD_MATH_SERIES) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, SeriesMath>(p, i, zs, zy, p.gw.zs, p.gw.zy);
    else if (mode == OD_MATH_FAST) for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, FastMath>(p, i, zs, zy, p.gw.zs, p.gw.zy);
    else for (int64_t i = 0; i < p.n; ++i) step_particle_full<S, F, E, ExactMath>(p, i, zs, zy, p.gw.zs, p.gw.zy);
}

template <int E>
static int launch2(const StepParams& p, int scheme, bool f, int mode) {
    switch (scheme) {
        case 0: f ? run2<0, true, E>(p, mode) : run2<0, false, E>(p, mode); return 0;
        case 1: f ? run2<1, true, E>(p, mode) : run2<1, false, E>(p, mode); return 0;
        case 2: f ? run2<2, true, E>(p, mode) : run2<2, false, E>(p, mode); return 0;
    }
    return -2;
}

extern "C" {

void hs_spec_mode(int on) { g_spec_on = on; }
void hs_spec_counts(int64_t* n, int64_t* redo, int reset) {
    if (n) *n = g_spec_n;
    if (redo) *redo = g_spec_redo;
    if (reset) g_spec_n = g_spec_redo = 0;
}

int hs2_advect(const od_advect_args* a, const hs_group* g, const hs_pair* t3, const hs_chain* chain) {
    hs_levels lv, lc[OD_MAX_CHAIN];
    StepParams p;
    int rc = fill_current2(a, g, t3, &p, lv);
    if (rc) return rc;
    if (a->n_chain > 0) fill_chain2(a, chain, &p, lc);
    return launch2<0>(p, a->scheme, a->factor_f64 != 0, a->fast);
}

int hs2_step(const od_step_args* a, const hs_group* g_uv, const hs_pair* t3, const hs_group* g_wind, const hs_pair* t_wind,
             const hs_group* g_w, const hs_pair* t_w, const hs_chain* chain) {
    hs_levels l1, l2, l3, lc[OD_MAX_CHAIN];
    StepParams p;
    int rc = fill_current2(&a->cur, g_uv, t3, &p, l1);
    if (rc) return rc;
    if (a->cur.n_chain > 0) fill_chain2(&a->cur, chain, &p, lc);
    if (a->group_wind >= 0) {
        if (!a->d_wdf || !g_wind || g_wind->nz != 1) return -4;
        p.wind_on = 1; p.wdf_f64 = a->wdf_f64; p.gwind = make_geom(*g_wind, l2); p.pwind = make_pair(*t_wind);
        p.wdf = a->d_wdf; p.wind_drift_depth = a->wind_drift_depth; p.noise_wind = a->d_noise_wind;
    }
    if (a->group_w >= 0) {
        if (!a->d_z_inout || !a->cur.d_z || !g_w) return -5;
        p.w_on = 1; p.w_at_surface = a->w_at_surface; p.gw = make_geom(*g_w, l3); p.pw = make_pair(*t_w);
        p.z_inout = a->d_z_inout; p.zio_f64 = a->z_inout_f64;
        p.w_same_grid = same_grid(*g_uv, *g_w) ? 1 : 0;
    }
    if (a->d_rand_x) {
        if (!a->d_rand_y) return -6;
        p.diff_on = 1; p.rand_x = a->d_rand_x; p.rand_y = a->d_rand_y; p.diffusivity = a->d_diffusivity;
        p.diffusivity_const = a->diffusivity_const;
    }
    if (!p.wind_on && !p.diff_on) return launch2<2>(p, a->cur.scheme, a->cur.factor_f64 != 0, a->cur.fast);
    return launch2<1>(p, a->cur.scheme, a->cur.factor_f64 != 0, a->cur.fast);
}

int hs2_mix(const od_mix_args* a, const hs_group* g, const hs_pair* pr) {
    hs_levels lv;
    MixParams p;
    memset(&p, 0, sizeof(p));
    std::vector<double> xs, xy, zl;
    if (a->model == OD_MIX_ENVIRONMENT) {
        if (!g || g->nz < 2) return -2;
        p.g = make_geom(*g, lv);
        p.pr = make_pair(*pr);
        const int nz = g->nz;
        xs.resize(nz); xy.resize(nz); zl.assign(g->z_levels, g->z_levels + nz);
        const bool inc = zl[1] > zl[0];
        for (int i = 0; i < nz; ++i) { int src = inc ? nz - 1 - i : i; xs[i] = -zl[src]; xy[i] = (double)src; }
        p.zl = zl.data(); p.xs = xs.data(); p.xy = xy.data();
        p.uniform_dz = 1; p.dz0 = zl[1] - zl[0];
        for (int k = 1; k + 1 < nz; ++k) if (zl[k + 1] - zl[k] != p.dz0) p.uniform_dz = 0;
    } else {
        if (a->nlev < 2 || a->nlev > 65535) return -3;
        if (a->model != OD_MIX_CONSTANT && !a->d_wind_speed) return -4;
        p.g.nz = a->nlev; p.uniform_dz = 1; p.dz0 = -1.0;
        p.wind_speed = a->d_wind_speed; p.mld = a->d_mld; p.mld_const = (float)a->mld_const;
        p.background = a->background; p.k_const = a->k_const;
    }
    p.model = a->model;
    p.n = a->n; p.lon = a->d_lon; p.lat = a->d_lat; p.z_in = a->d_z_in; p.z_out = a->d_z_out;
    p.moving = a->d_moving; p.terminal_velocity = a->d_terminal_velocity; p.ids = a->d_ids; p.rand = a->d_rand;
    p.dt_mix = a->dt_mix; p.zmin_const = -(double)(float)a->sea_floor_const; p.sea_floor = a->d_sea_floor;
    p.seed = a->seed; p.ntimes = a->ntimes; p.z_in_f64 = a->z_in_f64; p.tv_f64 = a->tv_f64;
    p.mix_at_surface = a->mix_at_surface; p.pos_f32 = a->pos_f32; p.step_index = a->step_index;
    p.seafloor_action = a->seafloor_action; p.seafloor_code = a->seafloor_code; p.status = a->d_status; p.moving_out = a->d_moving_out;
    p.iter0 = a->iter0; p.skip_surface_stick = a->skip_surface_stick;
    unsigned cnt = 0;
    p.counter = &cnt;
    if (p.g.proj_kind) for (int64_t i = 0; i < a->n; ++i) mix_particle<true>(p, i, p.xs, p.xy);
    else for (int64_t i = 0; i < a->n; ++i) mix_particle<false>(p, i, p.xs, p.xy);
    if (a->h_n_deactivated) *a->h_n_deactivated = cnt;
    return 0;
}

int hs2_leeway(const od_leeway_args* a, const hs_group* g_wind, const hs_pair* t_wind, const hs_group* g_cur, const hs_pair* t_cur) {
    hs_levels l1, l2;
    LeewayParams p;
    memset(&p, 0, sizeof(p));
    if (g_wind->nz != 1 || g_cur->nz != 1) return -2;
    p.gwind = make_geom(*g_wind, l1); p.gcur = make_geom(*g_cur, l2);
    p.pwind = make_pair(*t_wind); p.pcur = make_pair(*t_cur);
    p.n = a->n; p.lon = a->d_lon; p.lat = a->d_lat;
    p.dw_slope = a->d_dw_slope; p.dw_offset = a->d_dw_offset; p.dw_eps = a->d_dw_eps;
    p.cw_slope = a->d_cw_slope; p.cw_offset = a->d_cw_offset; p.cw_eps = a->d_cw_eps;
    p.orientation = a->d_orientation; p.capsized = a->d_capsized; p.jibe_probability = a->d_jibe_probability;
    p.moving = a->d_moving; p.status = a->d_status; p.ids = a->d_ids; p.rand = a->d_rand; p.dt = a->dt; p.seed = a->seed;
    p.capsize_fraction = a->capsize_fraction; p.jp_f64 = a->jp_f64; p.pos_f32 = a->pos_f32; p.step_index = a->step_index;
    p.capsize_on = a->capsize_on; p.capsize_from = a->capsize_from; p.wind_threshold = a->wind_threshold;
    p.wind_sigma = a->wind_sigma; p.rand_capsize = a->d_rand_capsize;
    p.noise_cur = a->d_noise_cur; p.noise_wind = a->d_noise_wind; p.noise_kinds = a->noise_kinds;
    if (a->capsize_on && !a->d_capsized) return -3;
    p.missing_code = a->missing_code;
    if (p.gwind.proj_kind || p.gcur.proj_kind) for (int64_t i = 0; i < a->n; ++i) leeway_particle<true>(p, i);
    else for (int64_t i = 0; i < a->n; ++i) leeway_particle<false>(p, i);
    return 0;
}

}  // extern "C"
