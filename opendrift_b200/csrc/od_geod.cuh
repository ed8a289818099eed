// od_geod.cuh -- WGS84 direct geodesic (Karney 2013, order-6 series), float64, one call per particle.
//
// Replaces pyproj.Geod(ellps='WGS84').fwd, which the reference calls for every position update
// (opendrift/models/basemodel/__init__.py:4643-4657) and every Runge-Kutta mid-point
// (opendrift/models/physics_methods.py:632-635, 649-652, 663-666).  The algorithm is the published one
// (J. Geodesy 87:43-55, eqs. 7-21).  It is written for a GPU thread whose bottleneck is the FP64 pipe:
//   * the part that depends only on the start latitude is split off (all four moves of an RK4 step start
//     from the same point);
//   * sin/cos of angles that are reduced to [-pi/4, pi/4] by construction (azimuth and latitude in degrees)
//     or tiny by nature (B11, and tau12/sig12 for step-sized distances) use the two minimax kernels directly,
//     with no range reduction and no slow path; larger arguments fall back to sincos();
//   * eps(k2) = (sqrt(1+k2)-1)/(sqrt(1+k2)+1) is a 7-term series in k2 <= e'^2 = 0.0067 (exact to round-off),
//     norm(ssig1, csig1) uses the identity ssig1^2 + csig1^2 = calp0^2, the two divisions of
//     tau12 = s12 / (b (1 + A1m1)) are folded into one, divisions by constants are multiplications;
//   * long-mantissa constants live in __constant__ memory so that DFMA/DMUL take them as c[bank][offset]
//     operands instead of materialising them through uniform-register moves;
//   * no back azimuth / reduced length / geodesic scale is computed.
//   * the series are truncated where the dropped terms are below double round-off for the WGS84 flattening
//     (C1 to order 4, C1' to 5, C3 to 3); latitude is returned as lat1 + (small arctangent);
// Accuracy is unchanged: <= 1e-13 deg against the mpmath evaluation of the exact geodesic integrals
// (tests/golden/geod_mpmath.npz), on the host build and on the GPU.
//
// The same source compiles for the host (tests/hostshim) so that the arithmetic can be checked against
// the oracle without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OD_HD __host__ __device__ __forceinline__
#define OD_HDS static __host__ __device__ __forceinline__
#else
#define OD_HD static inline
#define OD_HDS static inline
#endif

namespace od {

struct Wgs84 {
    static constexpr double a = 6378137.0;
    static constexpr double f = 1.0 / 298.257223563;
    static constexpr double f1 = 1.0 - f;
    static constexpr double e2 = f * (2.0 - f);
    static constexpr double ep2 = e2 / (f1 * f1);
    static constexpr double n = f / (2.0 - f);
    static constexpr double b = a * f1;
};

constexpr double kDeg = 0.017453292519943295769;       // pi / 180
constexpr double kRad2Deg = 57.295779513082320877;     // 180 / pi
constexpr double kTiny = 1.4916681462400413e-154;      // sqrt(DBL_MIN)
constexpr double kPio4 = 0.78539816339744830962;

// Every constant with a long mantissa that the geodesic needs.
struct GeodK {
    double ep2, f1, neg_f, inv_b, inv_c, deg, rad2deg, inv90, inv360;
    double S[6], C[6];        // sin / cos minimax kernels on [-pi/4, pi/4]
    double c1[6][3];          // C1[l]  / eps^l : polynomial in eps^2 (highest power first), pre-divided
    double c1p[6][3];         // C1'[l] / eps^l
    double a3[6];             // A3 = sum a3[k] eps^k
    double c3[5][5];          // C3[l] / eps^l = sum_k c3[l-1][k] eps^k
    double at[6];             // atan(t)/t - 1 series in t^2 for |t| <= 0.01
    double ts[3], tc[3];      // Taylor coefficients of sin (x^3, x^5, x^7) and cos (x^4, x^6, x^8)
};

constexpr GeodK make_geodk() {
    GeodK k = {};
    const double n = Wgs84::n;
    k.ep2 = Wgs84::ep2; k.f1 = Wgs84::f1; k.neg_f = -Wgs84::f; k.inv_b = 1.0 / Wgs84::b;
    k.inv_c = Wgs84::b / (Wgs84::a * Wgs84::a);
    k.deg = kDeg; k.rad2deg = kRad2Deg; k.inv90 = 1.0 / 90.0; k.inv360 = 1.0 / 360.0;
    // minimax kernels of sin and cos on [-pi/4, pi/4] (the classic fdlibm k_sin / k_cos coefficients)
    k.S[0] = -1.66666666666666324348e-01; k.S[1] = 8.33333333332248946124e-03; k.S[2] = -1.98412698298579493134e-04;
    k.S[3] = 2.75573137070700676789e-06; k.S[4] = -2.50507602534068634195e-08; k.S[5] = 1.58969099521155010221e-10;
    k.C[0] = 4.16666666666666019037e-02; k.C[1] = -1.38888888888741095749e-03; k.C[2] = 2.48015872894767294178e-05;
    k.C[3] = -2.75573143513906633035e-07; k.C[4] = 2.08757232129817482790e-09; k.C[5] = -1.13596475577881948265e-11;
    // C1 (Karney eq. 18)
    k.c1[0][0] = -1.0 / 32.0;      k.c1[0][1] = 6.0 / 32.0;       k.c1[0][2] = -16.0 / 32.0;
    k.c1[1][0] = -9.0 / 2048.0;    k.c1[1][1] = 64.0 / 2048.0;    k.c1[1][2] = -128.0 / 2048.0;
    k.c1[2][0] = 0.0;              k.c1[2][1] = 9.0 / 768.0;      k.c1[2][2] = -16.0 / 768.0;
    k.c1[3][0] = 0.0;              k.c1[3][1] = 3.0 / 512.0;      k.c1[3][2] = -5.0 / 512.0;
    k.c1[4][0] = 0.0;              k.c1[4][1] = 0.0;              k.c1[4][2] = -7.0 / 1280.0;
    k.c1[5][0] = 0.0;              k.c1[5][1] = 0.0;              k.c1[5][2] = -7.0 / 2048.0;
    // C1' (eq. 21)
    k.c1p[0][0] = 205.0 / 1536.0;    k.c1p[0][1] = -432.0 / 1536.0;   k.c1p[0][2] = 768.0 / 1536.0;
    k.c1p[1][0] = 4005.0 / 12288.0;  k.c1p[1][1] = -4736.0 / 12288.0; k.c1p[1][2] = 3840.0 / 12288.0;
    k.c1p[2][0] = 0.0;               k.c1p[2][1] = -225.0 / 384.0;    k.c1p[2][2] = 116.0 / 384.0;
    k.c1p[3][0] = 0.0;               k.c1p[3][1] = -7173.0 / 7680.0;  k.c1p[3][2] = 2695.0 / 7680.0;
    k.c1p[4][0] = 0.0;               k.c1p[4][1] = 0.0;               k.c1p[4][2] = 3467.0 / 7680.0;
    k.c1p[5][0] = 0.0;               k.c1p[5][1] = 0.0;               k.c1p[5][2] = 38081.0 / 61440.0;
    // A3 (eq. 24): coefficients of eps^k, polynomials in n
    k.a3[0] = 1.0;
    k.a3[1] = (n - 1.0) / 2.0;
    k.a3[2] = (n * (3.0 * n - 1.0) - 2.0) / 8.0;
    k.a3[3] = ((-n - 3.0) * n - 1.0) / 16.0;
    k.a3[4] = (-2.0 * n - 3.0) / 64.0;
    k.a3[5] = -3.0 / 128.0;
    // C3 (eq. 25): C3[l] = eps^l * sum_k c3[l-1][k] eps^k
    k.c3[0][0] = (1.0 - n) / 4.0; k.c3[0][1] = (1.0 - n * n) / 8.0; k.c3[0][2] = ((3.0 - n) * n + 3.0) / 64.0;
    k.c3[0][3] = (2.0 * n + 5.0) / 128.0; k.c3[0][4] = 3.0 / 128.0;
    k.c3[1][0] = ((n - 3.0) * n + 2.0) / 32.0; k.c3[1][1] = ((-3.0 * n - 2.0) * n + 3.0) / 64.0;
    k.c3[1][2] = (n + 3.0) / 128.0; k.c3[1][3] = 5.0 / 256.0;
    k.c3[2][0] = ((5.0 * n - 9.0) * n + 5.0) / 192.0; k.c3[2][1] = (9.0 - 10.0 * n) / 384.0; k.c3[2][2] = 7.0 / 512.0;
    k.c3[3][0] = (7.0 - 14.0 * n) / 512.0; k.c3[3][1] = 7.0 / 512.0;
    k.c3[4][0] = 21.0 / 2560.0;
    // atan(t) = t (1 + t^2 (at0 + t^2 (at1 + ...)))
    k.at[0] = -1.0 / 3.0; k.at[1] = 1.0 / 5.0; k.at[2] = -1.0 / 7.0; k.at[3] = 1.0 / 9.0; k.at[4] = -1.0 / 11.0;
    k.at[5] = 1.0 / 13.0;
    k.ts[0] = -1.0 / 6.0; k.ts[1] = 1.0 / 120.0; k.ts[2] = -1.0 / 5040.0;
    k.tc[0] = 1.0 / 24.0; k.tc[1] = -1.0 / 720.0; k.tc[2] = 1.0 / 40320.0;
    return k;
}

#if defined(__CUDACC__)
__constant__ GeodK d_geodk = make_geodk();
#endif
static const GeodK h_geodk = make_geodk();

#if defined(__CUDA_ARCH__)
#define OD_GK d_geodk
#else
#define OD_GK h_geodk
#endif

OD_HD void sincos_lib(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
    sincos(x, &s, &c);
#else
    s = sin(x);
    c = cos(x);
#endif
}

// sin and cos for |x| <= pi/4 (no range reduction): the two minimax kernels, < 1 ulp
OD_HD void sincos_q(double x, double& s, double& c) {
    const GeodK& K = OD_GK;
    const double z = x * x;
    const double rs = K.S[1] + z * (K.S[2] + z * (K.S[3] + z * (K.S[4] + z * K.S[5])));
    s = x + (z * x) * (K.S[0] + z * rs);
    const double rc = z * (K.C[0] + z * (K.C[1] + z * (K.C[2] + z * (K.C[3] + z * (K.C[4] + z * K.C[5])))));
    const double hz = 0.5 * z;
    if (fabs(x) < 0.3) {
        c = 1.0 - (hz - z * rc);
    } else {
        const double qx = fabs(x) > 0.78125 ? 0.28125 : 0.25 * fabs(x);
        c = (1.0 - qx) - ((hz - qx) - z * rc);
    }
}

// |x| <= 0.01 (64 km of arc): Taylor series to round-off (next terms: x^9/9! < 3e-24, x^10/10! < 3e-27)
OD_HD void sincos_tiny(double x, double& s, double& c) {
    const GeodK& K = OD_GK;
    const double z = x * x;
    s = x + x * (z * (K.ts[0] + z * (K.ts[1] + z * K.ts[2])));
    c = 1.0 + z * (-0.5 + z * (K.tc[0] + z * (K.tc[1] + z * K.tc[2])));
}

OD_HD void sincos_any(double x, double& s, double& c) {
    const double ax = fabs(x);
    if (ax <= 0.01) sincos_tiny(x, s, c);
    else if (ax <= kPio4) sincos_q(x, s, c);
    else sincos_lib(x, s, c);
}

// IEEE remainder(x, 360) with -180 -> 180
OD_HD double ang_normalize(double x) {
    if (fabs(x) < 180.0) return x;
    double y = x - 360.0 * rint(x * OD_GK.inv360);
    if (fabs(y) > 180.0) y = x - 360.0 * rint(x / 360.0);   // only within an ulp of a tie
    return y == -180.0 ? 180.0 : y;
}

OD_HD double ang_round(double x) {
    const double z = 0.0625;
    double y = fabs(x);
    y = y < z ? z - (z - y) : y;
    return copysign(y, x);
}

// sin and cos of an angle in degrees: exact reduction to [-45, 45] degrees, then the kernels
OD_HD void sincosd(double x, double& sx, double& cx) {
    const double q = rint(x * OD_GK.inv90);
    const double r = (x - 90.0 * q) * OD_GK.deg;
    const int iq = ((int)q) & 3;
    double s, c;
    sincos_q(r, s, c);
    sx = (iq == 0) ? s : (iq == 1) ? c : (iq == 2) ? -s : -c;
    cx = (iq == 0) ? c : (iq == 1) ? -s : (iq == 2) ? -c : s;
    if (x == 0.0) sx = x;
    cx += 0.0;
}

// atan2(y, x): for x > 0 and |y| <= 0.01 x (a step-sized longitude difference) the series of atan(y/x)
OD_HD double atan2_small(double y, double x) {
    if (x > 0.0 && fabs(y) <= 0.01 * x) {
        const GeodK& K = OD_GK;
        const double t = y / x, z = t * t;
        const double p = z * (K.at[0] + z * (K.at[1] + z * (K.at[2] + z * (K.at[3] + z * (K.at[4] + z * K.at[5])))));
        return t + t * p;
    }
    return atan2(y, x);
}

OD_HD double rsqrt_(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// sum_{l=1..6} c[l] sin(2 l x), Clenshaw
OD_HD double sin_series6(double sinx, double cosx, double c1, double c2, double c3, double c4,
                         double c5, double c6) {
    const double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y1 = c6;                       // n = 6 (even): y0 = 0
    double y0 = ar * y1 + c5;
    y1 = ar * y0 - y1 + c4;
    y0 = ar * y1 - y0 + c3;
    y1 = ar * y0 - y1 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// sum_{l=1..5} c[l] sin(2 l x)
OD_HD double sin_series5(double sinx, double cosx, double c1, double c2, double c3, double c4, double c5) {
    const double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y0 = c5, y1 = 0.0;             // n = 5 (odd): y0 = c5
    y1 = ar * y0 - y1 + c4;
    y0 = ar * y1 - y0 + c3;
    y1 = ar * y0 - y1 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// sum_{l=1..4} c[l] sin(2 l x)
OD_HD double sin_series4(double sinx, double cosx, double c1, double c2, double c3, double c4) {
    const double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y1 = c4;
    double y0 = ar * y1 + c3;
    y1 = ar * y0 - y1 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// sum_{l=1..3} c[l] sin(2 l x)
OD_HD double sin_series3(double sinx, double cosx, double c1, double c2, double c3) {
    const double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y0 = c3;
    double y1 = ar * y0 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// The part of the line initialisation that depends on the start latitude only.
struct GeodStart {
    double sbet1, cbet1;
    double lat1, sphi1, cphi1;     // start latitude (degrees) and its sine / cosine
};

OD_HD GeodStart geod_start(double lat1) {
    GeodStart p;
    if (fabs(lat1) > 90.0) lat1 = NAN;     // LatFix
    double sb, cb;
    sincosd(ang_round(lat1), sb, cb);
    p.lat1 = lat1;
    p.sphi1 = sb;
    p.cphi1 = cb;
    sb *= OD_GK.f1;
    const double r = rsqrt_(sb * sb + cb * cb);
    sb *= r;
    cb *= r;
    p.sbet1 = sb;
    p.cbet1 = cb > kTiny ? cb : kTiny;
    return p;
}

// Position at distance s12 (metres, may be negative) along azimuth azi1 (degrees) from (lon1, start).
OD_HD void geod_move(const GeodStart& p, double lon1, double azi1, double s12, double& lon2, double& lat2) {
    const GeodK& K = OD_GK;
    if (s12 == 0.0) {            // a particle that does not move stays where it is, to the bit (a zero velocity is the fallback of an
        lon2 = ang_normalize(ang_normalize(lon1));      // uncovered Runge-Kutta stage: the next stage must sample the start point itself,
        lat2 = p.lat1;                                  // e.g. on the boundary row of the block, not a point an ulp outside it)
        return;
    }
    double salp1, calp1;
    sincosd(ang_round(ang_normalize(azi1)), salp1, calp1);
    const double sbet1 = p.sbet1, cbet1 = p.cbet1;

    const double salp0 = salp1 * cbet1;
    const double t0 = salp1 * sbet1;
    const double calp0sq = calp1 * calp1 + t0 * t0;
    // sig1: (ssig1, csig1) = (sbet1, cbet1 calp1) / hypot(..) and sbet1^2 + cbet1^2 calp1^2 = calp0^2
    double ssig1, csig1, calp0;
    const double somg1 = salp0 * sbet1;
    const double comg1 = (sbet1 != 0.0 || calp1 != 0.0) ? cbet1 * calp1 : 1.0;
    if (calp0sq > 1e-200) {
        const double inv0 = rsqrt_(calp0sq);
        calp0 = calp0sq * inv0;
        ssig1 = sbet1 * inv0;
        csig1 = comg1 * inv0;
    } else {                                  // on the equator heading due east / west
        calp0 = sqrt(calp0sq);
        ssig1 = 0.0;
        csig1 = 1.0;
    }
    const double k2 = calp0sq * K.ep2;
    // eps = (sqrt(1+k2)-1)/(sqrt(1+k2)+1), k2 <= 0.0068
    const double eps = k2 * (0.25 + k2 * (-0.125 + k2 * (0.078125 + k2 * (-0.0546875 + k2 * (0.041015625 +
                       k2 * (-0.0322265625 + k2 * 0.02618408203125))))));
    const double eps2 = eps * eps;

    // A1 (eq. 17): 1 + A1m1 = (1 + tA) / (1 - eps)
    const double tA = eps2 * (eps2 * (eps2 + 4.0) + 64.0) * 0.00390625;
    // C1 (eq. 18), C1' (eq. 21), C3 (eq. 25).  eps <= 1.68e-3 on WGS84, so the dropped terms
    // C1[5..6] <= 7e-17, C1'[6] <= 1.4e-17, f C3[4..5] <= 4e-16 radians are below double round-off of sigma.
    double d = eps;
    const double C1_1 = d * ((K.c1[0][0] * eps2 + K.c1[0][1]) * eps2 + K.c1[0][2]);
    const double C1p_1 = d * ((K.c1p[0][0] * eps2 + K.c1p[0][1]) * eps2 + K.c1p[0][2]);
    const double C3_1 = d * ((((K.c3[0][4] * eps + K.c3[0][3]) * eps + K.c3[0][2]) * eps + K.c3[0][1]) * eps + K.c3[0][0]);
    d *= eps;
    const double C1_2 = d * ((K.c1[1][0] * eps2 + K.c1[1][1]) * eps2 + K.c1[1][2]);
    const double C1p_2 = d * ((K.c1p[1][0] * eps2 + K.c1p[1][1]) * eps2 + K.c1p[1][2]);
    const double C3_2 = d * (((K.c3[1][3] * eps + K.c3[1][2]) * eps + K.c3[1][1]) * eps + K.c3[1][0]);
    d *= eps;
    const double C1_3 = d * (K.c1[2][1] * eps2 + K.c1[2][2]);
    const double C1p_3 = d * (K.c1p[2][1] * eps2 + K.c1p[2][2]);
    const double C3_3 = d * ((K.c3[2][2] * eps + K.c3[2][1]) * eps + K.c3[2][0]);
    d *= eps;
    const double C1_4 = d * (K.c1[3][1] * eps2 + K.c1[3][2]);
    const double C1p_4 = d * (K.c1p[3][1] * eps2 + K.c1p[3][2]);
    d *= eps;
    const double C1p_5 = d * K.c1p[4][2];
    const double B11 = sin_series4(ssig1, csig1, C1_1, C1_2, C1_3, C1_4);
    // sin, cos of B11 (|B11| < 2e-3): Taylor to round-off
    const double zB = B11 * B11;
    const double sB = B11 + B11 * (zB * (K.ts[0] + zB * K.ts[1]));
    const double cB = 1.0 + zB * (-0.5 + zB * (K.tc[0] + zB * K.tc[1]));
    const double stau1 = ssig1 * cB + csig1 * sB;
    const double ctau1 = csig1 * cB - ssig1 * sB;
    // A3 (eq. 24)
    const double A3 = ((((K.a3[5] * eps + K.a3[4]) * eps + K.a3[3]) * eps + K.a3[2]) * eps + K.a3[1]) * eps + 1.0;
    const double A3c = K.neg_f * salp0 * A3;
    const double B31 = sin_series3(ssig1, csig1, C3_1, C3_2, C3_3);

    // position on the line: tau12 = s12 / (b (1 + A1m1)) = s12 (1 - eps) / (b (1 + tA))
    // 1 / (1 + tA) = 1 - tA + tA^2 - ... with tA <= 7.1e-7 (tA^3 < 4e-19)
    const double tau12 = s12 * K.inv_b * (1.0 - eps) * (1.0 - tA + tA * tA);
    double st, ct;
    sincos_any(tau12, st, ct);
    const double B12 = -sin_series5(stau1 * ct + ctau1 * st, ctau1 * ct - stau1 * st,
                                    C1p_1, C1p_2, C1p_3, C1p_4, C1p_5);
    const double sig12 = tau12 - (B12 - B11);
    double ssig12, csig12;
    sincos_any(sig12, ssig12, csig12);
    const double ssig2 = ssig1 * csig12 + csig1 * ssig12;
    double csig2 = csig1 * csig12 - ssig1 * ssig12;
    const double sbet2 = calp0 * ssig2;
    const double t2 = calp0 * csig2;
    double cbet2 = sqrt(salp0 * salp0 + t2 * t2);
    if (cbet2 == 0.0) cbet2 = csig2 = kTiny;
    const double somg2 = salp0 * ssig2, comg2 = csig2;
    const double omg12 = atan2_small(somg2 * comg1 - comg2 * somg1, comg2 * comg1 + somg2 * somg1);
    const double lam12 = omg12 + A3c * (sig12 + (sin_series3(ssig2, csig2, C3_1, C3_2, C3_3) - B31));
    const double lon12 = lam12 * K.rad2deg;
    lon2 = ang_normalize(ang_normalize(lon1) + ang_normalize(lon12));
    // lat2 = atan2(sbet2, f1 cbet2); evaluated as lat1 + atan2(sin(phi2 - phi1), cos(phi2 - phi1)) so that a
    // step-sized move needs the arctangent of a small ratio only (and a zero-length move returns lat1 itself)
    const double c2 = K.f1 * cbet2;
    lat2 = p.lat1 + atan2_small(sbet2 * p.cphi1 - c2 * p.sphi1, c2 * p.cphi1 + sbet2 * p.sphi1) * K.rad2deg;
}

OD_HD void geod_direct(double lon1, double lat1, double azi1, double s12, double& lon2, double& lat2) {
    GeodStart p = geod_start(lat1);
    geod_move(p, lon1, azi1, s12, lon2, lat2);
}

// ---------------------------------------------------------------------------------------------------------
// Short-arc direct problem: Taylor series of the geodesic in the northward / eastward displacement.
//
// Along a geodesic  dphi/ds = cos(alpha) V^3/c,  dlambda/ds = sin(alpha) V/(c cos phi),  dalpha/ds = sin(alpha) t V/c
// (c = a^2/b, V^2 = 1 + e'^2 cos^2 phi, t = tan phi).  Differentiating along the line gives phi2 - phi1 and
// lambda2 - lambda1 as polynomials in X = s cos(alpha)/N, Y = s sin(alpha)/N whose coefficients are polynomials in
// V^2 and t at the start point (the classical Legendre series; tools/gen_geod_series.py derives them with SymPy to
// fifth order and writes od_geod_series.inc).  No azimuth, no trigonometric function of the move: the displacement
// components dt*v, dt*u enter directly.  With r = (|X| + |Y|) max(1, |t|) the measured difference from the full
// solution is at double round-off (<= 5e-14 deg) for r <= 4e-3, 1.5e-13 for r <= 5e-3, 2.4e-12 for r <= 8e-3
// (sixth-order truncation).  kSeriesMaxR = 4e-3 admits every move of <= 10 km at 60N, <= 3 km at 80N; the caller
// hands longer or more polar moves (and NaN) to geod_move (tests/test_hostmath.py::test_series_*,
// tools/gen_geod_series.py --check).
// ---------------------------------------------------------------------------------------------------------
#ifndef OD_SERIES_MAX_R
#define OD_SERIES_MAX_R 4.0e-3
#endif
constexpr double kSeriesMaxR = OD_SERIES_MAX_R;

struct SeriesStart {
    double lat1;        // degrees
    double t, W;        // tan(phi1), V^2 = 1 + e'^2 cos^2(phi1)
    double vc;          // V / c = 1 / N(phi1)                      [1/m]
    double icd;         // (180/pi) / cos(phi1)
};

OD_HD SeriesStart series_start(double lat1) {
    SeriesStart p;
    if (fabs(lat1) > 90.0) lat1 = NAN;
    double sp, cp;
    sincosd(lat1, sp, cp);
    const double ic = 1.0 / cp;
    p.lat1 = lat1;
    p.t = sp * ic;
    p.W = 1.0 + OD_GK.ep2 * (cp * cp);
    p.vc = sqrt(p.W) * OD_GK.inv_c;
    p.icd = ic * OD_GK.rad2deg;
    return p;
}

// Position after moving xn metres north and ye metres east along the geodesic that starts with that direction.
// Returns false (lon2/lat2 untouched) when the move is outside the series' range (long steps, near the poles, NaN).
OD_HD bool series_move(const SeriesStart& p, double lon1, double xn, double ye, double& lon2, double& lat2) {
    const double X = xn * p.vc, Y = ye * p.vc;
    const double r = (fabs(X) + fabs(Y)) * fmax(1.0, fabs(p.t));
    if (!(r <= kSeriesMaxR)) return false;
    const double t = p.t, T = t * t, W = p.W;
#include "od_geod_series.inc"
    const double X2 = X * X, Y2 = Y * Y, XY2 = X2 * Y2, X4 = X2 * X2, Y4 = Y2 * Y2;
    const double Pe = (p12 * Y2 + p30 * X2) + (p14 * Y4 + p32 * XY2 + p50 * X4);          // even in t, times X
    const double Po = (p02 * Y2 + p20 * X2) + (p04 * Y4 + p22 * XY2 + p40 * X4);          // odd in t
    const double P = X + (X * Pe + t * Po);
    const double Qe = (q03 * Y2 + q21 * X2) + (q05 * Y4 + q23 * XY2 + q41 * X4);
    const double Qo = X * (1.0 + (q13 * Y2 + q31 * X2));
    const double Q = Y + Y * (Qe + t * Qo);
    (void)p10; (void)q01; (void)q11;
    lat2 = p.lat1 + (W * P) * OD_GK.rad2deg;
    lon2 = ang_normalize(ang_normalize(lon1) + Q * p.icd);
    return true;
}

// Third-order truncation of the same series, for positions that only feed the field sampler (Runge-Kutta mid-points).
// Neglected: the fourth- and fifth-order terms, bounded by ~3 r^4 radians with r as above: 2e-9 m for a 300 m half step at
// 60N (r = 8e-5), 2e-5 m at r = kSeries3MaxR = 1e-3 -- the size of the float32 azimuth / distance rounding of the
// reference's mid-points that SeriesMath skips anyway (od_advect.cuh).  Five coefficients instead of eighteen: what the
// Runge-Kutta loop keeps live fits in registers.  Returns false beyond kSeries3MaxR (the caller then takes series_move).
#ifndef OD_SERIES3_MAX_R
#define OD_SERIES3_MAX_R 1.0e-3
#endif
constexpr double kSeries3MaxR = OD_SERIES3_MAX_R;

// The third-order move without the longitude normalisation: lon1n = ang_normalize(lon1) is the caller's (the Runge-Kutta loop
// of the specialised step kernel normalises the start longitude once, od_spec.cuh); lonraw = lon1n + dlon is not wrapped.
// Always evaluates; returns whether the move was inside the series' range (lonraw / lat2 are meaningless otherwise).
OD_HD bool series_move3_raw(const SeriesStart& p, double lon1n, double xn, double ye, double& lonraw, double& lat2) {
    const double X = xn * p.vc, Y = ye * p.vc;
    const double r = (fabs(X) + fabs(Y)) * fmax(1.0, fabs(p.t));
    const double t = p.t, T = t * t, W = p.W;
    const double p20 = 1.5 - 1.5 * W;                                               // times t  (p02 = -1/2)
    const double p12 = -2.0 * T + W * (1.5 * T - 1.0 / 6.0);
    const double p30 = 2.0 * T + W * (-4.5 * T + W * (2.5 * T - 0.5) + 0.5);
    const double q03 = (-1.0 / 3.0) * T;
    const double q21 = T + (1.0 / 3.0) * W;
    const double X2 = X * X, Y2 = Y * Y;
    const double P = X + (X * (p12 * Y2 + p30 * X2) + t * (p20 * X2 - 0.5 * Y2));
    const double Q = Y + Y * ((q03 * Y2 + q21 * X2) + t * X);
    lat2 = p.lat1 + (W * P) * OD_GK.rad2deg;
    lonraw = lon1n + Q * p.icd;
    return r <= kSeries3MaxR;
}

OD_HD bool series_move3(const SeriesStart& p, double lon1, double xn, double ye, double& lon2, double& lat2) {
    const double X = xn * p.vc, Y = ye * p.vc;
    const double r = (fabs(X) + fabs(Y)) * fmax(1.0, fabs(p.t));
    if (!(r <= kSeries3MaxR)) return false;
    double lonraw, la;
    series_move3_raw(p, ang_normalize(lon1), xn, ye, lonraw, la);
    lat2 = la;
    lon2 = ang_normalize(lonraw);
    return true;
}

// the rarely taken full solution, kept out of line so that the callers stay small
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static
#endif
void geod_direct_ne(double lon1, double lat1, double xn, double ye, double* lon2, double* lat2) {
    const double az = atan2(ye, xn) * OD_GK.rad2deg;
    geod_direct(lon1, lat1, az, sqrt(xn * xn + ye * ye), *lon2, *lat2);
}

// (north, east) displacement in metres -> position, series first, full solution otherwise
OD_HD void geod_move_ne(const SeriesStart& p, double lon1, double xn, double ye, double& lon2, double& lat2) {
    if (series_move(p, lon1, xn, ye, lon2, lat2)) return;
    geod_direct_ne(lon1, p.lat1, xn, ye, &lon2, &lat2);
}

}  // namespace od
