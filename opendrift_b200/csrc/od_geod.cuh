// od_geod.cuh -- WGS84 direct geodesic (Karney 2013, order-6 series), float64, one call per particle.
//
// Replaces pyproj.Geod(ellps='WGS84').fwd, which the reference calls for every position update
// (opendrift/models/basemodel/__init__.py:4643-4657) and every Runge-Kutta mid-point
// (opendrift/models/physics_methods.py:632-635, 649-652, 663-666).  The algorithm is the published one
// (J. Geodesy 87:43-55, eqs. 7-21), written here for a GPU thread: the part that depends only on the start
// latitude is split off (all four moves of an RK4 step start from the same point), the series
// coefficients are Horner polynomials in registers, and no back azimuth / reduced length is computed.
//
// The same source compiles for the host (tests/hostshim) so that the arithmetic can be checked against
// the oracle without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OD_HD __host__ __device__ __forceinline__
#else
#define OD_HD static inline
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
    // A3 = sum_k A3x[k] eps^k  (Karney eq. 24), polynomials in n
    static constexpr double A3_0 = 1.0;
    static constexpr double A3_1 = (n - 1.0) * (1.0 / 2.0);
    static constexpr double A3_2 = (n * (3.0 * n - 1.0) - 2.0) * (1.0 / 8.0);
    static constexpr double A3_3 = ((-n - 3.0) * n - 1.0) * (1.0 / 16.0);
    static constexpr double A3_4 = (-2.0 * n - 3.0) * (1.0 / 64.0);
    static constexpr double A3_5 = -3.0 * (1.0 / 128.0);
    // C3[l] = sum_{k>=l} C3x[l][k] eps^k  (Karney eq. 25)
    static constexpr double C3_11 = (1.0 - n) * (1.0 / 4.0);
    static constexpr double C3_12 = (1.0 - n * n) * (1.0 / 8.0);
    static constexpr double C3_13 = ((3.0 - n) * n + 3.0) * (1.0 / 64.0);
    static constexpr double C3_14 = (2.0 * n + 5.0) * (1.0 / 128.0);
    static constexpr double C3_15 = 3.0 * (1.0 / 128.0);
    static constexpr double C3_22 = ((n - 3.0) * n + 2.0) * (1.0 / 32.0);
    static constexpr double C3_23 = ((-3.0 * n - 2.0) * n + 3.0) * (1.0 / 64.0);
    static constexpr double C3_24 = (n + 3.0) * (1.0 / 128.0);
    static constexpr double C3_25 = 5.0 * (1.0 / 256.0);
    static constexpr double C3_33 = ((5.0 * n - 9.0) * n + 5.0) * (1.0 / 192.0);
    static constexpr double C3_34 = (9.0 - 10.0 * n) * (1.0 / 384.0);
    static constexpr double C3_35 = 7.0 * (1.0 / 512.0);
    static constexpr double C3_44 = (7.0 - 14.0 * n) * (1.0 / 512.0);
    static constexpr double C3_45 = 7.0 * (1.0 / 512.0);
    static constexpr double C3_55 = 21.0 * (1.0 / 2560.0);
};

constexpr double kDeg = 0.017453292519943295769;       // pi / 180
constexpr double kRad2Deg = 57.295779513082320877;     // 180 / pi
constexpr double kTiny = 1.4916681462400413e-154;      // sqrt(DBL_MIN)

OD_HD void sincos_(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
    sincos(x, &s, &c);
#else
    s = sin(x);
    c = cos(x);
#endif
}

// IEEE remainder(x, 360) with -180 -> 180
OD_HD double ang_normalize(double x) {
    double y = x - 360.0 * rint(x / 360.0);
    return y == -180.0 ? 180.0 : y;
}

OD_HD double ang_round(double x) {
    const double z = 1.0 * (1.0 / 16.0);
    double y = fabs(x);
    y = y < z ? z - (z - y) : y;
    return copysign(y, x);
}

// sin and cos of an angle in degrees, exact quadrant reduction
OD_HD void sincosd(double x, double& sx, double& cx) {
    double q = rint(x / 90.0);
    double r = (x - 90.0 * q) * kDeg;
    int iq = ((int)q) & 3;
    double s, c;
    sincos_(r, s, c);
    sx = (iq == 0) ? s : (iq == 1) ? c : (iq == 2) ? -s : -c;
    cx = (iq == 0) ? c : (iq == 1) ? -s : (iq == 2) ? -c : s;
    if (x == 0.0) sx = x;
    cx += 0.0;
}

// sum_{l=1..6} c[l] sin(2 l x), Clenshaw
OD_HD double sin_series6(double sinx, double cosx, double c1, double c2, double c3, double c4,
                         double c5, double c6) {
    double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y1 = ar * 0.0 - 0.0 + c6;      // n = 6 (even): y0 = 0
    double y0 = ar * y1 - 0.0 + c5;
    y1 = ar * y0 - y1 + c4;
    y0 = ar * y1 - y0 + c3;
    y1 = ar * y0 - y1 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// sum_{l=1..5} c[l] sin(2 l x)
OD_HD double sin_series5(double sinx, double cosx, double c1, double c2, double c3, double c4, double c5) {
    double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
    double y0 = c5, y1 = 0.0;             // n = 5 (odd): y0 = c5
    y1 = ar * y0 - y1 + c4;
    y0 = ar * y1 - y0 + c3;
    y1 = ar * y0 - y1 + c2;
    y0 = ar * y1 - y0 + c1;
    return 2.0 * sinx * cosx * y0;
}

// The part of the line initialisation that depends on the start latitude only.
struct GeodStart {
    double sbet1, cbet1;
};

OD_HD GeodStart geod_start(double lat1) {
    GeodStart p;
    if (fabs(lat1) > 90.0) lat1 = NAN;     // LatFix
    double sb, cb;
    sincosd(ang_round(lat1), sb, cb);
    sb *= Wgs84::f1;
    double r = sqrt(sb * sb + cb * cb);
    sb /= r;
    cb /= r;
    p.sbet1 = sb;
    p.cbet1 = cb > kTiny ? cb : kTiny;
    return p;
}

// Position at distance s12 (metres, may be negative) along azimuth azi1 (degrees) from (lon1, start).
OD_HD void geod_move(const GeodStart& p, double lon1, double azi1, double s12, double& lon2, double& lat2) {
    typedef Wgs84 E;
    double salp1, calp1;
    sincosd(ang_round(ang_normalize(azi1)), salp1, calp1);
    const double sbet1 = p.sbet1, cbet1 = p.cbet1;

    const double salp0 = salp1 * cbet1;
    const double t0 = salp1 * sbet1;
    const double calp0 = sqrt(calp1 * calp1 + t0 * t0);
    double ssig1 = sbet1;
    const double somg1 = salp0 * sbet1;
    double csig1 = (sbet1 != 0.0 || calp1 != 0.0) ? cbet1 * calp1 : 1.0;
    const double comg1 = csig1;
    {
        double r = 1.0 / sqrt(ssig1 * ssig1 + csig1 * csig1);
        ssig1 *= r;
        csig1 *= r;
    }
    const double k2 = calp0 * calp0 * E::ep2;
    const double eps = k2 / (2.0 * (1.0 + sqrt(1.0 + k2)) + k2);
    const double eps2 = eps * eps;

    // A1 - 1  (eq. 17)
    const double tA = eps2 * (eps2 * (eps2 + 4.0) + 64.0) * (1.0 / 256.0);
    const double A1m1 = (tA + eps) / (1.0 - eps);
    // C1 (eq. 18)
    double d = eps;
    const double C1_1 = d * ((6.0 - eps2) * eps2 - 16.0) * (1.0 / 32.0);
    d *= eps;
    const double C1_2 = d * ((64.0 - 9.0 * eps2) * eps2 - 128.0) * (1.0 / 2048.0);
    d *= eps;
    const double C1_3 = d * (9.0 * eps2 - 16.0) * (1.0 / 768.0);
    d *= eps;
    const double C1_4 = d * (3.0 * eps2 - 5.0) * (1.0 / 512.0);
    d *= eps;
    const double C1_5 = -7.0 * d * (1.0 / 1280.0);
    d *= eps;
    const double C1_6 = -7.0 * d * (1.0 / 2048.0);
    const double B11 = sin_series6(ssig1, csig1, C1_1, C1_2, C1_3, C1_4, C1_5, C1_6);
    double sB, cB;
    sincos_(B11, sB, cB);
    const double stau1 = ssig1 * cB + csig1 * sB;
    const double ctau1 = csig1 * cB - ssig1 * sB;
    // C1' (eq. 21)
    d = eps;
    const double C1p_1 = d * (eps2 * (205.0 * eps2 - 432.0) + 768.0) * (1.0 / 1536.0);
    d *= eps;
    const double C1p_2 = d * (eps2 * (4005.0 * eps2 - 4736.0) + 3840.0) * (1.0 / 12288.0);
    d *= eps;
    const double C1p_3 = d * (116.0 - 225.0 * eps2) * (1.0 / 384.0);
    d *= eps;
    const double C1p_4 = d * (2695.0 - 7173.0 * eps2) * (1.0 / 7680.0);
    d *= eps;
    const double C1p_5 = 3467.0 * d * (1.0 / 7680.0);
    d *= eps;
    const double C1p_6 = 38081.0 * d * (1.0 / 61440.0);
    // A3, C3 (eqs. 24, 25)
    const double A3 = ((((E::A3_5 * eps + E::A3_4) * eps + E::A3_3) * eps + E::A3_2) * eps + E::A3_1) * eps + E::A3_0;
    d = eps;
    const double C3_1 = d * ((((E::C3_15 * eps + E::C3_14) * eps + E::C3_13) * eps + E::C3_12) * eps + E::C3_11);
    d *= eps;
    const double C3_2 = d * (((E::C3_25 * eps + E::C3_24) * eps + E::C3_23) * eps + E::C3_22);
    d *= eps;
    const double C3_3 = d * ((E::C3_35 * eps + E::C3_34) * eps + E::C3_33);
    d *= eps;
    const double C3_4 = d * (E::C3_45 * eps + E::C3_44);
    d *= eps;
    const double C3_5 = d * E::C3_55;
    const double A3c = -E::f * salp0 * A3;
    const double B31 = sin_series5(ssig1, csig1, C3_1, C3_2, C3_3, C3_4, C3_5);

    // position on the line
    const double tau12 = s12 / (E::b * (1.0 + A1m1));
    double st, ct;
    sincos_(tau12, st, ct);
    const double B12 = -sin_series6(stau1 * ct + ctau1 * st, ctau1 * ct - stau1 * st,
                                    C1p_1, C1p_2, C1p_3, C1p_4, C1p_5, C1p_6);
    const double sig12 = tau12 - (B12 - B11);
    double ssig12, csig12;
    sincos_(sig12, ssig12, csig12);
    const double ssig2 = ssig1 * csig12 + csig1 * ssig12;
    double csig2 = csig1 * csig12 - ssig1 * ssig12;
    const double sbet2 = calp0 * ssig2;
    const double t2 = calp0 * csig2;
    double cbet2 = sqrt(salp0 * salp0 + t2 * t2);
    if (cbet2 == 0.0) cbet2 = csig2 = kTiny;
    const double somg2 = salp0 * ssig2, comg2 = csig2;
    const double omg12 = atan2(somg2 * comg1 - comg2 * somg1, comg2 * comg1 + somg2 * somg1);
    const double lam12 = omg12 + A3c * (sig12 + (sin_series5(ssig2, csig2, C3_1, C3_2, C3_3, C3_4, C3_5) - B31));
    const double lon12 = lam12 * kRad2Deg;
    lon2 = ang_normalize(ang_normalize(lon1) + ang_normalize(lon12));
    lat2 = atan2(sbet2, E::f1 * cbet2) * kRad2Deg;
}

OD_HD void geod_direct(double lon1, double lat1, double azi1, double s12, double& lon2, double& lat2) {
    GeodStart p = geod_start(lat1);
    geod_move(p, lon1, azi1, s12, lon2, lat2);
}

}  // namespace od
