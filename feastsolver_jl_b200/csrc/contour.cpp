// Contour constructors of the C ABI (host, scalar).  Restates src/contour.jl:26-86 of the
// reference: same node order, same weights, same divisibility / corner errors.
#include <cmath>
#include <complex>
#include <vector>

#include "../../include/feast_cuda.h"

typedef std::complex<double> hc;
static const double kPi = 3.14159265358979323846264338327950288;

static inline void put(feast_c128* a, int i, hc v) { a[i].re = v.real(); a[i].im = v.imag(); }

// Gauss-Legendre nodes/weights on [-1,1], ascending (as FastGaussQuadrature.gausslegendre
// returns them; call sites src/contour.jl:37,52).  Newton on P_n with the three-term recurrence.
extern "C" int feast_gauss_legendre(int n, double* x, double* w) {
    if (n < 1) return -1;
    if (!x) return -2;
    if (!w) return -3;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double t = cosl((long double)kPi * (i + 0.75L) / (n + 0.5L));
        long double pp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1.0L, p1 = t;
            for (int k = 2; k <= n; ++k) {
                long double p2 = ((2.0L * k - 1.0L) * t * p1 - (k - 1.0L) * p0) / k;
                p0 = p1; p1 = p2;
            }
            if (n == 1) { p0 = 1.0L; p1 = t; }
            pp = n * (t * p1 - p0) / (t * t - 1.0L);
            long double dt = p1 / pp;
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        // recompute derivative at the converged root
        long double p0 = 1.0L, p1 = t;
        for (int k = 2; k <= n; ++k) {
            long double p2 = ((2.0L * k - 1.0L) * t * p1 - (k - 1.0L) * p0) / k;
            p0 = p1; p1 = p2;
        }
        pp = n * (t * p1 - p0) / (t * t - 1.0L);
        long double wi = 2.0L / ((1.0L - t * t) * pp * pp);
        x[i] = (double)(-t); x[n - 1 - i] = (double)t;
        w[i] = (double)wi;   w[n - 1 - i] = (double)wi;
    }
    if (n % 2 == 1) x[n / 2] = 0.0;
    return 0;
}

// src/contour.jl:26-31
extern "C" int feast_contour_circular_trapezoidal(feast_c128 c, double r, int N, feast_c128* z, feast_c128* w) {
    if (N < 1) return -3;
    if (!z) return -4;
    if (!w) return -5;
    const hc cc(c.re, c.im);
    const double a = kPi / N, b = 2 * kPi - kPi / N;
    for (int i = 0; i < N; ++i) {
        // LinRange(a, b, N)[i+1]
        const double th = (N == 1) ? a : a + (b - a) * ((double)i / (double)(N - 1));
        const hc e = std::exp(hc(0.0, th));
        put(z, i, r * e + cc);
        put(w, i, r * e / (double)N);
    }
    return 0;
}

// src/contour.jl:33-44
extern "C" int feast_contour_circular_gauss(feast_c128 c, double r, int N, feast_c128* z, feast_c128* w) {
    if (N < 2 || N % 2 != 0) return -3;  // "Number of nodes must be multiple of 2"
    if (!z) return -4;
    if (!w) return -5;
    const int n = N / 2;
    std::vector<double> gx(n), gw(n);
    feast_gauss_legendre(n, gx.data(), gw.data());
    const hc cc(c.re, c.im);
    for (int i = 0; i < n; ++i) {
        const double phi = (kPi / 2.0) * (gx[i] + 1.0);
        const hc e1 = std::exp(hc(0.0, phi)), e2 = std::exp(hc(0.0, phi + kPi));
        put(z, i, r * e1 + cc);
        put(z, n + i, r * e2 + cc);
        put(w, i, r * e1 * gw[i] / 4.0);
        put(w, n + i, r * e2 * gw[i] / 4.0);
    }
    return 0;
}

static bool corners_ok(hc bl, hc tr) { return bl.real() < tr.real() && bl.imag() < tr.imag(); }

// src/contour.jl:47-63 (clockwise: top, right, bottom, left)
extern "C" int feast_contour_rectangular_gauss(feast_c128 bottom_left, feast_c128 top_right, int N, feast_c128* z,
                                               feast_c128* w) {
    const hc bl(bottom_left.re, bottom_left.im), tr(top_right.re, top_right.im);
    if (!corners_ok(bl, tr)) return -1;  // "Invalid corners" (contour.jl:15)
    if (N < 4 || N % 4 != 0) return -3;  // "Number of nodes must be multiple of 4"
    if (!z) return -4;
    if (!w) return -5;
    const int n = N / 4;
    std::vector<double> gx(n), gw(n);
    feast_gauss_legendre(n, gx.data(), gw.data());
    const double top_len = tr.real() - bl.real(), side_len = tr.imag() - bl.imag();
    const hc I(0.0, 1.0);
    const hc scale = 1.0 / (-4.0 * kPi * I);
    for (int i = 0; i < n; ++i) {
        const double xr = gx[n - 1 - i];  // reverse(gq_nodes)
        put(z, i, (gx[i] + 1.0) * (top_len / 2.0) + hc(bl.real(), tr.imag()));
        put(z, n + i, (gx[i] + 1.0) * (I * side_len / 2.0) + hc(tr.real(), bl.imag()));
        put(z, 2 * n + i, (xr + 1.0) * (top_len / 2.0) + hc(bl.real(), bl.imag()));
        put(z, 3 * n + i, (xr + 1.0) * (I * side_len / 2.0) + hc(bl.real(), bl.imag()));
        put(w, i, hc(gw[i] * top_len, 0.0) * scale);
        put(w, n + i, (-I * gw[i] * side_len) * scale);
        put(w, 2 * n + i, hc(-gw[i] * top_len, 0.0) * scale);
        put(w, 3 * n + i, (I * gw[i] * side_len) * scale);
    }
    return 0;
}

static inline double linrange(double a, double b, int len, int i) {  // LinRange(a,b,len)[i+1]
    return len == 1 ? a : a + (b - a) * ((double)i / (double)(len - 1));
}

// src/contour.jl:66-86
extern "C" int feast_contour_rectangular_trapezoidal(feast_c128 bottom_left, feast_c128 top_right, int N,
                                                     feast_c128* z, feast_c128* w) {
    const hc bl(bottom_left.re, bottom_left.im), tr(top_right.re, top_right.im);
    if (!corners_ok(bl, tr)) return -1;
    if (N < 4 || N % 4 != 0) return -3;
    if (!z) return -4;
    if (!w) return -5;
    const int n = N / 4;
    const hc I(0.0, 1.0);
    const double top_len = tr.real() - bl.real(), side_len = tr.imag() - bl.imag();
    const hc scale = 1.0 / (-2.0 * kPi * I);
    for (int i = 0; i < n; ++i) {
        put(z, i, hc(linrange(bl.real(), tr.real(), n + 1, i), tr.imag()));
        put(z, n + i, hc(tr.real(), linrange(tr.imag(), bl.imag(), n + 1, i)));
        put(z, 2 * n + i, hc(linrange(tr.real(), bl.real(), n + 1, i), bl.imag()));
        put(z, 3 * n + i, hc(bl.real(), linrange(bl.imag(), tr.imag(), n + 1, i)));
    }
    std::vector<hc> ww(N);
    const double dn = (double)n;
    ww[0] = I * side_len / (2 * dn) + top_len / (2 * dn);
    for (int i = 1; i < n; ++i) ww[i] = top_len / dn;
    ww[n] = top_len / (2 * dn) - I * side_len / (2 * dn);
    for (int i = n + 1; i < 2 * n; ++i) ww[i] = -I * side_len / dn;
    ww[2 * n] = -I * side_len / (2 * dn) - top_len / (2 * dn);
    for (int i = 2 * n + 1; i < 3 * n; ++i) ww[i] = -top_len / dn;
    ww[3 * n] = -top_len / (2 * dn) + I * side_len / (2 * dn);
    for (int i = 3 * n + 1; i < 4 * n; ++i) ww[i] = I * side_len / dn;
    for (int i = 0; i < N; ++i) put(w, i, ww[i] * scale);
    return 0;
}
