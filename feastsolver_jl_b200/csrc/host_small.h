// Host-side helpers for m0 x m0 complex matrices (column-major).  m0 <= a few hundred, so
// plain loops are fine; the library deliberately links no LAPACK (none is installed
// system-wide, and the reduced eigenproblem belongs to the caller).
#pragma once
#include <complex>
#include <vector>
#include <algorithm>
#include <cmath>

typedef std::complex<double> hc128;

// Cholesky G = R^H R (R upper) of a Hermitian matrix (column-major).  Returns false when a pivot is not
// safely positive (<= min_pivot); the pivot is then replaced by min_pivot so that R stays finite, but the
// caller should not use such a factor (see orthonormalize in api.cu: it retries with a shifted matrix).
inline bool chol_upper(int m, const std::vector<hc128>& G, std::vector<hc128>& R, double min_pivot) {
    R.assign((size_t)m * m, hc128(0, 0));
    bool ok = true;
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < j; ++i) {
            hc128 s = G[(size_t)j * m + i];  // G(i,j)
            for (int k = 0; k < i; ++k) s -= std::conj(R[(size_t)i * m + k]) * R[(size_t)j * m + k];
            R[(size_t)j * m + i] = s / R[(size_t)i * m + i].real();
        }
        double d = G[(size_t)j * m + j].real();
        for (int k = 0; k < j; ++k) d -= std::norm(R[(size_t)j * m + k]);
        if (!(d > min_pivot)) { d = min_pivot; ok = false; }
        R[(size_t)j * m + j] = hc128(std::sqrt(d), 0.0);
    }
    return ok;
}

// inverse of an upper-triangular matrix (column-major)
inline void triu_inverse(int m, const std::vector<hc128>& R, std::vector<hc128>& Ri) {
    Ri.assign((size_t)m * m, hc128(0, 0));
    for (int j = 0; j < m; ++j) {
        Ri[(size_t)j * m + j] = 1.0 / R[(size_t)j * m + j];
        for (int i = j - 1; i >= 0; --i) {
            hc128 s(0, 0);
            for (int k = i + 1; k <= j; ++k) s += R[(size_t)k * m + i] * Ri[(size_t)j * m + k];
            Ri[(size_t)j * m + i] = -s / R[(size_t)i * m + i];
        }
    }
}

// C = A * B (all m x m column-major)
inline void matmul_small(int m, const std::vector<hc128>& A, const std::vector<hc128>& B, std::vector<hc128>& C) {
    C.assign((size_t)m * m, hc128(0, 0));
    for (int j = 0; j < m; ++j)
        for (int k = 0; k < m; ++k) {
            const hc128 b = B[(size_t)j * m + k];
            if (b == hc128(0, 0)) continue;
            for (int i = 0; i < m; ++i) C[(size_t)j * m + i] += A[(size_t)k * m + i] * b;
        }
}

// One pass of the column-scaled (shifted) Cholesky-QR used by orthonormalize() (api.cu) on the host side:
// G (in: Gram matrix V^H V, column-major; destroyed) -> Ri = D^-1 R^-1 (the block update V_new = V Ri) and RD = R D
// (V = V_new RD).  Returns true when V is already orthonormal to rounding (nothing to apply).
// While the scaled Gram matrix is far from the identity, or its factorisation meets a pivot below the shift level, the
// factor is that of G + s I with s = 11 (m n + m (m + 1)) u ||V D^-1||_2^2 (<= m): shifted Cholesky-QR (Fukaya et al. 2020).
inline bool cholqr_pass(int m, double n_rows, std::vector<hc128>& G, std::vector<hc128>& Ri, std::vector<hc128>& RD,
                        double& prev_err) {
    const double u = 1.1e-16;
    const double shift = 11.0 * ((double)m * n_rows + (double)m * (m + 1.0)) * u * (double)m;
    // err = ||V^H V - I||_max of the incoming block.  Converged when it is at the rounding level of an m-term Gram
    // entry, or when it has stopped contracting below 1e-12: the Gram matrix of n rows is itself only accurate to
    // ~sqrt(n) u, a further pass cannot improve on that (prev_err: the caller's state, start with a negative value).
    double err = 0.0;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) err = std::max(err, std::abs(G[(size_t)j * m + i] - (i == j ? hc128(1, 0) : hc128(0, 0))));
    const bool stalled = prev_err >= 0.0 && err <= 1e-12 && err >= 0.25 * prev_err;
    prev_err = err;
    if (err <= 8e-16 * std::sqrt((double)m) + 4e-16 || stalled) return true;
    std::vector<double> d(m);
    for (int j = 0; j < m; ++j) {
        const double g = G[(size_t)j * m + j].real();
        d[j] = (g > 0.0 && std::isfinite(g)) ? std::sqrt(g) : 1.0;
    }
    double dev = 0.0;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
            hc128 g = G[(size_t)j * m + i] / (d[i] * d[j]);
            G[(size_t)j * m + i] = g;
            dev = std::max(dev, std::abs(g - (i == j ? hc128(1, 0) : hc128(0, 0))));
        }
    std::vector<hc128> R;
    bool ok = false;
    if (dev < 0.5) ok = chol_upper(m, G, R, shift);   // a pivot below the shift level: not safely positive definite
    if (!ok || !std::isfinite(dev)) {
        for (int j = 0; j < m; ++j) G[(size_t)j * m + j] += shift;
        chol_upper(m, G, R, 0.5 * shift);
    }
    triu_inverse(m, R, Ri);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Ri[(size_t)j * m + i] /= d[i];     // D^-1 R^-1
    RD = R;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) RD[(size_t)j * m + i] *= d[j];     // R D
    return false;
}
