// Host-side helpers for m0 x m0 complex matrices (column-major).  m0 <= a few hundred, so
// plain loops are fine; the library deliberately links no LAPACK (none is installed
// system-wide, and the reduced eigenproblem belongs to the caller).
#pragma once
#include <complex>
#include <vector>
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
