// klda_host.hpp -- host side of the iterative KLDA solve (klda.cu): plain C++, no CUDA, so that the very
// same code is compiled and tested on the CPU (tests/host/klda_host_main.cpp, tests/test_klda_host_math.py).
#pragma once
#include <math.h>
#include <vector>

namespace fx_klda_host {

// ---- host side of the Lanczos-Galerkin solve (tiny, O(m) per column) ---------------------------
// Galerkin system of A = cw K^2 + shift I on the Krylov space of K:  with K Q = Q T + beta_m q_{m+1} e_m',
//   Q'AQ = cw (T^2 + beta_m^2 e_m e_m') + shift I   (pentadiagonal, SPD whatever the definiteness of K),
// solved by a banded LDL'.  al[0..m], be[0..m] hold alpha_1..alpha_{m+1}, beta_1..beta_{m+1}: the look-ahead
// pair gives the exact residual norm of the Galerkin solution,
//   ||b - A Q z|| = cw beta_m sqrt((e_m'T z + alpha_{m+1} z_m)^2 + (beta_{m+1} z_m)^2).
inline double galerkin_penta_solve(const std::vector<double>& al, const std::vector<double>& be, int m, double cw,
                            double shift, double rhs0, std::vector<double>& z) {
    typedef long double ld_t;
    std::vector<ld_t> d0(m), d1(m, 0), d2(m, 0), D(m), L1(m, 0), L2(m, 0), y(m, 0);
    for (int i = 0; i < m; ++i) {
        const ld_t bl = i > 0 ? be[i - 1] : 0.0, br = i < m - 1 ? be[i] : 0.0, a = al[i];
        d0[i] = (ld_t)cw * (a * a + bl * bl + br * br) + shift;
        if (i + 1 < m) d1[i] = (ld_t)cw * be[i] * ((ld_t)al[i] + al[i + 1]);
        if (i + 2 < m) d2[i] = (ld_t)cw * be[i] * be[i + 1];
    }
    d0[m - 1] += (ld_t)cw * be[m - 1] * be[m - 1];
    for (int i = 0; i < m; ++i) {
        ld_t v = d0[i];
        if (i >= 1) v -= L1[i - 1] * L1[i - 1] * D[i - 1];
        if (i >= 2) v -= L2[i - 2] * L2[i - 2] * D[i - 2];
        D[i] = v;
        if (i + 1 < m) {
            ld_t t = d1[i];
            if (i >= 1) t -= L1[i - 1] * L2[i - 1] * D[i - 1];
            L1[i] = t / v;
        }
        if (i + 2 < m) L2[i] = d2[i] / v;
    }
    y[0] = rhs0;
    for (int i = 1; i < m; ++i) y[i] = -L1[i - 1] * y[i - 1] - (i >= 2 ? L2[i - 2] * y[i - 2] : (ld_t)0);
    for (int i = 0; i < m; ++i) y[i] /= D[i];
    for (int i = m - 2; i >= 0; --i) y[i] -= L1[i] * y[i + 1] + (i + 2 < m ? L2[i] * y[i + 2] : (ld_t)0);
    z.resize(m);
    for (int i = 0; i < m; ++i) z[i] = (double)y[i];
    const ld_t tz = (m >= 2 ? (ld_t)be[m - 2] * y[m - 2] : (ld_t)0) + (ld_t)al[m - 1] * y[m - 1];
    const ld_t r1 = tz + (ld_t)al[m] * y[m - 1], r2 = (ld_t)be[m] * y[m - 1];
    return (double)((ld_t)cw * be[m - 1] * sqrtl(r1 * r1 + r2 * r2));
}

// X = W^-1 B for a c x c symmetric positive definite W (column-major, ld c), B c x nb; false if not SPD
inline bool host_spd_solve(std::vector<double> W, int c, std::vector<double>& B, int nb) {
    for (int j = 0; j < c; ++j) {
        long double d = W[j + j * c];
        for (int k = 0; k < j; ++k) d -= (long double)W[j + k * c] * W[j + k * c];
        if (!(d > 0)) return false;
        const double dj = (double)sqrtl(d);
        W[j + j * c] = dj;
        for (int i = j + 1; i < c; ++i) {
            long double t = W[i + j * c];
            for (int k = 0; k < j; ++k) t -= (long double)W[i + k * c] * W[j + k * c];
            W[i + j * c] = (double)(t / dj);
        }
    }
    for (int r = 0; r < nb; ++r) {
        double* b = B.data() + (size_t)r * c;
        for (int i = 0; i < c; ++i) {
            long double t = b[i];
            for (int k = 0; k < i; ++k) t -= (long double)W[i + k * c] * b[k];
            b[i] = (double)(t / W[i + i * c]);
        }
        for (int i = c - 1; i >= 0; --i) {
            long double t = b[i];
            for (int k = i + 1; k < c; ++k) t -= (long double)W[k + i * c] * b[k];
            b[i] = (double)(t / W[i + i * c]);
        }
    }
    return true;
}

}  // namespace fx_klda_host
