// Host harness for the device exp of the Gaussian kernel map (fx_common.cuh: fx_exp_nonpos is
// __host__ __device__).  Compiled with nvcc for the HOST only by tests/test_device_math_host.py and
// compared with expl() (x87 80-bit long double on x86-64).  Test infrastructure, not product.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../featureextraction.jl_b200/csrc/fx_common.cuh"

static double ulp_of(double y) {
    int e;
    frexp(y, &e);
    return ldexp(1.0, e - 53);
}

int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : 2000000;
    unsigned long long st = 88172645463325252ull;
    double worst = 0.0, worst_x = 0.0, worst_clamped = 0.0;
    long bad_monotone = 0;
    for (long i = 0; i < N; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;                 // xorshift64
        const double u = (double)(st >> 11) / 9007199254740992.0;      // [0, 1)
        double x;
        switch (i & 3) {
            case 0: x = -708.0 * u; break;                             // whole range
            case 1: x = -u; break;                                     // small distances
            case 2: x = -40.0 * u * u; break;                          // typical kernel arguments
            default: x = -(double)(st & 65535) * 0.010830424696905538 - 1e-9 * u; break;   // near reduction breakpoints
        }
        const double got = fx_exp_nonpos(x);
        const long double ref = expl((long double)x);
        if (x < -708.0) {                                              // clamped region: absolute error only
            const double aerr = (double)fabsl((long double)got - ref);
            if (aerr > worst_clamped) worst_clamped = aerr;
            continue;
        }
        const double err = (double)(fabsl((long double)got - ref) / (long double)ulp_of((double)ref));
        if (err > worst) { worst = err; worst_x = x; }
    }
    // a few exact / edge values
    const double e0 = fx_exp_nonpos(0.0), ep = fx_exp_nonpos(1e-16), elo = fx_exp_nonpos(-1e6);
    // monotone on a fine grid around a table boundary
    double prev = fx_exp_nonpos(-0.0108304246969 - 1e-7);
    for (int k = 1; k <= 2000; ++k) {
        const double v = fx_exp_nonpos(-0.0108304246969 - 1e-7 + k * 1e-10);
        if (v < prev) ++bad_monotone;
        prev = v;
    }
    printf("{\"n\": %ld, \"max_ulp\": %.4f, \"at\": %.17g, \"exp0\": %.17g, \"exp_tiny_pos\": %.17g, "
           "\"exp_minus_1e6\": %.17g, \"non_monotone\": %ld, \"max_abs_err_below_minus_708\": %.3g}\n",
           N, worst, worst_x, e0, ep, elo, bad_monotone, worst_clamped);
    return 0;
}
