// Host harness for featureextraction.jl_b200/csrc/klda_host.hpp (the host side of the iterative KLDA
// solve).  Test infrastructure, not product.  stdin:
//   "penta m cw shift rhs0" then m+1 alphas and m+1 betas  -> prints est then z[0..m)
//   "spd c nb" then c*c entries of W (column-major) and c*nb entries of B -> prints ok then X
#include <stdio.h>
#include <string.h>
#include "../../featureextraction.jl_b200/csrc/klda_host.hpp"

int main() {
    char what[16];
    while (scanf("%15s", what) == 1) {
        if (!strcmp(what, "penta")) {
            int m; double cw, shift, rhs0;
            if (scanf("%d %lf %lf %lf", &m, &cw, &shift, &rhs0) != 4) return 2;
            std::vector<double> al(m + 1), be(m + 1), z;
            for (auto& v : al) if (scanf("%lf", &v) != 1) return 2;
            for (auto& v : be) if (scanf("%lf", &v) != 1) return 2;
            const double est = fx_klda_host::galerkin_penta_solve(al, be, m, cw, shift, rhs0, z);
            printf("%.17g", est);
            for (double v : z) printf(" %.17g", v);
            printf("\n");
        } else if (!strcmp(what, "spd")) {
            int c, nb;
            if (scanf("%d %d", &c, &nb) != 2) return 2;
            std::vector<double> W((size_t)c * c), B((size_t)c * nb);
            for (auto& v : W) if (scanf("%lf", &v) != 1) return 2;
            for (auto& v : B) if (scanf("%lf", &v) != 1) return 2;
            const bool ok = fx_klda_host::host_spd_solve(W, c, B, nb);
            printf("%d", ok ? 1 : 0);
            for (double v : B) printf(" %.17g", v);
            printf("\n");
        } else {
            return 3;
        }
    }
    return 0;
}
