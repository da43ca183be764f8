// TEST INFRASTRUCTURE: a plain-loop back end for the secular compression solver of
// eagleeverything_b200/csrc/secular.cuh, so that the CPU test suite can check the SAME arithmetic and host
// orchestration the CUDA back end (csrc/eigbasis.cu) runs, against numpy's dense eigensolver.  Not linked into
// libeaglegpu.so; built by tests/test_secular_cpu.py with g++.
#include <cstring>

#include "../../eagleeverything_b200/csrc/secular.cuh"

using namespace eg::sec;

struct LoopBackend {
    int solve(int m, const double* d, const double* z, const double* V, int r, int* origin, double* mu, double* Vout,
              int* max_iters) {
        std::vector<double> z2(m), zhat(m);
        for (int i = 0; i < m; i++) z2[i] = z[i] * z[i];
        int worst = 0;
        for (int j = 0; j + 1 < m; j++) {
            auto eval = [&](int o, double x) {
                Sums s = sums_zero();
                for (int i = 0; i < m; i++) sums_term(s, i <= j, z2[i], d[i] - d[o], x);
                return s;
            };
            int it = 0;
            find_root(j, d[j + 1] - d[j], eval, origin[j], mu[j], it);
            worst = std::max(worst, it);
        }
        for (int i = 0; i < m; i++) {
            double p = 1.0;
            for (int j = 0; j + 1 < m; j++) p *= lowner_factor(i, j, d, origin, mu);
            zhat[i] = z[i] < 0 ? -sqrt(p) : sqrt(p);
        }
        for (int j = 0; j + 1 < m; j++) {
            double nn = 0.0;
            std::vector<double> acc(r, 0.0);
            for (int i = 0; i < m; i++) {
                const double x = vec_comp(zhat[i], d[i], d[origin[j]], mu[j]);
                nn += x * x;
                for (int c = 0; c < r; c++) acc[c] += x * V[i + (size_t)c * m];
            }
            const double inv = 1.0 / sqrt(nn);
            for (int c = 0; c < r; c++) Vout[j + (size_t)c * (m - 1)] = acc[c] * inv;
        }
        *max_iters = worst;
        return 0;
    }
};

extern "C" int th_secular_compress(int64_t n, int q, const double* xi, const double* Xt, const double* yt, double* vals,
                                   double* etas, int64_t* stats4) {
    LoopBackend be;
    Stats st;
    const int rc = compress(be, n, q, xi, Xt, yt, vals, etas, &st);
    if (stats4) {
        stats4[0] = st.steps;
        stats4[1] = st.deflated;
        stats4[2] = st.max_iters;
        stats4[3] = st.roots;
    }
    return rc;
}
