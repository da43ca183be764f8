// Eigenvalues of a COMPRESSION of a diagonal matrix, and projections onto its eigenvectors, in O(q n^2):
//
//     T = (I - Q Q^T) D (I - Q Q^T),   D = diag(d_1 .. d_n),   Q (n x q) orthonormal,
//
// which is what EMMA's  eigen(S (K + I) S)  (reference: R/emma_eigen_R_wo_Z.R:7-20, S = I - X (X^T X)^-1 X^T) becomes in
// the basis of eigen(K): K never changes after the first forward iteration (R/AM.R:414-423), so its eigenvectors are
// computed ONCE and every later iteration only needs the n - q non-trivial eigenvalues of T and eta = (eigenvectors)^T y
// (all that R/emma_REMLE.R:40-76 and R/emma_MLE.R:27-56 consume).  The reference pays a dense n^3 eigendecomposition per
// iteration for them (the author's "eigen calculation in emma.REMLE is a bottleneck", MyPackage/MyREADME:1-4).
//
// Method: the projector is applied one direction at a time.  For a unit vector z,  (I - z z^T) D (I - z z^T)  has the
// eigenvalue 0 (vector z) and the n - 1 roots of the secular function
//         f(lambda) = sum_i z_i^2 / (d_i - lambda),
// one in every gap (d_j, d_j+1), with eigenvectors  x_j = (D - lambda_j)^-1 z / |.|  -- the limit rho -> infinity of the
// rank-one update that the divide-and-conquer eigensolver (LAPACK dlaed2/3/4; Gu & Eisenstat 1994) is built on.  In the
// new eigenbasis the compressed matrix is diagonal again, the remaining directions (and y) are carried along by
// Cauchy-like matrix-vector products  sum_i zhat_i v_i / (d_i - lambda_j),  and the next direction is treated the same
// way.  Three O(m^2) data-parallel pieces per direction (roots, Loewner weights, transforms) and O(m) host bookkeeping
// (sorting, deflation).  Numerical care follows the published algorithm:
//   * every root is stored as (origin pole, offset mu) with the origin the nearer of its two neighbouring poles, and all
//     differences d_i - lambda_j are formed as (d_i - d_origin) - mu: no cancellation, high relative accuracy;
//   * roots by safeguarded rational ("middle way") interpolation with the two neighbouring poles kept exact;
//   * deflation of negligible components of z and of (nearly) equal poles by Givens rotations, tolerance 8 eps max|d|;
//   * eigenvectors from the Loewner weights zhat (Gu & Eisenstat): with them the computed roots are the EXACT eigenvalues
//     of a nearby compression, so the vectors are orthogonal to working precision whatever the accuracy of the roots.
//
// This header holds the arithmetic (host + device) and the host orchestration as a template over a back end that
// supplies the three data-parallel pieces: csrc/eigbasis.cu runs them as CUDA kernels (the product); the CPU unit tests
// compile the same header with a loop back end (tests/csrc/secular_host.cpp) to check the numerics where no GPU exists.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#ifdef __CUDACC__
#define EG_HD __host__ __device__ __forceinline__
#else
#define EG_HD inline
#endif

namespace eg {
namespace sec {

constexpr double kEps = 2.220446049250313e-16;

struct Sums {
    double psi, dpsi;  // sum over the poles left of the gap (terms < 0) and its derivative
    double phi, dphi;  // right of the gap (terms > 0)
    double asum;       // sum of |terms|: the scale of the rounding error of psi + phi
};
EG_HD Sums sums_zero() { return Sums{0.0, 0.0, 0.0, 0.0, 0.0}; }
EG_HD Sums sums_add(const Sums& a, const Sums& b) {
    return Sums{a.psi + b.psi, a.dpsi + b.dpsi, a.phi + b.phi, a.dphi + b.dphi, a.asum + b.asum};
}
// one pole: z2 = z_i^2, del = d_i - d_origin, left = (i <= j)
EG_HD void sums_term(Sums& s, bool left, double z2, double del, double mu) {
    const double inv = 1.0 / (del - mu);
    const double t = z2 * inv, r = t * inv;
    if (left) { s.psi += t; s.dpsi += r; } else { s.phi += t; s.dphi += r; }
    s.asum += fabs(t);
}

// Root of f in the gap (d_j, d_j+1) of width `gap`.  eval(origin, mu) returns the Sums of f at lambda = d[origin] + mu
// (collectively on the device: every thread of the block calls it with the same arguments and receives the same
// result, so the scalar logic below runs redundantly and uniformly).  Result: origin in {j, j+1}, mu.
template <class Eval>
EG_HD void find_root(int j, double gap, Eval&& eval, int& origin_out, double& mu_out, int& iters_out) {
    Sums s = eval(j, 0.5 * gap);
    double f = s.psi + s.phi;
    int origin;
    double lo, hi, mu;
    if (f >= 0.0) { origin = j; lo = 0.0; hi = 0.5 * gap; mu = hi; }
    else { origin = j + 1; lo = -0.5 * gap; hi = 0.0; mu = lo; s = eval(origin, mu); f = s.psi + s.phi; }
    const double pl = origin == j ? 0.0 : -gap, pr = origin == j ? gap : 0.0;  // the two poles, relative to the origin
    int it = 0;
    for (; it < 120; ++it) {
        if (fabs(f) <= 4.0 * kEps * s.asum) break;
        if (f < 0.0) lo = mu; else hi = mu;
        const double width = hi - lo;
        if (width <= 2.0 * kEps * fmax(fabs(lo), fabs(hi))) { mu = f < 0.0 ? hi : lo; break; }
        // osculatory model  a + b / (D1 - eta) + c / (D2 - eta)  with the left sum attributed to pole j and the right sum
        // to pole j+1 (value and slope of each side matched):  A eta^2 - B eta + C = 0
        const double D1 = pl - mu, D2 = pr - mu;  // D1 < 0 < D2
        const double b = s.dpsi * D1 * D1, c = s.dphi * D2 * D2;
        const double a = f - s.dpsi * D1 - s.dphi * D2;
        const double B = a * (D1 + D2) + b + c, C = D1 * D2 * f;
        double eta = 0.0;
        bool ok = false;
        const double disc = B * B - 4.0 * a * C;
        if (disc >= 0.0) {
            const double q = 0.5 * (B + (B >= 0.0 ? sqrt(disc) : -sqrt(disc)));
            const double e1 = q != 0.0 ? C / q : 0.0, e2 = a != 0.0 ? q / a : e1;
            if (q != 0.0 && e1 > D1 && e1 < D2) { eta = e1; ok = true; }
            else if (a != 0.0 && e2 > D1 && e2 < D2) { eta = e2; ok = true; }
        }
        double next = mu + eta;
        const bool force_bisect = it >= 12 && (it & 1);   // the model converges in a handful of steps; if not, halve
        if (!ok || !(next > lo && next < hi) || force_bisect) {
            if (lo > 0.0 && hi > 16.0 * lo) next = sqrt(lo) * sqrt(hi);          // geometric: the root may be tiny
            else if (hi < 0.0 && lo < 16.0 * hi) next = -sqrt(-lo) * sqrt(-hi);
            else if (lo == 0.0 && force_bisect) next = hi * (1.0 / 1024.0);
            else if (hi == 0.0 && force_bisect) next = lo * (1.0 / 1024.0);
            else next = 0.5 * (lo + hi);
        }
        mu = next;
        s = eval(origin, mu);
        f = s.psi + s.phi;
    }
    origin_out = origin;
    mu_out = mu;
    iters_out = it;
}

// Loewner weight of pole i:  zhat_i^2 = prod_j (lambda_j - d_i) / prod_{k != i} (d_k - d_i), paired so that every factor
// lies in [0, 1]:  j < i with pole j, j >= i with pole j + 1.  One factor (root j), for pole i.
EG_HD double lowner_factor(int i, int j, const double* d, const int* origin, const double* mu) {
    const double num = (d[origin[j]] - d[i]) + mu[j];  // lambda_j - d_i (exactly mu_j when i is the root's origin)
    const double den = (j < i ? d[j] : d[j + 1]) - d[i];
    return num / den;
}
// component i of the (unnormalised) eigenvector of root j
EG_HD double vec_comp(double zhat_i, double d_i, double d_origin, double mu_j) { return zhat_i / ((d_i - d_origin) - mu_j); }

// ------------------------------------------------------------------------------------------------ host orchestration
// Back end:  int solve(int m, const double* d, const double* z, const double* V, int r, int* origin, double* mu,
//                      double* Vout, int* max_iters)
//   d[m] strictly increasing, z[m] (all |z_i| above the deflation tolerance, any norm), V: m x r column-major;
//   -> origin / mu of the m - 1 roots (root j in (d_j, d_j+1)) and Vout ((m-1) x r column-major): row j = x_j^T V.
struct Stats {
    int steps = 0, deflated = 0, max_iters = 0;
    int64_t roots = 0;
};

// xi[n]: eigenvalues of K (any order).  Xt: n x q column-major = U^T X;  yt[n] = U^T y  (U = eigenvectors of K, in the
// order of xi).  out_values[n-q]: the non-trivial eigenvalues of (I - QQ^T) diag(xi) (I - QQ^T), Q = orth(Xt), in
// DEcreasing order (= eigen(S (K+I) S)$values[1:(n-q)] - 1);  out_etas[n-q]: the matching (eigenvector)^T y, up to sign.
// Returns 0, or 1 when a column of X is (numerically) in the span of the earlier ones.
template <class Backend>
int compress(Backend& be, int64_t n64, int q, const double* xi, const double* Xt, const double* yt, double* out_values,
             double* out_etas, Stats* stats) {
    const int n = (int)n64;
    const int r0 = q + 1;  // carried columns: X's and y (last)
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return xi[a] < xi[b]; });
    int m = n;
    std::vector<double> d(m), V((size_t)m * r0);
    for (int i = 0; i < m; i++) {
        d[i] = xi[perm[i]];
        for (int c = 0; c < q; c++) V[i + (size_t)c * m] = Xt[perm[i] + (size_t)c * n];
        V[i + (size_t)q * m] = yt[perm[i]];
    }
    std::vector<double> col_norm0(q);
    for (int c = 0; c < q; c++) {
        double s = 0;
        for (int i = 0; i < n; i++) s += Xt[i + (size_t)c * n] * Xt[i + (size_t)c * n];
        col_norm0[c] = sqrt(s);
    }
    Stats st;
    for (int k = 0; k < q; k++) {
        const int r = q - k;  // columns carried beyond this one (the last is y)
        // direction: column 0 of the current V (earlier directions have already been projected out of it)
        std::vector<double> z(V.begin(), V.begin() + m);
        double nz = 0;
        for (int i = 0; i < m; i++) nz += z[i] * z[i];
        nz = sqrt(nz);
        if (!(nz > 1e-10 * col_norm0[k]) || !(nz > 0.0)) return 1;
        for (int i = 0; i < m; i++) z[i] /= nz;
        std::vector<double> W((size_t)m * r);  // the carried columns (rows follow d)
        for (int c = 0; c < r; c++) std::copy(V.begin() + (size_t)(c + 1) * m, V.begin() + (size_t)(c + 2) * m, W.begin() + (size_t)c * m);
        double dmax = 0;
        for (int i = 0; i < m; i++) dmax = std::max(dmax, fabs(d[i]));
        const double tol = 8.0 * kEps * std::max(dmax, 1e-300);
        // ---- deflation (dlaed2's scheme): walk the sorted poles; `p` is the last pole still carrying weight
        std::vector<char> defl(m, 0);
        int p = -1;
        for (int i = 0; i < m; i++) {
            if (fabs(z[i]) * dmax <= tol) { defl[i] = 1; z[i] = 0.0; continue; }   // negligible component: (d_i, e_i) is an eigenpair
            if (p < 0) { p = i; continue; }
            // poles p and i: rotate the weight of p into i when the off-diagonal this creates is negligible
            const double tau = hypot(z[p], z[i]);
            const double c = z[i] / tau, s = -z[p] / tau;
            if (fabs((d[i] - d[p]) * c * s) <= tol) {
                z[i] = tau;
                z[p] = 0.0;
                for (int cc = 0; cc < r; cc++) {
                    double& vp = W[p + (size_t)cc * m];
                    double& vi = W[i + (size_t)cc * m];
                    const double a = vp, b2 = vi;
                    vp = c * a + s * b2;    // row p <- the rotated basis vector that carries no weight
                    vi = -s * a + c * b2;
                }
                const double dp = d[p] * c * c + d[i] * s * s, di = d[p] * s * s + d[i] * c * c;
                d[p] = dp;
                d[i] = di;
                defl[p] = 1;
                p = i;
            } else {
                p = i;
            }
        }
        std::vector<int> act;
        for (int i = 0; i < m; i++)
            if (!defl[i]) act.push_back(i);
        const int ma = (int)act.size();
        st.deflated += m - ma;
        if (ma < 1) return 1;
        // active poles must be strictly increasing for the root finder (rotations keep the order; guard anyway)
        std::vector<double> da(ma), za(ma), Wa((size_t)ma * r);
        for (int a = 0; a < ma; a++) {
            da[a] = d[act[a]];
            za[a] = z[act[a]];
            for (int cc = 0; cc < r; cc++) Wa[a + (size_t)cc * ma] = W[act[a] + (size_t)cc * m];
        }
        for (int a = 1; a < ma; a++)
            if (!(da[a] > da[a - 1])) return 2;
        std::vector<int> origin(std::max(ma - 1, 1));
        std::vector<double> mu(std::max(ma - 1, 1)), Wout((size_t)std::max(ma - 1, 1) * r);
        if (ma > 1) {
            int iters = 0;
            const int rc = be.solve(ma, da.data(), za.data(), Wa.data(), r, origin.data(), mu.data(), Wout.data(), &iters);
            if (rc) return 100 + rc;
            st.max_iters = std::max(st.max_iters, iters);
            st.roots += ma - 1;
        }
        // ---- new diagonal: deflated poles keep their value and their (rotated) rows; the roots bring theirs
        const int m1 = m - 1;
        std::vector<double> d1(m1), V1((size_t)m1 * r);
        std::vector<std::pair<double, int>> order;  // (value, source): source >= 0 deflated row, < 0 root -1-j
        order.reserve(m1);
        for (int i = 0; i < m; i++)
            if (defl[i]) order.emplace_back(d[i], i);
        for (int j = 0; j + 1 < ma; j++) order.emplace_back(da[origin[j]] + mu[j], -1 - j);
        std::stable_sort(order.begin(), order.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });
        for (int t = 0; t < m1; t++) {
            d1[t] = order[t].first;
            const int src = order[t].second;
            for (int cc = 0; cc < r; cc++)
                V1[t + (size_t)cc * m1] = src >= 0 ? W[src + (size_t)cc * m] : Wout[(-1 - src) + (size_t)cc * (ma - 1)];
        }
        d.swap(d1);
        V.swap(V1);
        m = m1;
        st.steps++;
    }
    for (int t = 0; t < m; t++) {  // decreasing, as R's eigen()
        out_values[t] = d[m - 1 - t];
        out_etas[t] = V[m - 1 - t];  // one column left: y
    }
    if (stats) *stats = st;
    return 0;
}

}  // namespace sec
}  // namespace eg
