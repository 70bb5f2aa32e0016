// pimc_propint.cu -- HOST code (no kernels): the pair-propagator term table and the neighbour cut-off of interacting Systems.
//
// Reference: prop_rel_interpolate_terms / build_prop_int (src/propagator.jl:34-89) tabulate on a 600 x 600 grid over [1e-20, L]^2
//     A(r1, r2) = -T1 - T2 + T3,    T_i = (1/2pi) Int_0^inf k exp(-tau k^2) w_i(k) B_i(k r1, k r2) dk
//     w_1 = w_3 = t^2/(1+t^2),  w_2 = t/(1+t^2),  t(k) = 1 / ((2/pi)(gamma + ln(k/2)) - 4/g0)
//     B_1 = J0 J0,  B_2 = J0(k r1) Y0(k r2) + J0(k r2) Y0(k r1),  B_3 = Y0 Y0
// with 3 x 360 000 adaptive Gauss-Kronrod quadratures (QuadGK, rtol 1e-11); determine_nnrange (src/system.jl:10-15) then finds the
// cut-off radius r_a with Optim + Roots.  Both run on the host in the reference (once per System) and on the host here: every integrand
// is a product f(k) u(k r1) v(k r2), so on a FIXED rule {k_q, w_q} the table is three matrix products
//     A = -J' W1 J - (J' W2 Y + Y' W2 J) + Y' W1 Y,   J = J0(k_q r_i), Y = Y0(k_q r_i)
// (16-point Gauss-Legendre panels, uniform with <= 10 rad phase advance per panel, dyadically graded towards k = 0 where ln k makes the
// integrand non-analytic, cut where exp(-tau k^2) < 1e-19; with D = 1/t the weights 1/(1+D^2), D/(1+D^2) are smooth through the pole
// of t).  Same rule as pimc_jl_b200/propint.py (numpy, kept as the independent cross-check: tests/test_propint_cpu.py holds both against
// adaptive quadrature and against each other).  The Julia shim calls these through ccall so that examples/density_SRL_lattice.jl:17-19
// (`propint = build_prop_int(...)`, System(...; interactions = true, propint) without r_a) runs unchanged.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/pimc_b200.h"

namespace {
const double EULER_GAMMA = 0.5772156649015329;

void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0); w.assign(n, 0.0);
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp; z -= dz;
            if (fabs(dz) < 1e-16) break;
        }
        {   // derivative at the converged node
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
        }
        x[i] = -z; x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

// composite rule on (0, kmax): nodes k, weights wk
void gl_rule(double tau, double rmax, std::vector<double> &k, std::vector<double> &wk)
{
    const int order = 16, levels = 48; const double phase = 10.0;
    const double kmax = sqrt(44.0 / tau);
    const double h = std::min(phase / (2.0 * rmax), kmax / 8.0);
    std::vector<double> x, w; gauss_legendre(order, x, w);
    std::vector<double> edges; edges.push_back(0.0);
    for (int j = levels; j >= 1; --j) edges.push_back(ldexp(h, -j));
    const int npan = (int)ceil((kmax - h) / h);
    for (int i = 0; i <= npan; ++i) edges.push_back(h + (kmax - h) * (double)i / (double)npan);
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    k.clear(); wk.clear();
    for (size_t p = 0; p + 1 < edges.size(); ++p) {
        const double a = edges[p], b = edges[p + 1], hw = 0.5 * (b - a), mid = 0.5 * (b + a);
        for (int q = 0; q < order; ++q) { k.push_back(hw * x[q] + mid); wk.push_back(hw * w[q]); }
    }
}

void parallel_for(int n, const std::function<void(int, int)> &body)
{
    unsigned hc = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(hc ? hc : 1, 32u); if (nt > n) nt = n; if (nt < 1) nt = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=, &body]() { body(t, nt); });
    for (auto &x : th) x.join();
}

// bilinear `terms(r1_norm, r2_norm)` (Interpolations.jl scale(interpolate(A, BSpline(Linear())), ...)); tab column-major n x n
double terms_lookup(const double *tab, int n, double lo, double hi, double x, double y)
{
    const double h = (hi - lo) / (n - 1);
    const double tx = (x - lo) / h, ty = (y - lo) / h;
    int ix = (int)floor(tx), iy = (int)floor(ty);
    ix = std::min(std::max(ix, 0), n - 2); iy = std::min(std::max(iy, 0), n - 2);
    const double fx = tx - ix, fy = ty - iy;
    const double c0 = (1 - fx) * tab[ix + (size_t)n * iy] + fx * tab[ix + 1 + (size_t)n * iy];
    const double c1 = (1 - fx) * tab[ix + (size_t)n * (iy + 1)] + fx * tab[ix + 1 + (size_t)n * (iy + 1)];
    return (1 - fy) * c0 + fy * c1;
}
double prop_int_eval(const double *tab, int n, double lo, double hi, const double *r1, const double *r2, int dim, double tau)
{
    double d2 = 0.0, n1 = 0.0, n2 = 0.0;
    for (int d = 0; d < dim; ++d) { const double q = r1[d] - r2[d]; d2 += q * q; n1 += r1[d] * r1[d]; n2 += r2[d] * r2[d]; }
    const double rel0 = exp(-d2 / (4 * tau)) / (4 * M_PI * tau);                       // prop_rel0 (propagator.jl:73-76)
    return 1 + terms_lookup(tab, n, lo, hi, sqrt(n1), sqrt(n2)) / rel0;                 // propagator.jl:82-86
}
}  // namespace

extern "C" int pimc_build_prop_table(double L, double g0, double tau, int32_t delta, double *tab, double *lo, double *hi)
{
    if (!(L > 0) || !(tau > 0) || g0 == 0.0 || delta < 2 || !tab) return PIMC_ERR_INVALID;
    const double r_lo = 1e-20;                                                          // propagator.jl:38-39
    const int n = delta;
    std::vector<double> r(n);
    for (int i = 0; i < n; ++i) r[i] = r_lo + (L - r_lo) * (double)i / (double)(n - 1); // range(1e-20, L, delta)
    r[n - 1] = L;
    std::vector<double> k, wk; gl_rule(tau, std::max(L, 1e-3), k, wk);
    const size_t Q = k.size();
    std::vector<double> w1(Q), w2(Q);
    for (size_t q = 0; q < Q; ++q) {
        const double D = (2.0 / M_PI) * (EULER_GAMMA + log(k[q] / 2.0)) - 4.0 / g0;     // 1 / tk(k)
        const double base = wk[q] * k[q] * exp(-tau * k[q] * k[q]) / (2.0 * M_PI);
        w1[q] = base / (1.0 + D * D); w2[q] = base * D / (1.0 + D * D);
    }
    // J[i][q], Y[i][q] (node index fastest: the products below are dot products over q)
    std::vector<double> J((size_t)n * Q), Y((size_t)n * Q), J1((size_t)n * Q), J2((size_t)n * Q), Y1((size_t)n * Q), Y2((size_t)n * Q);
    parallel_for(n, [&](int t, int nt) {
        for (int i = t; i < n; i += nt)
            for (size_t q = 0; q < Q; ++q) {
                const double a = k[q] * r[i], j = j0(a), y = y0(a);
                J[i * Q + q] = j; Y[i * Q + q] = y; J1[i * Q + q] = j * w1[q]; J2[i * Q + q] = j * w2[q]; Y1[i * Q + q] = y * w1[q]; Y2[i * Q + q] = y * w2[q];
            }
    });
    parallel_for(n, [&](int t, int nt) {
        for (int i = t; i < n; i += nt) {
            const double *j1 = &J1[i * Q], *j2 = &J2[i * Q], *y1 = &Y1[i * Q], *y2 = &Y2[i * Q];
            for (int j = i; j < n; ++j) {                                                // A is symmetric: J'W1J, Y'W1Y and the B_2 sum all are
                const double *jj = &J[j * Q], *yj = &Y[j * Q];
                double t1 = 0.0, t2a = 0.0, t2b = 0.0, t3 = 0.0;
                for (size_t q = 0; q < Q; ++q) { t1 += j1[q] * jj[q]; t2a += j2[q] * yj[q]; t2b += y2[q] * jj[q]; t3 += y1[q] * yj[q]; }
                const double s = -t1 - (t2a + t2b) + t3;
                tab[i + (size_t)n * j] = s; tab[j + (size_t)n * i] = s;
            }
        }
    });
    if (lo) *lo = r_lo; if (hi) *hi = L;
    return PIMC_OK;
}

extern "C" int pimc_prop_int(const double *tab, int32_t n, double lo, double hi, const double *r1_rel, const double *r2_rel, int32_t dim, double tau, double *out)
{
    if (!tab || n < 2 || !r1_rel || !r2_rel || dim < 1 || !out || !(tau > 0)) return PIMC_ERR_INVALID;
    *out = prop_int_eval(tab, n, lo, hi, r1_rel, r2_rel, dim, tau);
    return PIMC_OK;
}

// determine_nnrange(propint, tau, a, b) (system.jl:10-15): zero of p([r],[r]) - 0.999 right of its minimiser.  Optim's Nelder-Mead from
// x0 = a and this scan + golden section land in the same basin (the function has one minimum); Roots' bisection = the loop below.
extern "C" int pimc_determine_nnrange(const double *tab, int32_t n, double lo, double hi, double tau, double a, double b, double *r_a)
{
    if (!tab || n < 2 || !r_a || !(tau > 0) || !(b > a)) return PIMC_ERR_INVALID;
    auto f = [&](double r) { return prop_int_eval(tab, n, lo, hi, &r, &r, 1, tau) - 0.999; };
    const int NS = 2001; const double x0 = std::max(a, lo);
    int best = 0; double fbest = 0.0;
    auto xs = [&](int i) { return i == NS - 1 ? b : x0 + (b - x0) * (double)i / (double)(NS - 1); };
    for (int i = 0; i < NS; ++i) { const double v = f(xs(i)); if (i == 0 || v < fbest) { fbest = v; best = i; } }
    double l = xs(std::max(best - 1, 0)), h = xs(std::min(best + 1, NS - 1));
    const double gr = (sqrt(5.0) - 1) / 2;
    for (int it = 0; it < 80; ++it) {
        const double c = h - gr * (h - l), d = l + gr * (h - l);
        if (f(c) < f(d)) h = d; else l = c;
    }
    const double rmin = 0.5 * (l + h), fa = f(rmin), fb = f(b);
    if (!((fa < 0.0 && 0.0 <= fb) || (fb < 0.0 && 0.0 <= fa))) return PIMC_ERR_STATE;   // Roots.find_zero would throw
    l = rmin; h = b;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (l + h);
        if ((f(mid) < 0.0) == (fa < 0.0)) l = mid; else h = mid;
        if (h - l <= 1e-15 * std::max(1.0, fabs(h))) break;
    }
    *r_a = 0.5 * (l + h);
    return PIMC_OK;
}
