// pimc_moves.cuh -- the update functors and estimators as device functions (reference: src/updates/*.jl, src/measurement.jl).
#pragma once
#include "pimc_device.cuh"

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pair action: interaction_action! (helper.jl:306-366) and the inlined copies of ReshapeSwapLinear (reshape.jl:166-199 old
// configuration, :209-240 new configuration)
static __device__ __noinline__ double d_pairs_old(const DevSys &S, int c, int p, int sl)
{
    // find_nns(s, p, sl, exceptions=[p]) with the STORED bin of p, then lnU of (bead, next bead) distances
    double w = 0.0;
    const int M = S.M;
    double x = S.r[RIDX(S, c, p, 0, sl)], y = S.dim > 1 ? S.r[RIDX(S, c, p, 1, sl)] : 0.0;
    int b = S.bins[VIDX(S, c, p, sl)];
    int pn = sl == M - 1 ? S.next[(size_t)c * S.N + p] : p, sn = (sl + 1) % M;
    double xn = S.r[RIDX(S, c, pn, 0, sn)], yn = S.dim > 1 ? S.r[RIDX(S, c, pn, 1, sn)] : 0.0;
    NbFirst F; d_nb_first(S, c, sl, b, F);
    PIMC_FOR_STENCIL(S, c, sl, F, {
        if (o != p && d_peuclid(S, ox, oy, x, y) <= S.cellw) {
            int onx = sl == M - 1 ? S.next[(size_t)c * S.N + o] : o;
            double oxn = S.r[RIDX(S, c, onx, 0, sn)], oyn = S.dim > 1 ? S.r[RIDX(S, c, onx, 1, sn)] : 0.0;
            double lu = d_lnU(S, d_distance(ox, x, S.L), d_distance(oy, y, S.L), d_distance(oxn, xn, S.L), d_distance(oyn, yn, S.L));
            for (int rep = S.mult[VIDX(S, c, o, sl)]; rep > 0; --rep) w += lu;
        }
    })
    return w;
}
static __device__ __noinline__ double d_pairs_new(const DevSys &S, int c, double x, double y, double xn, double yn, int sl, int e1, int e2, bool skip_next_exc)
{
    double w = 0.0;
    const int M = S.M;
    int sn = (sl + 1) % M;
    NbFirst F; d_nb_first(S, c, sl, d_bin(S, x, y), F);
    PIMC_FOR_STENCIL(S, c, sl, F, {
        if (o != e1 && o != e2 && d_peuclid(S, ox, oy, x, y) <= S.cellw) {
            int onx = sl == M - 1 ? S.next[(size_t)c * S.N + o] : o;
            if (!(skip_next_exc && (onx == e1 || onx == e2))) {
                double oxn = S.r[RIDX(S, c, onx, 0, sn)], oyn = S.dim > 1 ? S.r[RIDX(S, c, onx, 1, sn)] : 0.0;
                double lu = d_lnU(S, d_distance(ox, x, S.L), d_distance(oy, y, S.L), d_distance(oxn, xn, S.L), d_distance(oyn, yn, S.L));
                for (int rep = S.mult[VIDX(S, c, o, sl)]; rep > 0; --rep) w += lu;
            }
        }
    })
    return w;
}

// ---- ReshapeLinear body (reshape.jl:56-87), one thread.  n 0-based, j0 1-based, scratch slot `slot`.
// returns 1 accepted, 0 rejected, -1 bridge failed
__device__ __forceinline__ int d_reshape_linear(const DevSys &S, int c, int n, int j0, int m, const GSrc &g, double u, int commit,
                                                int slot, double *wi_out, double *wu_out)
{
    const int M = S.M, dim = S.dim;
    const int jm = j0 + m, rows = m + 1;
    const int nx = S.next[(size_t)c * S.N + n];
    // pcycle (helper.jl:113-115) for j0 <= j <= jm < 2M: slices beyond M belong to the next particle of the cycle
    const int pe = jm <= M ? n : nx, je = (jm <= M ? jm : jm - M) - 1;
    double bx = S.r[RIDX(S, c, n, 0, j0 - 1)], by = dim > 1 ? S.r[RIDX(S, c, n, 1, j0 - 1)] : 0.0;
    double ex = S.r[RIDX(S, c, pe, 0, je)], ey = dim > 1 ? S.r[RIDX(S, c, pe, 1, je)] : 0.0;
    double *px = S.prop + RIDX(S, c, slot, 0, 0), *py = px + M, *pv = S.propV + VIDX(S, c, slot, 0);
    double w_initial = 0.0, w_updated = 0.0;
    int ret = -1;
    if (d_bridge(S, c, bx, by, ex, ey, rows, j0, n, g, px, py, pv)) {
        const double mht = -0.5 * S.tau;
        double sv = 0.0;
        for (int jp = 1; jp <= m; ++jp) {
            int j = j0 + jp - 1;
            int p = j <= M ? n : nx, sl = (j <= M ? j : j - M) - 1;
            w_initial += S.Vl[VIDX(S, c, p, sl)];
            if (S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE)) { // intended mode: interaction_action! does count (reshape.jl:72,75)
                w_initial += d_pairs_old(S, c, p, sl);
                w_updated += d_pairs_new(S, c, px[jp - 1], dim > 1 ? py[jp - 1] : 0.0, px[jp], dim > 1 ? py[jp] : 0.0, sl, p, -1, true);
            }
            double vl = mht * (pv[jp - 1] + pv[jp]); // lnV (propagator.jl:26-28)
            pv[jp - 1] = vl;
            sv = jp == 1 ? vl : sv + vl;
        }
        w_updated += sv;
        ret = d_metropolis(pimc_exp(w_updated - w_initial), u) ? 1 : 0;
        if (ret == 1 && commit) {
            for (int jp = 1; jp <= m; ++jp) {
                int j = j0 + jp - 1;
                int p = j <= M ? n : nx, sl = (j <= M ? j : j - M) - 1;
                S.r[RIDX(S, c, p, 0, sl)] = px[jp - 1];
                if (dim > 1) S.r[RIDX(S, c, p, 1, sl)] = py[jp - 1];
                S.Vl[VIDX(S, c, p, sl)] = pv[jp - 1];
                d_cell_update(S, c, sl, p, px[jp - 1], dim > 1 ? py[jp - 1] : 0.0);
            }
        }
    }
    if (wi_out) *wi_out = w_initial;
    if (wu_out) *wu_out = w_updated;
    return ret;
}

// sampleparticles weight table (helper.jl:230-260), one thread, w[N]
__device__ __forceinline__ void d_swap_weights(const DevSys &S, int c, int n1, int j0, int m, double *w)
{
    const int M = S.M, N = S.N, dim = S.dim;
    const int jmw = (j0 + m - 1) % M; // mod1(j0 + m, M) - 1
    const bool wrap = j0 + m > M;
    const int n1next = wrap ? S.next[(size_t)c * N + n1] : n1;
    const double mt = m * S.tau;
    double ax = S.r[RIDX(S, c, n1, 0, j0 - 1)], ay = dim > 1 ? S.r[RIDX(S, c, n1, 1, j0 - 1)] : 0.0;
    double cx = S.r[RIDX(S, c, n1next, 0, jmw)], cy = dim > 1 ? S.r[RIDX(S, c, n1next, 1, jmw)] : 0.0;
    for (int i = 0; i < N; ++i) {
        int inext = wrap ? S.next[(size_t)c * N + i] : i;
        double t = d_lnK2(ax, ay, S.r[RIDX(S, c, inext, 0, jmw)], dim > 1 ? S.r[RIDX(S, c, inext, 1, jmw)] : 0.0, dim, S.lambda, mt, S.L);
        double y = d_lnK2(S.r[RIDX(S, c, i, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, i, 1, j0 - 1)] : 0.0, cx, cy, dim, S.lambda, mt, S.L);
        w[i] = pimc_exp(t + y);
    }
}
// StatsBase.sample(::Weights): t = u * sum(w), walk the cumulative sum ; returns 0-based index
__device__ __forceinline__ int d_sample_weighted(const double *w, int n, double u)
{
    double wsum = w[0];
    for (int i = 1; i < n; ++i) wsum = wsum + w[i];
    double t = u * wsum, cw = w[0]; int i = 0;
    while (cw < t && i < n - 1) { i += 1; cw += w[i]; }
    return i;
}

// ---- ReshapeSwapLinear body (reshape.jl:138-279), one thread; scratch slots 0 and 1.
// returns 1 accepted, 0 rejected, -1 bridge failed, -2 n1 == n2
__device__ __forceinline__ int d_reshape_swap(const DevSys &S, int c, int n1, int n2, int j0, int m, const GSrc &g1, const GSrc &g2,
                                              double u, int commit, double *wi_out, double *wu_out)
{
    if (n1 == n2) return -2;
    const int M = S.M, N = S.N, dim = S.dim, jm = j0 + m, rows = m + 1;
    int *nextc = S.next + (size_t)c * N;
    const int x1 = nextc[n1], x2 = nextc[n2];
    const bool wrap = jm > M;
    const int je = (wrap ? jm - M : jm) - 1;
    const int e1 = wrap ? x2 : n2; // r1 ends on the cycle of n2
    const int e2 = wrap ? x1 : n1;
    double *p1x = S.prop + RIDX(S, c, 0, 0, 0), *p1y = p1x + M, *p1v = S.propV + VIDX(S, c, 0, 0);
    double *p2x = S.prop + RIDX(S, c, 1, 0, 0), *p2y = p2x + M, *p2v = S.propV + VIDX(S, c, 1, 0);
    int b1 = d_bridge(S, c, S.r[RIDX(S, c, n1, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, n1, 1, j0 - 1)] : 0.0,
                      S.r[RIDX(S, c, e1, 0, je)], dim > 1 ? S.r[RIDX(S, c, e1, 1, je)] : 0.0, rows, j0, n2, g1, p1x, p1y, p1v);
    int b2 = d_bridge(S, c, S.r[RIDX(S, c, n2, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, n2, 1, j0 - 1)] : 0.0,
                      S.r[RIDX(S, c, e2, 0, je)], dim > 1 ? S.r[RIDX(S, c, e2, 1, je)] : 0.0, rows, j0, n1, g2, p2x, p2y, p2v);
    double w_initial = 0.0, w_updated = 0.0;
    int ret = -1;
    if (b1 && b2) {
        for (int j = j0; j <= jm - 1; ++j) {
            int q1 = j <= M ? n1 : x1, q2 = j <= M ? n2 : x2, sl = (j <= M ? j : j - M) - 1;
            w_initial += S.Vl[VIDX(S, c, q1, sl)] + S.Vl[VIDX(S, c, q2, sl)];
            if (S.interactions) { w_initial += d_pairs_old(S, c, q1, sl); w_initial += d_pairs_old(S, c, q2, sl); }
        }
        const double mht = -0.5 * S.tau;
        double s1 = 0.0, s2 = 0.0;
        for (int jp = 1; jp <= m; ++jp) {
            int j = j0 + jp - 1;
            double v1 = mht * (p1v[jp - 1] + p1v[jp]), v2 = mht * (p2v[jp - 1] + p2v[jp]);
            p1v[jp - 1] = v1; p2v[jp - 1] = v2;
            s1 = jp == 1 ? v1 : s1 + v1; s2 = jp == 1 ? v2 : s2 + v2;
            if (S.interactions) {
                int q1 = j <= M ? n1 : x1, q2 = j <= M ? n2 : x2, sl = (j <= M ? j : j - M) - 1;
                double add = d_pairs_new(S, c, p1x[jp - 1], dim > 1 ? p1y[jp - 1] : 0.0, p1x[jp], dim > 1 ? p1y[jp] : 0.0, sl, q1, q2, false);
                add += d_pairs_new(S, c, p2x[jp - 1], dim > 1 ? p2y[jp - 1] : 0.0, p2x[jp], dim > 1 ? p2y[jp] : 0.0, sl, q1, q2, false);
                if (S.compat & PIMC_COMPAT_SWAP_SIGN) w_initial += add; else w_updated += add; // reshape.jl:224,239
            }
        }
        w_updated += s1 + s2;
        ret = d_metropolis(pimc_exp(w_updated - w_initial), u) ? 1 : 0;
        if (ret == 1 && commit) {
            nextc[n1] = x2; nextc[n2] = x1; // reshape.jl:254 ; the closures then see the re-computed cycles
            for (int jr = 2; jr <= m + 1; ++jr) {
                int j = j0 + jr - 1;
                int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
                S.r[RIDX(S, c, q1, 0, sl)] = p1x[jr - 1]; if (dim > 1) S.r[RIDX(S, c, q1, 1, sl)] = p1y[jr - 1];
                d_cell_update(S, c, sl, q1, p1x[jr - 1], dim > 1 ? p1y[jr - 1] : 0.0);
                S.r[RIDX(S, c, q2, 0, sl)] = p2x[jr - 1]; if (dim > 1) S.r[RIDX(S, c, q2, 1, sl)] = p2y[jr - 1];
                d_cell_update(S, c, sl, q2, p2x[jr - 1], dim > 1 ? p2y[jr - 1] : 0.0);
            }
            for (int jp = 1; jp <= m; ++jp) {
                int j = j0 + jp - 1;
                int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
                S.Vl[VIDX(S, c, q1, sl)] = p1v[jp - 1];
                S.Vl[VIDX(S, c, q2, sl)] = p2v[jp - 1];
            }
            if (jm < M) { // tails jm+1..M change owner (reshape.jl:269-275)
                for (int sl = jm; sl < M; ++sl) {
                    for (int k = 0; k < dim; ++k) {
                        double t = S.r[RIDX(S, c, n1, k, sl)]; S.r[RIDX(S, c, n1, k, sl)] = S.r[RIDX(S, c, n2, k, sl)]; S.r[RIDX(S, c, n2, k, sl)] = t;
                    }
                    double tv = S.Vl[VIDX(S, c, n1, sl)]; S.Vl[VIDX(S, c, n1, sl)] = S.Vl[VIDX(S, c, n2, sl)]; S.Vl[VIDX(S, c, n2, sl)] = tv;
                    if (S.need_cells) {
                        d_cell_update(S, c, sl, n1, S.r[RIDX(S, c, n1, 0, sl)], dim > 1 ? S.r[RIDX(S, c, n1, 1, sl)] : 0.0);
                        d_cell_update(S, c, sl, n2, S.r[RIDX(S, c, n2, 0, sl)], dim > 1 ? S.r[RIDX(S, c, n2, 1, sl)] : 0.0);
                    }
                }
            }
            if (!(S.compat & PIMC_COMPAT_SWAP_STALE_LINK) && jm <= M) { // intended: the link leaving slice j_m changes owner too
                double tv = S.Vl[VIDX(S, c, n1, jm - 1)]; S.Vl[VIDX(S, c, n1, jm - 1)] = S.Vl[VIDX(S, c, n2, jm - 1)]; S.Vl[VIDX(S, c, n2, jm - 1)] = tv;
            }
            if (S.need_cells) {
                // rm_nn!(old cycles) ... add_nn!(new pol1), add_nn!(new pol2) (reshape.jl:250-251,277-278): when the swap MERGED two
                // cycles the new pol1 and pol2 are the same set and every member is pushed twice into every slice's list
                bool merged = false; { int p = nextc[n1], cnt = 0; while (p != n1 && cnt <= N) { if (p == n2) merged = true; p = nextc[p]; cnt++; } }
                unsigned char mu = merged ? 2 : 1;
                for (int pass = 0; pass < 2; ++pass) {
                    int s0 = pass == 0 ? n1 : n2, p = s0, cnt = 0;
                    do { for (int sl = 0; sl < M; ++sl) S.mult[VIDX(S, c, p, sl)] = mu; p = nextc[p]; cnt++; } while (p != s0 && cnt <= N);
                }
            }
        }
    }
    if (wi_out) *wi_out = w_initial;
    if (wu_out) *wu_out = w_updated;
    return ret;
}

// ---- centre-of-mass move of the permutation cycle of n (com.jl:47-100 / :168-220, move_polymer! helper.jl:368-395).
// One WARP per task: lanes stride the slices, Delta-U reduced with warp shuffles.  Displacement source: explicit d or stream.
// returns 1 accepted, 0 rejected, -1 no admissible displacement
struct DSrc { const double *d; pimc_stream st; uint32_t slot; };
__device__ __forceinline__ int d_com_warp(const DevSys &S, int c, int n, double maxd, const DSrc &ds, double u, int commit,
                                          double *wi_out, double *wu_out, int *npol_out)
{
    const int lane = threadIdx.x & 31, M = S.M, N = S.N, dim = S.dim;
    const int *nextc = S.next + (size_t)c * N;
    // w_initial: cached links of every member of the cycle
    double w_initial = 0.0; int npol = 0;
    { int p = n; do { double part = 0.0;
            for (int j = lane; j < M; j += 32) {
                part += S.Vl[VIDX(S, c, p, j)];
                if (S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE)) part += d_pairs_old(S, c, p, j); // com.jl:54,173 (intended)
            }
            w_initial += warp_sum(part); npol += 1; p = nextc[p]; } while (p != n && npol <= N); }
    double dx = 0.0, dy = 0.0; bool ok = false;
    for (long long ctr = 1; ctr <= S.ctr; ++ctr) {
        if (ds.d) { if (ctr > 1) break; dx = ds.d[0]; dy = dim > 1 ? ds.d[1] : 0.0; }
        else {
            pimc_u4 w = pimc_draw(ds.st, ds.slot, PIMC_K_COM, (uint32_t)(ctr - 1), 0);
            dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5);
            dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        }
        bool hit = false;
        if (S.a > 0.0) {
            int p = n, cnt = 0;
            do { for (int j = lane; j < M; j += 32) {
                    double x = d_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? d_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    if (d_hardcore_hit(S, c, x, y, j, p)) hit = true;
                }
                p = nextc[p]; cnt++; } while (p != n && cnt <= N);
        }
        if (!__any_sync(0xffffffffu, hit)) { ok = true; break; }
    }
    int ret = -1; double w_updated = 0.0;
    if (ok) {
        const double mht = -0.5 * S.tau;
        double part = 0.0;
        int p = n, cnt = 0;
        do { int pn = nextc[p];
            for (int j = lane; j < M; j += 32) {
                int q = j == M - 1 ? pn : p, jn = j == M - 1 ? 0 : j + 1;
                double x = d_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? d_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                double xn = d_teleport(S.r[RIDX(S, c, q, 0, jn)] + dx, S.L), yn = dim > 1 ? d_teleport(S.r[RIDX(S, c, q, 1, jn)] + dy, S.L) : 0.0;
                part += mht * (d_pot(S.pot, x, y, dim) + d_pot(S.pot, xn, yn, dim));
                if (S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE)) part += d_pairs_new(S, c, x, y, xn, yn, j, p, -1, false); // com.jl:79,198
            }
            p = pn; cnt++; } while (p != n && cnt <= N);
        w_updated += warp_sum(part);
        ret = d_metropolis(pimc_exp(w_updated - w_initial), u) ? 1 : 0;
        if (ret == 1 && commit) {
            // link cache first (it reads the still-unshifted neighbours), then the positions
            p = n; cnt = 0;
            do { int pn = nextc[p];
                for (int j = lane; j < M; j += 32) {
                    int q = j == M - 1 ? pn : p, jn = j == M - 1 ? 0 : j + 1;
                    double x = d_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? d_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    double xn = d_teleport(S.r[RIDX(S, c, q, 0, jn)] + dx, S.L), yn = dim > 1 ? d_teleport(S.r[RIDX(S, c, q, 1, jn)] + dy, S.L) : 0.0;
                    S.Vl[VIDX(S, c, p, j)] = mht * (d_pot(S.pot, x, y, dim) + d_pot(S.pot, xn, yn, dim));
                }
                p = pn; cnt++; } while (p != n && cnt <= N);
            __syncwarp();
            p = n; cnt = 0;
            do { for (int j = lane; j < M; j += 32) {
                    double x = d_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? d_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    S.r[RIDX(S, c, p, 0, j)] = x; if (dim > 1) S.r[RIDX(S, c, p, 1, j)] = y;
                    d_cell_update(S, c, j, p, x, y);
                }
                p = nextc[p]; cnt++; } while (p != n && cnt <= N);
            __syncwarp();
        }
    }
    if (wi_out) *wi_out = w_initial;
    if (wu_out) *wu_out = w_updated;
    if (npol_out) *npol_out = npol;
    return ret;
}

// ---- Energy functor (measurement.jl:92-122), whole CTA; result valid in thread 0. red = 3 * 32 doubles of shared memory.
// A warp owns a worldline; every lane first issues the loads of EST_U beads and of their successors (independent, coalesced; the
// successor loads hit the lines just fetched), then does the arithmetic: inside the persistent kernel a chain has one or two warps
// and the estimator is bound by memory latency, so the loads in flight per lane are what counts.
#define EST_U 4
static __device__ __noinline__ double f_pot(const PotDev &p, double x, double y, int dim) { return d_pot(p, x, y, dim); }   // one copy (code size)
static __device__ __noinline__ double f_rdv(const PotDev &p, double x, double y, int dim) { return d_rdv(p, x, y, dim); }
__device__ __forceinline__ void d_energy_block(const DevSys &S, int c, double *red, double *E, double *Ev, double *parts)
{
    const int M = S.M, N = S.N, dim = S.dim;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double link = 0.0, pot = 0.0, vkin = 0.0;
    for (int i = w; i < N; i += nw) {
        const int in = S.next[(size_t)c * N + i];
        const double *xr = S.r + RIDX(S, c, i, 0, 0), *yr = xr + M;
        const double *xn = S.r + RIDX(S, c, in, 0, 0), *yn = xn + M;
        for (int jb = 0; jb < M; jb += 32 * EST_U) {
            double ax[EST_U], ay[EST_U], bx[EST_U], by[EST_U];
#pragma unroll
            for (int u = 0; u < EST_U; ++u) {
                const int j = jb + 32 * u + lane;
                ax[u] = ay[u] = bx[u] = by[u] = 0.0;
                if (j < M) {
                    ax[u] = xr[j]; bx[u] = j == M - 1 ? xn[0] : xr[j + 1];
                    if (dim > 1) { ay[u] = yr[j]; by[u] = j == M - 1 ? yn[0] : yr[j + 1]; }
                }
            }
#pragma unroll
            for (int u = 0; u < EST_U; ++u) {
                const int j = jb + 32 * u + lane;
                if (j < M) {
                    double dr = d_distance(ax[u], bx[u], S.L), d2 = dr * dr;
                    if (dim > 1) { dr = d_distance(ay[u], by[u], S.L); d2 = d2 + dr * dr; }
                    link += d2;
                    if (S.pot.kind != PIMC_POT_ZERO) pot += f_pot(S.pot, ax[u], ay[u], dim) + f_pot(S.pot, bx[u], by[u], dim);
                    if (S.pot.dv_kind != PIMC_DV_ZERO) vkin += f_rdv(S.pot, ax[u], ay[u], dim);
                }
            }
        }
    }
    link = warp_sum(link); pot = warp_sum(pot); vkin = warp_sum(vkin);
    __syncthreads();
    if (lane == 0) { red[w] = link; red[32 + w] = pot; red[64 + w] = vkin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        link = 0.0; pot = 0.0; vkin = 0.0;
        for (int i = 0; i < nw; ++i) { link += red[i]; pot += red[32 + i]; vkin += red[64 + i]; }
        *E = (double)(S.dim * S.N) / (2 * S.tau) - 1 / (4 * S.lambda * (S.tau * S.tau) * S.M) * link + 1.0 / (2 * S.M) * pot;
        *Ev = 1.0 / (2 * S.M) * vkin + 1.0 / (2 * S.M) * pot;
        if (parts) { parts[0] = link; parts[1] = pot; parts[2] = vkin; }
    }
}

// ---- Density functor (measurement.jl:45-55), whole CTA, integer counters in HBM (column-major like the Julia array).
// ibin = floor((r + L) / bin) (measurement.jl:47): evaluated as floor((r + L) * (1 / bin)), and re-evaluated with the true division
// whenever the product lies within a few ulps of an integer -- the only case in which the two floors can differ.  A warp owns a
// worldline and loads EST_U beads per lane before binning them (see d_energy_block).
__device__ __forceinline__ long long d_density_bin(double x, double L, double bin, double inv)
{
    const double s = x + L, t = s * inv;
    double f = floor(t);
    const double fr = t - f, thr = (fabs(t) + 1.0) * 4e-15;
    if (fr < thr || 1.0 - fr < thr) f = floor(s / bin);
    return (long long)f;
}
__device__ __forceinline__ void d_density_block(const DevSys &S, int c, const DeDev &D)
{
    const int M = S.M, N = S.N, dim = S.dim;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const bool shift = (S.compat & PIMC_COMPAT_DENSITY_SHIFT) != 0;
    const double inv = 1.0 / D.bin;
    for (int i = w; i < N; i += nw) {
        const double *xr = S.r + RIDX(S, c, i, 0, 0), *yr = xr + M;
        for (int jb = 0; jb < M; jb += 32 * EST_U) {
            double xv[EST_U], yv[EST_U];
#pragma unroll
            for (int u = 0; u < EST_U; ++u) {
                const int j = jb + 32 * u + lane;
                xv[u] = yv[u] = 0.0;
                if (j < M) { xv[u] = xr[j]; if (dim > 1) yv[u] = yr[j]; }
            }
#pragma unroll
            for (int u = 0; u < EST_U; ++u) {
                const int j = jb + 32 * u + lane;
                if (j < M) {
                    long long ib0 = d_density_bin(xv[u], S.L, D.bin, inv);
                    long long ib1 = dim > 1 ? d_density_bin(yv[u], S.L, D.bin, inv) : 1;
                    bool ok;
                    if (shift) ok = ib0 > 0 && ib0 < D.nbins + 1 && (dim == 1 || (ib1 > 0 && ib1 < D.nbins + 1));
                    else { ok = ib0 >= 0 && ib0 < D.nbins && (dim == 1 || (ib1 >= 0 && ib1 < D.nbins)); ib0 += 1; ib1 += 1; }
                    if (ok) atomicAdd(D.dens + (ib0 - 1) + (dim > 1 ? D.nbins * (ib1 - 1) : 0), 1ull);
                }
            }
        }
    }
}

// ---- Counter / adjust! (helper.jl:6-52, simulation.jl:1-10) ----
struct RingReg { int head, len, sum; long long tries; };
__device__ __forceinline__ void d_ring_push(const UpdDev &U, int c, RingReg &R, int acc)
{
    unsigned *ring = U.ring + (size_t)c * U.ring_words;
    long long cap = U.range + 1;
    R.tries += 1;
    long long pos = (long long)R.head + R.len; if (pos >= cap) pos -= cap;   // head < cap, len <= range = cap - 1
    unsigned bit = 1u << (pos & 31);
    if (acc) ring[pos >> 5] |= bit; else ring[pos >> 5] &= ~bit;
    R.len += 1; R.sum += acc ? 1 : 0;
    if (R.len > U.range) {
        R.sum -= (ring[R.head >> 5] >> (R.head & 31)) & 1u;
        R.head = R.head + 1 >= cap ? 0 : R.head + 1;
        R.len -= 1;
    }
}
__device__ __forceinline__ void d_adjust(const UpdDev &U, int c, const RingReg &R)
{
    double acc = (double)R.sum / (double)R.len; // empty window: 0/0 = NaN -> only the clamps act
    double v = U.var[c];
    if (U.kind == PIMC_UPD_RESHAPE_LINEAR || U.kind == PIMC_UPD_RESHAPE_SWAP) {
        if (acc < U.minacc) v -= 1; else if (acc > U.maxacc) v += 1;
    } else {
        if (acc < U.minacc) v *= 0.9; else if (acc > U.maxacc) v *= 1.1;
    }
    v = U.vmin > v ? U.vmin : v;
    v = U.vmax < v ? U.vmax : v;
    U.var[c] = v;
}
