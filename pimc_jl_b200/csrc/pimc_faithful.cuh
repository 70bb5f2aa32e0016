// pimc_faithful.cuh -- warp-cooperative bodies of the FAITHFUL schedule (one proposal per chain and run! iteration) for systems
// WITH a hard core / pair action / cell list: ReshapeLinear (reshape.jl:31-91), ReshapeSwapLinear (reshape.jl:123-283) and
// hardspherelevy! (helper.jl:141-181).  One warp executes one proposal:
//   * every independent piece (Gaussians of retry 0, teleports, potentials, hard-core tests, the pair sums of every slice, the
//     weight table of sampleparticles, the commit with its cell-list surgery -- one list per (chain, slice), so slices are
//     independent) is spread over the lanes;
//   * every floating-point SUM is then formed by lane 0 in the reference order from the staged terms, and every staged term is
//     produced by the same device function the one-thread bodies of pimc_moves.cuh call (d_pairs_old/new, d_hardcore_hit,
//     d_gauss, d_pot), so results are bit-identical to d_reshape_linear / d_reshape_swap and to the oracle;
//   * the hard core is handled SPECULATIVELY: the bridge is laid with the retry-0 Gaussians and no tests, all beads are tested
//     at once (a lane per bead), and only if a bead fails the tail from the first failing bead is redone by the serial
//     redraw loop of helper.jl:160-176 (executed uniformly by the warp).  Draws are addressed (include/pimc_rng.h), so the
//     speculation consumes exactly the reference's random numbers.
// Scratch per proposal (shared memory, or HBM when a chain's rows do not fit): w[N] | 2 bridges x { x, y, v, link, gx, gy }[M + 1].
#pragma once
#include "pimc_moves.cuh"
#include "pimc_launch.h"


// ---- out-of-line leaves.  The persistent reference-schedule kernel runs ONE warp per proposal through long straight-line code:
// with every helper inlined at every call site the kernel was 1.2 MB of SASS and the warps stalled on instruction fetch
// (ncu: 28.6 cycles of "no instruction" per issue, profiles/r01e_*).  One shared copy of each helper keeps the hot code of all
// proposal kinds within the instruction caches.  Same arithmetic, same bits.
static __device__ __noinline__ pimc_u4 f_draw(pimc_stream st, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead) { return pimc_draw(st, slot, kind, retry, bead); }
static __device__ __noinline__ double2 f_gauss(const GSrc &g, int dim, int bead, int retry) { double a, b; d_gauss(g, dim, bead, retry, a, b); return make_double2(a, b); }
static __device__ __noinline__ double f_teleport(double x, double L) { return d_teleport_q(x, L); }
static __device__ __noinline__ double f_exp(double x) { return pimc_exp(x); }
static __device__ __noinline__ double f_lnK2(double ax, double ay, double bx, double by, int dim, double tau, double lambda, double L) { return d_lnK2(ax, ay, bx, by, dim, tau, lambda, L); }
static __device__ __noinline__ bool f_hardcore_hit(const DevSys &S, int c, double x, double y, int j, int exc) { return d_hardcore_hit(S, c, x, y, j, exc); }
static __device__ __noinline__ void f_cell_update(const DevSys &S, int c, int j, int n, double x, double y) { d_cell_update(S, c, j, n, x, y); }

__device__ __forceinline__ int warp_min_int(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}

// hardspherelevy! (helper.jl:141-181); == levy! (helper.jl:118-139) when a == 0.  Every lane passes the same arguments.
// bb: x = bb, y = bb + R1, v = bb + 2 R1 (receives V(row)), gx = bb + 4 R1, gy = bb + 5 R1.  Returns 1, or 0 when a bead exhausted s.ctr.
static __device__ __noinline__ int d_bridge_w(const DevSys &S, int c, double bx, double by, double ex, double ey, int rows, int j0,
                                          int exc, const GSrc &g, double *bb, int R1)
{
    const int lane = threadIdx.x & 31, dim = S.dim, m = rows - 2;
    const double L = S.L;
    double *px = bb, *py = bb + R1, *pv = bb + 2 * R1, *gx = bb + 4 * R1, *gy = bb + 5 * R1;
    if (fabs(bx - ex) > L) ex += d_sign(bx) * (2 * L);
    if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * (2 * L);
    for (int j = 1 + lane; j <= m; j += 32) {              // retry-0 draws of every interior bead, scaled by sigma_j; alpha_j parked in v
        const double alpha = (double)(m + 1 - j) / (double)(m + 2 - j);
        const double sig = sqrt(2 * S.lambda * alpha * S.tau);
        const double2 gg = f_gauss(g, dim, j, 0); const double g0 = gg.x, g1 = gg.y;
        gx[j] = g0 * sig; gy[j] = dim > 1 ? g1 * sig : 0.0; pv[j] = alpha;
    }
    __syncwarp();
    // Speculate -> test -> repair, repeated: lay rows jstart..m with the retry-0 draws (helper.jl:128-135), teleport and test them all
    // at once (a lane per bead); if bead jf is the first to hit the hard core, redraw THAT bead serially (helper.jl:160-176) and
    // speculate again from jf + 1.  Beads behind jf always start from their retry-0 draw, exactly as the serial loop does.
    double sx0 = bx, sy0 = by;                             // un-teleported row jstart - 1
    int jstart = 1;
    if (lane == 0) { px[0] = f_teleport(bx, L); py[0] = dim > 1 ? f_teleport(by, L) : 0.0; px[rows - 1] = f_teleport(ex, L); py[rows - 1] = dim > 1 ? f_teleport(ey, L) : 0.0; }
    while (jstart <= m) {
        {
            double qx = sx0, qy = sy0;
            for (int j = jstart; j <= m; ++j) {
                const double alpha = pv[j], om = 1 - alpha;
                qx = alpha * qx + om * ex + gx[j];
                if (dim > 1) qy = alpha * qy + om * ey + gy[j];
                if (lane == 0) { px[j] = qx; py[j] = qy; }
            }
        }
        __syncwarp();
        for (int j = jstart + lane; j <= m; j += 32) {     // teleport (helper.jl:136-138)
            px[j] = f_teleport(px[j], L);
            py[j] = dim > 1 ? f_teleport(py[j], L) : 0.0;
        }
        __syncwarp();
        if (!(S.a > 0.0)) break;
        int jf = 0x7fffffff;
        for (int j = jstart + lane; j <= m; j += 32) {
            const int sl = (j0 + j - 1) % S.M;             // mod1(j0 + j, M) - 1
            if (f_hardcore_hit(S, c, px[j], py[j], sl, exc) && j < jf) jf = j;
        }
        jf = warp_min_int(jf);
        if (jf > m) break;
        for (int j = jstart; j < jf; ++j) {                // un-teleported row jf - 1
            const double alpha = pv[j], om = 1 - alpha;
            sx0 = alpha * sx0 + om * ex + gx[j];
            if (dim > 1) sy0 = alpha * sy0 + om * ey + gy[j];
        }
        {                                                  // bead jf: its retry-0 draw is known to hit; redraw (warp-uniform)
            const double alpha = pv[jf], om = 1 - alpha;
            const double sig = sqrt(2 * S.lambda * alpha * S.tau);
            const int sl = (j0 + jf - 1) % S.M;
            double nx = 0.0, ny = 0.0, tx = 0.0, ty = 0.0; long long ctr = 1; bool pass = true;
            while (pass) {
                pass = false; ctr += 1;
                if (ctr > S.ctr) { pass = true; break; }
                const double2 gg = f_gauss(g, dim, jf, (int)(ctr - 1));
                nx = alpha * sx0 + om * ex + gg.x * sig;
                if (dim > 1) ny = alpha * sy0 + om * ey + gg.y * sig;
                tx = f_teleport(nx, L); ty = dim > 1 ? f_teleport(ny, L) : 0.0;
                if (f_hardcore_hit(S, c, tx, ty, sl, exc)) pass = true;
            }
            if (pass) return 0;
            sx0 = nx; sy0 = ny;
            __syncwarp();
            if (lane == 0) { px[jf] = tx; py[jf] = ty; }
        }
        jstart = jf + 1;
        __syncwarp();
    }
    __syncwarp();
    for (int row = lane; row < rows; row += 32) pv[row] = f_pot(S.pot, px[row], py[row], dim);
    __syncwarp();
    return 1;
}

// ---- ReshapeLinear body (reshape.jl:56-87), one warp.  n 0-based, j0 1-based; scr = FA_ARR * (M + 1) doubles.
// returns 1 accepted, 0 rejected, -1 bridge failed (identical on every lane)
__device__ __forceinline__ int d_reshape_linear_w(const DevSys &S, int c, int n, int j0, int m, const GSrc &g, double u, double *scr)
{
    const int lane = threadIdx.x & 31, M = S.M, dim = S.dim, R1 = M + 1;
    const int jm = j0 + m, rows = m + 1;
    const int nx = S.next[(size_t)c * S.N + n];
    const int pe = jm <= M ? n : nx, je = (jm <= M ? jm : jm - M) - 1;      // pcycle (helper.jl:113-115)
    const double bx = S.r[RIDX(S, c, n, 0, j0 - 1)], by = dim > 1 ? S.r[RIDX(S, c, n, 1, j0 - 1)] : 0.0;
    const double ex = S.r[RIDX(S, c, pe, 0, je)], ey = dim > 1 ? S.r[RIDX(S, c, pe, 1, je)] : 0.0;
    double *px = scr, *py = scr + R1, *pv = scr + 2 * R1, *lk = scr + 3 * R1, *vo = scr + 4 * R1, *po = scr + 5 * R1;
    if (!d_bridge_w(S, c, bx, by, ex, ey, rows, j0, n, g, scr, R1)) return -1;
    const bool pairs = S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE);   // intended mode (reshape.jl:72,75)
    const double mht = -0.5 * S.tau;
    for (int jp = 1 + lane; jp <= m; jp += 32) {
        const int j = j0 + jp - 1;
        const int p = j <= M ? n : nx, sl = (j <= M ? j : j - M) - 1;
        lk[jp - 1] = mht * (pv[jp - 1] + pv[jp]);                             // lnV (propagator.jl:26-28)
        vo[jp - 1] = S.Vl[VIDX(S, c, p, sl)];
    }
    __syncwarp();
    if (pairs) {
        for (int jp = 1 + lane; jp <= m; jp += 32) {
            const int j = j0 + jp - 1;
            const int p = j <= M ? n : nx, sl = (j <= M ? j : j - M) - 1;
            po[jp - 1] = d_pairs_old(S, c, p, sl);
            pv[jp - 1] = d_pairs_new(S, c, px[jp - 1], py[jp - 1], px[jp], py[jp], sl, p, -1, true);
        }
        __syncwarp();
    }
    int ret = 0;
    if (lane == 0) {                                                          // the sums in the reference order
        double w_initial = 0.0, w_updated = 0.0, sv = 0.0;
        for (int jp = 1; jp <= m; ++jp) {
            w_initial += vo[jp - 1];
            if (pairs) { w_initial += po[jp - 1]; w_updated += pv[jp - 1]; }
            sv = jp == 1 ? lk[0] : sv + lk[jp - 1];
        }
        w_updated += sv;
        ret = d_metropolis(f_exp(w_updated - w_initial), u) ? 1 : 0;
    }
    ret = __shfl_sync(0xffffffffu, ret, 0);
    if (ret) {
        for (int jp = 1 + lane; jp <= m; jp += 32) {                          // rows 1..m and all m links (reshape.jl:82-86)
            const int j = j0 + jp - 1;
            const int p = j <= M ? n : nx, sl = (j <= M ? j : j - M) - 1;
            S.r[RIDX(S, c, p, 0, sl)] = px[jp - 1];
            if (dim > 1) S.r[RIDX(S, c, p, 1, sl)] = py[jp - 1];
            S.Vl[VIDX(S, c, p, sl)] = lk[jp - 1];
            f_cell_update(S, c, sl, p, px[jp - 1], py[jp - 1]);               // one list per slice: lanes touch different lists
        }
        __syncwarp();
    }
    return ret;
}

// ---- sampleparticles weight table (helper.jl:224-267) across the lanes, normalised; sums on lane 0.  Returns n2 (0-based).
__device__ __forceinline__ int d_sample_partner_w(const DevSys &S, int c, int n1, int j0, int m, double u, double *w)
{
    const int lane = threadIdx.x & 31, M = S.M, N = S.N, dim = S.dim;
    const int *nextc = S.next + (size_t)c * N;
    const int jmw = (j0 + m - 1) % M;                                         // mod1(j0 + m, M) - 1
    const bool wrap = j0 + m > M;
    const int n1next = wrap ? nextc[n1] : n1;
    const double mt = m * S.tau;
    const double ax = S.r[RIDX(S, c, n1, 0, j0 - 1)], ay = dim > 1 ? S.r[RIDX(S, c, n1, 1, j0 - 1)] : 0.0;
    const double cx = S.r[RIDX(S, c, n1next, 0, jmw)], cy = dim > 1 ? S.r[RIDX(S, c, n1next, 1, jmw)] : 0.0;
    for (int i = lane; i < N; i += 32) {
        const int inext = wrap ? nextc[i] : i;
        const double t = f_lnK2(ax, ay, S.r[RIDX(S, c, inext, 0, jmw)], dim > 1 ? S.r[RIDX(S, c, inext, 1, jmw)] : 0.0, dim, S.lambda, mt, S.L);
        const double y = f_lnK2(S.r[RIDX(S, c, i, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, i, 1, j0 - 1)] : 0.0, cx, cy, dim, S.lambda, mt, S.L);
        w[i] = f_exp(t + y);
    }
    __syncwarp();
    double norm = 0.0;
    if (lane == 0) { norm = w[0]; for (int i = 1; i < N; ++i) norm = norm + w[i]; }
    norm = __shfl_sync(0xffffffffu, norm, 0);
    for (int i = lane; i < N; i += 32) w[i] = w[i] / norm;
    __syncwarp();
    int n2 = 0;
    if (lane == 0) n2 = d_sample_weighted(w, N, u);
    return __shfl_sync(0xffffffffu, n2, 0);
}

// ---- ReshapeSwapLinear body (reshape.jl:138-279), one warp; scr = 2 * FA_ARR * (M + 1) doubles.
// returns 1 accepted, 0 rejected, -1 bridge failed, -2 n1 == n2 (identical on every lane)
__device__ __forceinline__ int d_reshape_swap_w(const DevSys &S, int c, int n1, int n2, int j0, int m, const GSrc &g1, const GSrc &g2,
                                                double u, double *scr)
{
    if (n1 == n2) return -2;
    const int lane = threadIdx.x & 31, M = S.M, N = S.N, dim = S.dim, R1 = M + 1, jm = j0 + m, rows = m + 1;
    int *nextc = S.next + (size_t)c * N;
    const int x1 = nextc[n1], x2 = nextc[n2];
    const bool wrap = jm > M;
    const int je = (wrap ? jm - M : jm) - 1;
    const int e1 = wrap ? x2 : n2, e2 = wrap ? x1 : n1;                       // bridge 1 ends on the cycle of n2 and vice versa
    double *b1 = scr, *b2 = scr + FA_ARR * R1;
    if (!d_bridge_w(S, c, S.r[RIDX(S, c, n1, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, n1, 1, j0 - 1)] : 0.0,
                    S.r[RIDX(S, c, e1, 0, je)], dim > 1 ? S.r[RIDX(S, c, e1, 1, je)] : 0.0, rows, j0, n2, g1, b1, R1)) return -1;
    if (!d_bridge_w(S, c, S.r[RIDX(S, c, n2, 0, j0 - 1)], dim > 1 ? S.r[RIDX(S, c, n2, 1, j0 - 1)] : 0.0,
                    S.r[RIDX(S, c, e2, 0, je)], dim > 1 ? S.r[RIDX(S, c, e2, 1, je)] : 0.0, rows, j0, n1, g2, b2, R1)) return -1;
    const double mht = -0.5 * S.tau;
    // per bridge: link[jp] = lnV of the new link, gx[jp] = cached old link, then (interactions) gy[jp] = old pair sum, v[jp] = new pair sum
    for (int idx = lane; idx < 2 * m; idx += 32) {
        const int b = idx >= m ? 1 : 0, jp = idx - b * m, j = j0 + jp;
        const int q = j <= M ? (b ? n2 : n1) : (b ? x2 : x1), sl = (j <= M ? j : j - M) - 1;
        double *bb = b ? b2 : b1;
        bb[3 * R1 + jp] = mht * (bb[2 * R1 + jp] + bb[2 * R1 + jp + 1]);
        bb[4 * R1 + jp] = S.Vl[VIDX(S, c, q, sl)];
    }
    __syncwarp();
    if (S.interactions) {
        for (int idx = lane; idx < 2 * m; idx += 32) {
            const int b = idx >= m ? 1 : 0, jp = idx - b * m, j = j0 + jp;
            const int q1 = j <= M ? n1 : x1, q2 = j <= M ? n2 : x2, sl = (j <= M ? j : j - M) - 1;
            double *bb = b ? b2 : b1;
            bb[5 * R1 + jp] = d_pairs_old(S, c, b ? q2 : q1, sl);             // reshape.jl:166-199
            bb[2 * R1 + jp] = d_pairs_new(S, c, bb[jp], bb[R1 + jp], bb[jp + 1], bb[R1 + jp + 1], sl, q1, q2, false);   // :209-240
        }
        __syncwarp();
    }
    int ret = 0;
    if (lane == 0) {                                                          // the sums in the reference order
        double w_initial = 0.0, w_updated = 0.0, s1 = 0.0, s2 = 0.0;
        for (int jp = 0; jp < m; ++jp) {
            w_initial += b1[4 * R1 + jp] + b2[4 * R1 + jp];
            if (S.interactions) { w_initial += b1[5 * R1 + jp]; w_initial += b2[5 * R1 + jp]; }
        }
        for (int jp = 0; jp < m; ++jp) {
            const double v1 = b1[3 * R1 + jp], v2 = b2[3 * R1 + jp];
            s1 = jp == 0 ? v1 : s1 + v1; s2 = jp == 0 ? v2 : s2 + v2;
            if (S.interactions) {
                double add = b1[2 * R1 + jp];
                add += b2[2 * R1 + jp];
                if (S.compat & PIMC_COMPAT_SWAP_SIGN) w_initial += add; else w_updated += add;   // reshape.jl:224,239
            }
        }
        w_updated += s1 + s2;
        ret = d_metropolis(f_exp(w_updated - w_initial), u) ? 1 : 0;
    }
    ret = __shfl_sync(0xffffffffu, ret, 0);
    if (ret) {
        if (lane == 0) { nextc[n1] = x2; nextc[n2] = x1; }                    // reshape.jl:254; the closures then see the re-computed cycles
        __syncwarp();
        for (int jr = 2 + lane; jr <= m + 1; jr += 32) {                      // new rows 2..m+1; one slice (one cell list) per lane
            const int j = j0 + jr - 1;
            const int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
            S.r[RIDX(S, c, q1, 0, sl)] = b1[jr - 1]; if (dim > 1) S.r[RIDX(S, c, q1, 1, sl)] = b1[R1 + jr - 1];
            f_cell_update(S, c, sl, q1, b1[jr - 1], b1[R1 + jr - 1]);
            S.r[RIDX(S, c, q2, 0, sl)] = b2[jr - 1]; if (dim > 1) S.r[RIDX(S, c, q2, 1, sl)] = b2[R1 + jr - 1];
            f_cell_update(S, c, sl, q2, b2[jr - 1], b2[R1 + jr - 1]);
        }
        for (int jp = 1 + lane; jp <= m; jp += 32) {
            const int j = j0 + jp - 1;
            const int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
            S.Vl[VIDX(S, c, q1, sl)] = b1[3 * R1 + jp - 1];
            S.Vl[VIDX(S, c, q2, sl)] = b2[3 * R1 + jp - 1];
        }
        if (jm < M)                                                           // tails jm+1..M change owner (reshape.jl:269-275)
            for (int sl = jm + lane; sl < M; sl += 32) {
                for (int k = 0; k < dim; ++k) {
                    const double t = S.r[RIDX(S, c, n1, k, sl)]; S.r[RIDX(S, c, n1, k, sl)] = S.r[RIDX(S, c, n2, k, sl)]; S.r[RIDX(S, c, n2, k, sl)] = t;
                }
                const double tv = S.Vl[VIDX(S, c, n1, sl)]; S.Vl[VIDX(S, c, n1, sl)] = S.Vl[VIDX(S, c, n2, sl)]; S.Vl[VIDX(S, c, n2, sl)] = tv;
                if (S.need_cells) {
                    f_cell_update(S, c, sl, n1, S.r[RIDX(S, c, n1, 0, sl)], dim > 1 ? S.r[RIDX(S, c, n1, 1, sl)] : 0.0);
                    f_cell_update(S, c, sl, n2, S.r[RIDX(S, c, n2, 0, sl)], dim > 1 ? S.r[RIDX(S, c, n2, 1, sl)] : 0.0);
                }
            }
        __syncwarp();
        if (lane == 0 && !(S.compat & PIMC_COMPAT_SWAP_STALE_LINK) && jm <= M) {   // intended: the link leaving slice j_m changes owner too
            const double tv = S.Vl[VIDX(S, c, n1, jm - 1)]; S.Vl[VIDX(S, c, n1, jm - 1)] = S.Vl[VIDX(S, c, n2, jm - 1)]; S.Vl[VIDX(S, c, n2, jm - 1)] = tv;
        }
        if (S.need_cells) {
            // rm_nn!(old cycles) ... add_nn!(new pol1), add_nn!(new pol2) (reshape.jl:250-251,277-278): a cycle-MERGING swap pushes every
            // member twice into every slice's list (multiplicity 2), any other swap leaves multiplicity 1
            bool merged = false; { int p = nextc[n1], cnt = 0; while (p != n1 && cnt <= N) { if (p == n2) merged = true; p = nextc[p]; cnt++; } }
            const unsigned char mu = merged ? 2 : 1;
            for (int pass = 0; pass < 2; ++pass) {
                const int s0 = pass == 0 ? n1 : n2; int p = s0, cnt = 0;
                do { for (int sl = lane; sl < M; sl += 32) S.mult[VIDX(S, c, p, sl)] = mu; p = nextc[p]; cnt++; } while (p != s0 && cnt <= N);
            }
        }
        __syncwarp();
    }
    return ret;
}

// ---- centre-of-mass move of the permutation cycle of n (com.jl:47-100 / :168-220, move_polymer! helper.jl:368-395) by the whole
// CTA: every thread calls with the same arguments.  The hard-core tests of a displacement (one find_nn per bead of every member,
// helper.jl:385-390) are the expensive part and are spread over all threads; a displacement is redrawn until no bead hits.
// Sums are reduced warp-then-block (parity with the oracle to 1e-12 like d_com_warp; the accept decision and the committed
// positions / links are bit-identical).  red: 34 doubles of shared memory.  returns 1 accepted, 0 rejected, -1 no admissible displacement
__device__ __forceinline__ double block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double s = red[0]; for (int i = 1; i < nw; ++i) s += red[i]; red[32] = s; }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ int d_com_cta(const DevSys &S, int c, int n, double maxd, const DSrc &ds, double u, double *red, int *npol_out,
                                         double *scr, size_t scr_cap)
{
    const int tid = threadIdx.x, nt = blockDim.x, M = S.M, N = S.N, dim = S.dim;
    const int *nextc = S.next + (size_t)c * N;
    const bool pairs = S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE);   // com.jl:54,79,173,198 (intended mode)
    double part = 0.0; int npol = 0;
    { int p = n; do {
            for (int j = tid; j < M; j += nt) {
                part += S.Vl[VIDX(S, c, p, j)];
                if (pairs) part += d_pairs_old(S, c, p, j);
            }
            npol += 1; p = nextc[p]; } while (p != n && npol <= N); }
    const double w_initial = block_sum(part, red);
    // short cycles: the displaced positions and their potentials are staged once (scr: x | y | V, npol * M each) and reused by the
    // hard-core test, the action and the commit; longer cycles recompute them (same expressions, same bits)
    const bool staged = (size_t)3 * npol * M <= scr_cap;
    double *sx = scr, *sy = scr + (size_t)npol * M, *sv = scr + (size_t)2 * npol * M;
    double dx = 0.0, dy = 0.0; bool ok = false;
    for (long long ctr = 1; ctr <= S.ctr; ++ctr) {
        pimc_u4 w = f_draw(ds.st, ds.slot, PIMC_K_COM, (uint32_t)(ctr - 1), 0);
        dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5);
        dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        int hit = 0;
        if (S.a > 0.0 || staged) {
            int p = n, cnt = 0;
            do { for (int j = tid; j < M; j += nt) {
                    double x = f_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? f_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    if (staged) { sx[(size_t)cnt * M + j] = x; sy[(size_t)cnt * M + j] = y; }
                    if (S.a > 0.0 && f_hardcore_hit(S, c, x, y, j, p)) hit = 1;
                }
                p = nextc[p]; cnt++; } while (p != n && cnt <= N);
        }
        if (!__syncthreads_or(hit)) { ok = true; break; }
    }
    int ret = -1;
    if (ok && staged) {
        const double mht = -0.5 * S.tau;
        for (int idx = tid; idx < npol * M; idx += nt) sv[idx] = f_pot(S.pot, sx[idx], sy[idx], dim);
        __syncthreads();
        part = 0.0;
        int p = n, k = 0;
        do { const int kn = k + 1 == npol ? 0 : k + 1;
            for (int j = tid; j < M; j += nt) {
                const size_t a0 = (size_t)k * M + j, a1 = j == M - 1 ? (size_t)kn * M : a0 + 1;
                part += mht * (sv[a0] + sv[a1]);
                if (pairs) part += d_pairs_new(S, c, sx[a0], sy[a0], sx[a1], sy[a1], j, p, -1, false);
            }
            p = nextc[p]; k++; } while (p != n && k <= N);
        const double w_updated = block_sum(part, red);
        ret = d_metropolis(f_exp(w_updated - w_initial), u) ? 1 : 0;             // same value on every thread
        if (ret == 1) {
            p = n; k = 0;
            do { const int kn = k + 1 == npol ? 0 : k + 1;
                for (int j = tid; j < M; j += nt) {                               // a thread owns its slices: per-slice list order as in d_com_warp
                    const size_t a0 = (size_t)k * M + j, a1 = j == M - 1 ? (size_t)kn * M : a0 + 1;
                    S.Vl[VIDX(S, c, p, j)] = mht * (sv[a0] + sv[a1]);
                    S.r[RIDX(S, c, p, 0, j)] = sx[a0]; if (dim > 1) S.r[RIDX(S, c, p, 1, j)] = sy[a0];
                    f_cell_update(S, c, j, p, sx[a0], sy[a0]);
                }
                p = nextc[p]; k++; } while (p != n && k <= N);
            __syncthreads();
        }
    } else if (ok) {
        const double mht = -0.5 * S.tau;
        part = 0.0;
        int p = n, cnt = 0;
        do { int pn = nextc[p];
            for (int j = tid; j < M; j += nt) {
                int q = j == M - 1 ? pn : p, jn = j == M - 1 ? 0 : j + 1;
                double x = f_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? f_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                double xn = f_teleport(S.r[RIDX(S, c, q, 0, jn)] + dx, S.L), yn = dim > 1 ? f_teleport(S.r[RIDX(S, c, q, 1, jn)] + dy, S.L) : 0.0;
                part += mht * (f_pot(S.pot, x, y, dim) + f_pot(S.pot, xn, yn, dim));
                if (pairs) part += d_pairs_new(S, c, x, y, xn, yn, j, p, -1, false);
            }
            p = pn; cnt++; } while (p != n && cnt <= N);
        const double w_updated = block_sum(part, red);
        ret = d_metropolis(f_exp(w_updated - w_initial), u) ? 1 : 0;          // same value on every thread
        if (ret == 1) {
            // link cache first (it reads the still-unshifted neighbours), then the positions
            p = n; cnt = 0;
            do { int pn = nextc[p];
                for (int j = tid; j < M; j += nt) {
                    int q = j == M - 1 ? pn : p, jn = j == M - 1 ? 0 : j + 1;
                    double x = f_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? f_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    double xn = f_teleport(S.r[RIDX(S, c, q, 0, jn)] + dx, S.L), yn = dim > 1 ? f_teleport(S.r[RIDX(S, c, q, 1, jn)] + dy, S.L) : 0.0;
                    S.Vl[VIDX(S, c, p, j)] = mht * (f_pot(S.pot, x, y, dim) + f_pot(S.pot, xn, yn, dim));
                }
                p = pn; cnt++; } while (p != n && cnt <= N);
            __syncthreads();
            p = n; cnt = 0;
            do { for (int j = tid; j < M; j += nt) {                              // a thread owns its slices: per-slice list order as in d_com_warp
                    double x = f_teleport(S.r[RIDX(S, c, p, 0, j)] + dx, S.L), y = dim > 1 ? f_teleport(S.r[RIDX(S, c, p, 1, j)] + dy, S.L) : 0.0;
                    S.r[RIDX(S, c, p, 0, j)] = x; if (dim > 1) S.r[RIDX(S, c, p, 1, j)] = y;
                    f_cell_update(S, c, j, p, x, y);
                }
                p = nextc[p]; cnt++; } while (p != n && cnt <= N);
            __syncthreads();
        }
    }
    if (npol_out) *npol_out = npol;
    return ret;
}
