// pimc_isweep.cuh -- SWEEP schedule for INTERACTING worldlines (hard core a > 0, cell list), optimistic-parallel execution.
//
// Definition (oracle: ORA_SCHED_SWEEP_SEQ): per iteration every worldline of the chain proposes once, strictly in index order, proposal n
// sees the committed results of the proposals before it.  In the reference as shipped (compat flag PAIR_BYVALUE) the pair action
// does not enter ReshapeLinear / centre-of-mass moves, so proposals of one sweep are coupled ONLY through the hard-core tests of
// hardspherelevy! (helper.jl:158-171) and move_polymer! (helper.jl:385-390): `nearest other particle at this slice closer than a`.
// The kernels below execute that sequential definition exactly, but in parallel:
//
//   P1  every proposal is evaluated against the configuration S0 at the START of the sweep (one warp per proposal, lanes = beads): bridge /
//       displacement with the retry-0 draws, hard-core test of every new bead through the cell list, Delta-U, Metropolis.  Nothing is
//       committed: new rows go to the scratch slab `prop / propV`, and are linked into a second cell list NW over the PROPOSED positions.
//       A proposal whose retry-0 draw hits the hard core is marked DIRTY (it needs the serial redraw loop).
//   P3  validation: a proposal is also DIRTY when one of its new beads lies within a(1 + 1e-9) of a new bead of a tentatively accepted
//       proposal of LOWER index at the same slice (found through NW).  For every other proposal the sequential execution would have seen
//       the same hit / no-hit booleans, hence produced the same rows and the same decision: its P1 result is final.
//       (A changed particle can flip a test only by being closer than a before or after its move: old positions closer than a are hits in
//       S0 -> DIRTY by P1; new positions closer than a -> DIRTY by P3.)
//   P4  DIRTY proposals are replayed in index order against the exact sequential state, represented virtually: S0 lists minus the beads
//       moved by finally accepted lower proposals, plus NW restricted to those proposals.  A replay that ends accepted re-inserts its final
//       rows into NW and marks every higher proposal with a bead within a(1 + 1e-9) of them DIRTY as well.
//   P5  commit of all accepted rows (positions, link cache, bins, multiplicity) and deterministic rebuild of the touched slices' lists.
//
// Draws are addressed (include/pimc_rng.h), so P1 and the replays consume exactly the random numbers of the sequential sweep.
// GPU == oracle bit for bit (tests/test_gpu_parity.py::test_interacting_sweep_*).
#pragma once
#include "pimc_sweep_util.cuh"

#define ISW_ACC      1u   // tentatively / finally accepted
#define ISW_DIRTY    2u   // needs a sequential replay
#define ISW_INS      4u   // rows are linked into NW
#define ISW_DONE     8u   // replayed: final
#define ISW_PROP    16u   // this slot proposes in this sweep

// ---- which proposal of this sweep writes bead (o, sl)?  ReshapeLinear: rows t = 0..m_k-1 of proposal k start at slice `first` on particle k
// and wrap onto next[k]; centre-of-mass: every bead of a proposing cycle, proposal index = leader (smallest index) of the cycle.
struct IsCtx {
    int first, M, N;
    const unsigned short *mlen;   // [N] links m_k of proposal k            (reshape)
    const unsigned short *prev;   // [N] inverse permutation                 (reshape)
    const int *lead;              // [N] proposal that moves particle o, -1  (centre of mass)
    unsigned *stat;               // [N] ISW_* bits per proposal
};
template <bool COM> __device__ __forceinline__ int d_is_writer(const IsCtx &X, int o, int sl, int &t)
{
    if (COM) { t = 1; return X.lead[o]; }
    int k;
    if (sl >= X.first) { k = o; t = sl - X.first; } else { k = X.prev[o]; t = sl + X.M - X.first; }
    return t <= (int)X.mlen[k] - 1 ? k : -1;
}

__device__ __forceinline__ void d_nw_insert(const DevSys &S, const ISweepParams &P, int c, int p, int sl, double x, double y)
{
    int *head = P.nw_head + HIDX(S, c, sl, d_bin(S, x, y));
    const int old = atomicExch(head, p);                      // concurrent inserts of other warps; readers only after a CTA barrier
    P.nw_next[NIDX(S, c, sl, p)] = old;
}
__device__ __forceinline__ void d_nw_remove(const DevSys &S, const ISweepParams &P, int c, int p, int sl, double x, double y)   // sequential phases only
{
    int *head = P.nw_head + HIDX(S, c, sl, d_bin(S, x, y));
    int q = *head, pr = -1;
    while (q >= 0 && q != p) { pr = q; q = P.nw_next[NIDX(S, c, sl, q)]; }
    if (q < 0) return;
    const int nx = P.nw_next[NIDX(S, c, sl, p)];
    if (pr < 0) *head = nx; else P.nw_next[NIDX(S, c, sl, pr)] = nx;
}
__device__ __forceinline__ double d_prop_x(const DevSys &S, int c, int p, int sl) { return S.prop[RIDX(S, c, p, 0, sl)]; }
__device__ __forceinline__ double d_prop_y(const DevSys &S, int c, int p, int sl) { return S.dim > 1 ? S.prop[RIDX(S, c, p, 1, sl)] : 0.0; }

// NW entries within a(1 + 1e-9) of (x, y) at slice sl.  MODE 0 (validation of proposal n): is there one written by a tentatively accepted
// proposal k < n?  MODE 1 (after the replay of n): mark every proposal k > n that owns one DIRTY.  Returns the MODE 0 answer.
template <bool COM, int MODE>
static __device__ __noinline__ bool d_nw_near(const DevSys &S, const ISweepParams &P, const IsCtx &X, int c, double x, double y, int sl, int n)
{
    const double ap = S.a * (1.0 + 1e-9), inv = 1.0 / S.cellw;
    const int nb = S.nbins;
    const int ixa = d_floor_div(x - ap + S.L, S.cellw, inv), ixb = d_floor_div(x + ap + S.L, S.cellw, inv);
    int iya = 0, iyb = 0;
    if (S.dim > 1) { iya = d_floor_div(y - ap + S.L, S.cellw, inv); iyb = d_floor_div(y + ap + S.L, S.cellw, inv); }
    bool found = false;
    const int nx = min(ixb - ixa + 1, 3), ny = min(iyb - iya + 1, 3);
    for (int dy = 0; dy < ny; ++dy)
        for (int dx = 0; dx < nx; ++dx) {
            int ix = ixa + dx; ix = ix < 0 ? ix + nb : (ix >= nb ? ix - nb : ix); ix = ix < 0 ? 0 : (ix > nb - 1 ? nb - 1 : ix);
            int iy = iya + dy; iy = iy < 0 ? iy + nb : (iy >= nb ? iy - nb : iy); iy = iy < 0 ? 0 : (iy > nb - 1 ? nb - 1 : iy);
            const int b = S.dim > 1 ? ix + nb * iy : ix;
            for (int o = P.nw_head[HIDX(S, c, sl, b)]; o >= 0; o = P.nw_next[NIDX(S, c, sl, o)]) {
                int t; const int k = d_is_writer<COM>(X, o, sl, t);
                if (k < 0 || k == n) continue;
                if (MODE == 0 ? !(k < n && (X.stat[k] & ISW_ACC)) : !(k > n)) continue;
                if (d_peuclid(S, d_prop_x(S, c, o, sl), d_prop_y(S, c, o, sl), x, y) < ap) {
                    if (MODE == 0) return true;
                    atomicOr(&X.stat[k], ISW_DIRTY);
                }
            }
        }
    return found;
}

// hard-core test of the replay of proposal d at (x, y), slice sl, exception particle exc: find_nn (nearest_neighbours.jl:156-179) + the
// distance test of helper.jl:167-170 on the SEQUENTIAL state = S0 lists minus beads moved by accepted proposals k < d, plus their new beads
template <bool COM>
static __device__ __noinline__ bool d_hit_virtual(const DevSys &S, const ISweepParams &P, const IsCtx &X, int c, double x, double y, int sl, int exc, int d)
{
    const int b = d_bin(S, x, y), nst = S.dim == 2 ? 9 : 3;
    int best = -1; double bd = 0.0, bx = 0.0, by = 0.0;
    for (int q = 0; q < nst; ++q) {
        const int cell = d_stencil(S, b, q);
        for (int o = S.cell_head[HIDX(S, c, sl, cell)]; o >= 0; o = S.cell_next[NIDX(S, c, sl, o)]) {
            if (o == exc) continue;
            int t; const int k = d_is_writer<COM>(X, o, sl, t);
            if (k >= 0 && k < d && t >= 1 && (X.stat[k] & ISW_ACC)) continue;          // moved away by an accepted lower proposal
            const double ox = S.r[RIDX(S, c, o, 0, sl)], oy = S.dim > 1 ? S.r[RIDX(S, c, o, 1, sl)] : 0.0;
            const double pe = d_peuclid(S, ox, oy, x, y);
            if (best < 0 || pe < bd) { best = o; bd = pe; bx = ox; by = oy; }
        }
        for (int o = P.nw_head[HIDX(S, c, sl, cell)]; o >= 0; o = P.nw_next[NIDX(S, c, sl, o)]) {
            if (o == exc) continue;
            int t; const int k = d_is_writer<COM>(X, o, sl, t);
            if (!(k >= 0 && k < d && (X.stat[k] & ISW_ACC))) continue;                  // only final rows of accepted lower proposals exist
            const double ox = d_prop_x(S, c, o, sl), oy = d_prop_y(S, c, o, sl);
            const double pe = d_peuclid(S, ox, oy, x, y);
            if (best < 0 || pe < bd) { best = o; bd = pe; bx = ox; by = oy; }
        }
    }
    if (best < 0) return false;
    const double dx = d_distance(x, bx, S.L), dy = S.dim > 1 ? d_distance(y, by, S.L) : 0.0;
    return d_norm2(dx, dy, S.dim) < S.a;
}

// ---- P5: lists of one slice rebuilt from `bins` by one warp, ascending particle order in every list (== k_cells_build).  The heads of the
// cells occupied BEFORE the commit were cleared by d_clear_slice_heads.
__device__ __forceinline__ void d_clear_slice_heads(const DevSys &S, int c, int sl)
{
    const int lane = threadIdx.x & 31;
    for (int p = lane; p < S.N; p += 32) S.cell_head[HIDX(S, c, sl, S.bins[VIDX(S, c, p, sl)])] = -1;
}
__device__ __forceinline__ void d_rebuild_slice(const DevSys &S, int c, int sl)
{
    const int lane = threadIdx.x & 31, N = S.N;
    for (int base = ((N - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int p = base + lane; const bool valid = p < N;
        const int b = valid ? S.bins[VIDX(S, c, p, sl)] : -1 - lane;
        const unsigned same = __match_any_sync(0xffffffffu, b);
        const unsigned higher = lane == 31 ? 0u : (same & ~((2u << lane) - 1u)), lower = same & ((1u << lane) - 1u);
        int nxt = -1;
        if (valid) {
            if (higher) nxt = base + (__ffs(higher) - 1);
            else nxt = S.cell_head[HIDX(S, c, sl, b)];          // what the chunks above left (or -1)
            S.cell_next[NIDX(S, c, sl, p)] = nxt;
        }
        __syncwarp();
        if (valid && !lower) S.cell_head[HIDX(S, c, sl, b)] = p;
        __syncwarp();
    }
}

// =====================================================================================================================
// ReshapeLinear sweep (reshape.jl:31-91 + hardspherelevy! helper.jl:141-181) of interacting worldlines
// =====================================================================================================================
// rows scratch of one warp: px | py | pv | gx | gy, R1 = M + 2 doubles each

// bridge rows 0..m of proposal n with the retry-0 draws, teleported; returns (warp-uniform) whether any interior bead hits the hard core in S0
__device__ __forceinline__ bool d_isw_rs_eval(const DevSys &S, const ISweepParams &P, const pimc_stream &st, int c, int n, int m, int first,
                                              double *rows, int R1)
{
    const int lane = threadIdx.x & 31, M = S.M, dim = S.dim, nfirst = M - first, mb = m - 1;
    double *px = rows, *py = rows + R1, *gx = rows + 3 * R1, *gy = rows + 4 * R1;
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL;
    const double *rc = S.r + (size_t)c * S.N * dim * M;
    const int nx = S.next[(size_t)c * S.N + n];
    const int pe = m < nfirst ? n : nx, je = m < nfirst ? first + m : m - nfirst;      // reshape.jl:56-58, pcycle helper.jl:113-115
    double bx = rc[(size_t)(n * dim) * M + first], by = dim > 1 ? rc[(size_t)(n * dim + 1) * M + first] : 0.0;
    double ex = rc[(size_t)(pe * dim) * M + je], ey = dim > 1 ? rc[(size_t)(pe * dim + 1) * M + je] : 0.0;
    if (fabs(bx - ex) > L) ex += d_sign(bx) * twoL;                                    // helper.jl:120-125
    if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * twoL;
    for (int j = 1 + lane; j <= mb; j += 32) {
        double g0, g1;
        pimc_gauss_pair_t(pimc_draw_rk(st, &P.sp.rk, (uint32_t)n, PIMC_K_BRIDGE, 0, (uint32_t)j), S.logtab, &g0, &g1);
        const double sig = S.tab_sig[m + 1 - j];
        gx[j] = g0 * sig; gy[j] = dim > 1 ? g1 * sig : 0.0;
    }
    __syncwarp();
    if (lane < dim) {                                                                   // the serial recurrence of levy! (helper.jl:128-135), one lane per dimension
        double *arr = lane ? py : px; const double *g = lane ? gy : gx;
        double prev = lane ? by : bx; const double e = lane ? ey : ex;
        arr[0] = prev; arr[m] = e;
        for (int j = 1; j <= mb; ++j) {
            const double a = S.tab_alpha[m + 1 - j];
            const double t = (1 - a) * e;
            prev = a * prev + t + g[j];
            arr[j] = prev;
        }
    }
    if (dim < 2 && lane == 1) for (int j = 0; j <= m; ++j) py[j] = 0.0;
    __syncwarp();
    for (int t = lane; t <= m; t += 32) {                                               // helper.jl:136-138 (tests see teleported beads, :166)
        px[t] = d_teleport_fast(px[t], L, twoL, inv2L);
        if (dim > 1) py[t] = d_teleport_fast(py[t], L, twoL, inv2L);
    }
    __syncwarp();
    bool hit = false;
    for (int j = 1 + lane; j <= mb; j += 32) {
        const int a = first + j, sl = a >= M ? a - M : a;
        if (d_hardcore_hit(S, c, px[j], py[j], sl, n)) hit = true;
    }
    return __any_sync(0xffffffffu, hit);
}
// Delta-U, Metropolis, rows -> prop / propV (owner particle, slice).  rows: teleported px, py.  Returns acc (warp-uniform).
__device__ __forceinline__ int d_isw_rs_decide(const DevSys &S, const ISweepParams &P, const pimc_stream &st, int c, int n, int m, int first,
                                               double *rows, int R1)
{
    const int lane = threadIdx.x & 31, M = S.M, dim = S.dim, nfirst = M - first;
    double *px = rows, *py = rows + R1, *pv = rows + 2 * R1;
    const double mht = -0.5 * S.tau;
    const int nx = S.next[(size_t)c * S.N + n];
    for (int t = lane; t <= m; t += 32) pv[t] = d_pot(S.pot, px[t], py[t], dim);
    __syncwarp();
    double wi = 0.0, wu = 0.0;
    for (int t = lane; t < m; t += 32) {
        const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
        wi += S.Vl[VIDX(S, c, p, sl)];
        wu += mht * (pv[t] + pv[t + 1]);                                                // lnV (propagator.jl:26-28)
    }
    wi = 0.0 + warp_sum(wi); wu = 0.0 + warp_sum(wu);
    int acc = 0;
    if (lane == 0) {
        const double dw = wu - wi;
        if (dw >= 0.0) acc = 1;
        else {
            const double delta = pimc_exp(dw);
            if (delta >= 1.0) acc = 1;
            else { const pimc_u4 dm = pimc_draw_rk(st, &P.sp.rk, (uint32_t)n, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
        }
    }
    acc = __shfl_sync(0xffffffffu, acc, 0);
    for (int t = lane; t < m; t += 32) {                                                // rows 0..m-1 and links 0..m-1 are what a commit writes (reshape.jl:82-86)
        const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
        S.prop[RIDX(S, c, p, 0, sl)] = px[t];
        if (dim > 1) S.prop[RIDX(S, c, p, 1, sl)] = py[t];
        S.propV[VIDX(S, c, p, sl)] = mht * (pv[t] + pv[t + 1]);
    }
    return acc;
}
__device__ __forceinline__ void d_isw_rs_link(const DevSys &S, const ISweepParams &P, int c, int n, int m, int first, const double *rows, int R1, bool insert)
{
    const int lane = threadIdx.x & 31, M = S.M, nfirst = M - first;
    const int nx = S.next[(size_t)c * S.N + n];
    for (int t = 1 + lane; t < m; t += 32) {                                            // interior rows: the beads that move
        const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
        if (insert) d_nw_insert(S, P, c, p, sl, rows[t], rows[R1 + t]);
        else d_nw_remove(S, P, c, p, sl, d_prop_x(S, c, p, sl), d_prop_y(S, c, p, sl));
    }
}

// replay of proposal d on one warp against the sequential state (hardspherelevy! with its redraw loop, speculate -> test -> repair as in
// pimc_faithful.cuh::d_bridge_w).  Returns 1 accepted, 0 rejected / bridge failed.
static __device__ __noinline__ int d_isw_rs_replay(const DevSys &S, const ISweepParams &P, const IsCtx &X, const pimc_stream &st, int c, int d, int m,
                                                   int first, double *rows, int R1)
{
    const int lane = threadIdx.x & 31, M = S.M, dim = S.dim, nfirst = M - first, mb = m - 1;
    double *px = rows, *py = rows + R1, *pv = rows + 2 * R1, *gx = rows + 3 * R1, *gy = rows + 4 * R1;
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL;
    const double *rc = S.r + (size_t)c * S.N * dim * M;
    const int nx = S.next[(size_t)c * S.N + d];
    const int pe = m < nfirst ? d : nx, je = m < nfirst ? first + m : m - nfirst;
    const double bx = rc[(size_t)(d * dim) * M + first], by = dim > 1 ? rc[(size_t)(d * dim + 1) * M + first] : 0.0;
    double ex = rc[(size_t)(pe * dim) * M + je], ey = dim > 1 ? rc[(size_t)(pe * dim + 1) * M + je] : 0.0;
    if (fabs(bx - ex) > L) ex += d_sign(bx) * twoL;
    if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * twoL;
    for (int j = 1 + lane; j <= mb; j += 32) {
        double g0, g1;
        pimc_gauss_pair_t(pimc_draw_rk(st, &P.sp.rk, (uint32_t)d, PIMC_K_BRIDGE, 0, (uint32_t)j), S.logtab, &g0, &g1);
        const double sig = S.tab_sig[m + 1 - j];
        gx[j] = g0 * sig; gy[j] = dim > 1 ? g1 * sig : 0.0; pv[j] = S.tab_alpha[m + 1 - j];
    }
    __syncwarp();
    double sx0 = bx, sy0 = by;                             // un-teleported row jstart - 1
    int jstart = 1;
    if (lane == 0) { px[0] = d_teleport_fast(bx, L, twoL, inv2L); py[0] = dim > 1 ? d_teleport_fast(by, L, twoL, inv2L) : 0.0;
                     px[m] = d_teleport_fast(ex, L, twoL, inv2L); py[m] = dim > 1 ? d_teleport_fast(ey, L, twoL, inv2L) : 0.0; }
    while (jstart <= mb) {
        {
            double qx = sx0, qy = sy0;
            for (int j = jstart; j <= mb; ++j) {
                const double alpha = pv[j], om = 1 - alpha;
                qx = alpha * qx + om * ex + gx[j];
                if (dim > 1) qy = alpha * qy + om * ey + gy[j];
                if (lane == 0) { px[j] = qx; py[j] = qy; }
            }
        }
        __syncwarp();
        for (int j = jstart + lane; j <= mb; j += 32) {
            px[j] = d_teleport_fast(px[j], L, twoL, inv2L);
            py[j] = dim > 1 ? d_teleport_fast(py[j], L, twoL, inv2L) : 0.0;
        }
        __syncwarp();
        int jf = 0x7fffffff;
        for (int j = jstart + lane; j <= mb; j += 32) {
            const int a = first + j, sl = a >= M ? a - M : a;
            if (d_hit_virtual<false>(S, P, X, c, px[j], py[j], sl, d, d) && j < jf) jf = j;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, jf, o); jf = t < jf ? t : jf; }
        if (jf > mb) break;
        for (int j = jstart; j < jf; ++j) {                // un-teleported row jf - 1
            const double alpha = pv[j], om = 1 - alpha;
            sx0 = alpha * sx0 + om * ex + gx[j];
            if (dim > 1) sy0 = alpha * sy0 + om * ey + gy[j];
        }
        {                                                  // bead jf: its retry-0 draw hits; redraw (helper.jl:160-176), warp-uniform
            const double alpha = pv[jf], om = 1 - alpha;
            const double sig = S.tab_sig[m + 1 - jf];
            const int a = first + jf, sl = a >= M ? a - M : a;
            double nxp = 0.0, nyp = 0.0, tx = 0.0, ty = 0.0; long long ctr = 1; bool pass = true;
            while (pass) {
                pass = false; ctr += 1;
                if (ctr > S.ctr) { pass = true; break; }
                double g0, g1;
                pimc_gauss_pair_t(pimc_draw_rk(st, &P.sp.rk, (uint32_t)d, PIMC_K_BRIDGE, (uint32_t)(ctr - 1), (uint32_t)jf), S.logtab, &g0, &g1);
                nxp = alpha * sx0 + om * ex + g0 * sig;
                if (dim > 1) nyp = alpha * sy0 + om * ey + g1 * sig;
                tx = d_teleport_fast(nxp, L, twoL, inv2L); ty = dim > 1 ? d_teleport_fast(nyp, L, twoL, inv2L) : 0.0;
                if (d_hit_virtual<false>(S, P, X, c, tx, ty, sl, d, d)) pass = true;
            }
            if (pass) return 0;                            // s.ctr draws exhausted: the functor returns false (helper.jl:160-164)
            sx0 = nxp; sy0 = nyp;
            __syncwarp();
            if (lane == 0) { px[jf] = tx; py[jf] = ty; }
        }
        jstart = jf + 1;
        __syncwarp();
    }
    __syncwarp();
    return d_isw_rs_decide(S, P, st, c, d, m, first, rows, R1);
}

#define ISW_TICK(i) do { if (P.prof && threadIdx.x == 0) { const long long t_ = clock64(); pacc[i] += (unsigned long long)(t_ - tlast); tlast = t_; } } while (0)

__device__ __forceinline__ void d_isw_commit_rebuild(const DevSys &S, int c, int sl0, int nsl)
{
    // lists of slices sl0 .. sl0 + nsl - 1 (mod M): one warp per slice
    const int warp = threadIdx.x >> 5, NW = blockDim.x >> 5, M = S.M;
    for (int i = warp; i < nsl; i += NW) { int sl = sl0 + i; if (sl >= M) sl -= M; d_rebuild_slice(S, c, sl); }
}

__global__ void __launch_bounds__(ISW_THREADS, 4) k_isweep_reshape(const __grid_constant__ DevSys S, const __grid_constant__ ISweepParams P)
{
    extern __shared__ double smd[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = ISW_THREADS / 32;
    const int M = S.M, N = S.N, dim = S.dim;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.sp.iter);
    const pimc_u4 di = pimc_draw_rk(st, &P.sp.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P.sp, di);
    if (P.sp.kind[pick] != PIMC_UPD_RESHAPE_LINEAR) return;
    const UpdDev &U = P.upd[pick];
    unsigned char *flag = (unsigned char *)smd;
    unsigned *stat = (unsigned *)(flag + ((N + 15) & ~15));
    unsigned short *mlen = (unsigned short *)(stat + N), *prev = mlen + N;
    double *rows = (double *)(((size_t)(prev + N) + 4 * (size_t)N + 15) & ~(size_t)15) + (size_t)warp * ISW_RARR * (M + 2);
    const int R1 = M + 2;
    __shared__ BookPre s_pre; __shared__ unsigned long long s_bead; __shared__ int s_maxm, s_nrep;
    unsigned long long pacc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; long long tlast = P.prof ? clock64() : 0;
    if (tid == 0) { s_pre = d_book_prefetch(U, c); s_bead = 0; s_maxm = 0; s_nrep = 0; }
    __syncthreads();                                                                    // the counters are accumulated with atomics right below
    const int var = (int)U.var[c], vmax = (int)P.sp.vmax[pick];
    const int first = (int)pimc_index(di.w[2], (uint32_t)M);                            // j0 - 1
    const int *nextc = S.next + (size_t)c * N;
    unsigned long long my_beads = 0; int my_maxm = 0;
    for (int n = tid; n < N; n += ISW_THREADS) {
        const pimc_u4 dt = pimc_draw_rk(st, &P.sp.rk, (uint32_t)n, PIMC_K_TASK, 0, 0);
        const int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
        const int m = vmax < mm ? vmax : mm;
        mlen[n] = (unsigned short)m; prev[nextc[n]] = (unsigned short)n; stat[n] = ISW_PROP; flag[n] = 0;
        my_beads += (unsigned long long)(m - 1); my_maxm = m > my_maxm ? m : my_maxm;
    }
    my_maxm = __reduce_max_sync(0xffffffffu, my_maxm);
    { const unsigned ws = __reduce_add_sync(0xffffffffu, (unsigned)my_beads); if (lane == 0) { if (ws) atomicAdd(&s_bead, (unsigned long long)ws); atomicMax(&s_maxm, my_maxm); } }
    __syncthreads();
    IsCtx X; X.first = first; X.M = M; X.N = N; X.mlen = mlen; X.prev = prev; X.lead = nullptr; X.stat = stat;
    ISW_TICK(0);
    // ---- P1: every proposal against S0 ----
    for (int n = warp; n < N; n += NW) {
        const int m = mlen[n];
        const bool hit = d_isw_rs_eval(S, P, st, c, n, m, first, rows, R1);
        unsigned s = ISW_PROP;
        if (hit) s |= ISW_DIRTY;
        else {
            const int acc = d_isw_rs_decide(S, P, st, c, n, m, first, rows, R1);
            d_isw_rs_link(S, P, c, n, m, first, rows, R1, true);
            s |= ISW_INS | (acc ? ISW_ACC : 0u);
        }
        if (lane == 0) stat[n] = s;
        __syncwarp();
    }
    __threadfence_block();
    __syncthreads();
    ISW_TICK(1);
    // ---- P3: validation against the new beads of lower, tentatively accepted proposals ----
    for (int n = warp; n < N; n += NW) {
        if (stat[n] & ISW_DIRTY) continue;
        const int m = mlen[n], nx = nextc[n], nfirst = M - first;
        bool bad = false;
        for (int t = 1 + lane; t < m; t += 32) {
            const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
            if (d_nw_near<false, 0>(S, P, X, c, d_prop_x(S, c, p, sl), d_prop_y(S, c, p, sl), sl, n)) bad = true;
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&stat[n], ISW_DIRTY);
    }
    ISW_TICK(2);
    // ---- P4: dirty proposals in index order against the sequential state, all on warp 0 (no CTA barrier per replay; the scan of `stat`
    //      and the replay must not race: a replay sets DONE on its own slot and DIRTY on higher ones) ----
    __syncthreads();
    if (warp == 0) {
        for (int d = 0; d < N; ++d) {
            if ((stat[d] & (ISW_DIRTY | ISW_DONE)) != ISW_DIRTY) continue;                // warp-uniform (one shared-memory word)
            const int m = mlen[d];
            if (stat[d] & ISW_INS) d_isw_rs_link(S, P, c, d, m, first, rows, R1, false);  // its tentative rows leave NW
            __syncwarp();
            const int acc = d_isw_rs_replay(S, P, X, st, c, d, m, first, rows, R1);
            unsigned s = ISW_PROP | ISW_DIRTY | ISW_DONE;
            if (acc) {
                d_isw_rs_link(S, P, c, d, m, first, rows, R1, true);
                __syncwarp();
                for (int t = 1 + lane; t < m; t += 32) {                                  // higher proposals with a bead next to the final rows
                    const int a = first + t, sl = a >= M ? a - M : a;
                    d_nw_near<false, 1>(S, P, X, c, rows[t], rows[R1 + t], sl, d);
                }
                s |= ISW_ACC | ISW_INS;
            }
            __syncwarp();
            if (lane == 0) { stat[d] = s; s_nrep += 1; }
            __syncwarp();
        }
    }
    __threadfence_block();
    __syncthreads();
    ISW_TICK(3);
    // ---- P5: commit ----
    const int nsl = s_maxm;                                                               // rows t = 0 .. maxm - 1 can change
    for (int i = warp; i < nsl; i += NW) { int sl = first + i; if (sl >= M) sl -= M; d_clear_slice_heads(S, c, sl); }
    __syncthreads();
    for (int n = warp; n < N; n += NW) {
        const unsigned s = stat[n];
        const int m = mlen[n], nx = nextc[n], nfirst = M - first;
        if (s & ISW_INS)
            for (int t = 1 + lane; t < m; t += 32) {                                      // NW back to empty
                const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
                P.nw_head[HIDX(S, c, sl, d_bin(S, d_prop_x(S, c, p, sl), d_prop_y(S, c, p, sl)))] = -1;
            }
        if (s & ISW_ACC) {
            for (int t = lane; t < m; t += 32) {                                          // reshape.jl:82-86: rows 1..m, update_nn_bead! each
                const int p = t < nfirst ? n : nx, sl = t < nfirst ? first + t : t - nfirst;
                const double x = d_prop_x(S, c, p, sl), y = d_prop_y(S, c, p, sl);
                S.r[RIDX(S, c, p, 0, sl)] = x; if (dim > 1) S.r[RIDX(S, c, p, 1, sl)] = y;
                S.Vl[VIDX(S, c, p, sl)] = S.propV[VIDX(S, c, p, sl)];
                S.bins[VIDX(S, c, p, sl)] = d_bin(S, x, y);
                S.mult[VIDX(S, c, p, sl)] = 1;
            }
            if (lane == 0) flag[n] = 1;
        }
    }
    __threadfence_block();
    __syncthreads();
    d_isw_commit_rebuild(S, c, first, nsl);
    ISW_TICK(4);
    __syncthreads();
    if (warp == 0) d_bookkeep_sweep_warp(U, c, flag, N, s_bead, P.sp.stats, s_pre);
    ISW_TICK(5);
    if (P.prof && tid == 0) for (int i = 0; i < 6; ++i) atomicAdd(P.prof + i, pacc[i]);
    if (tid == 0) { atomicAdd(P.sp.stats + 12, (unsigned long long)s_nrep); atomicAdd(P.sp.stats + 13, (unsigned long long)N); }   // replay rate: the host's dispatch heuristic
}

// =====================================================================================================================
// centre-of-mass sweeps (com.jl:31-104,136-224; move_polymer! helper.jl:368-395) of interacting worldlines
// =====================================================================================================================
// whole-CTA replay of proposal d (cycle led by particle d) against the sequential state; same draws / retry rule as d_com_cta.
// Writes the final rows of an accepted move to prop / propV.  Returns 1 accepted, 0 rejected, -1 no admissible displacement (every thread).
static __device__ __noinline__ int d_isw_com_replay(const DevSys &S, const ISweepParams &P, const IsCtx &X, const pimc_stream &st, int c, int d, double maxd,
                                                    double *red, int *npol_out)
{
    const int tid = threadIdx.x, nt = blockDim.x, M = S.M, N = S.N, dim = S.dim;
    const int *nextc = S.next + (size_t)c * N;
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    double part = 0.0; int npol = 0;
    { int p = d; do { for (int j = tid; j < M; j += nt) part += S.Vl[VIDX(S, c, p, j)]; npol += 1; p = nextc[p]; } while (p != d && npol <= N); }
    // block_sum (pimc_faithful.cuh) restated: warp sums, then warp 0 order
    auto bsum = [&](double v) {
        const int lane = tid & 31, w = tid >> 5, nw = (nt + 31) >> 5;
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) red[w] = v;
        __syncthreads();
        if (tid == 0) { double s = red[0]; for (int i = 1; i < nw; ++i) s += red[i]; red[32] = s; }
        __syncthreads();
        return red[32];
    };
    const double w_initial = bsum(part);
    double dx = 0.0, dy = 0.0; bool ok = false;
    for (long long ctr = 1; ctr <= S.ctr; ++ctr) {
        const pimc_u4 w = pimc_draw_rk(st, &P.sp.rk, (uint32_t)d, PIMC_K_COM, (uint32_t)(ctr - 1), 0);
        dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5);
        dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        int hit = 0;
        int p = d, cnt = 0;
        do { for (int j = tid; j < M; j += nt) {
                const double x = d_teleport_fast(S.r[RIDX(S, c, p, 0, j)] + dx, L, twoL, inv2L), y = dim > 1 ? d_teleport_fast(S.r[RIDX(S, c, p, 1, j)] + dy, L, twoL, inv2L) : 0.0;
                if (d_hit_virtual<true>(S, P, X, c, x, y, j, p, d)) hit = 1;
            }
            p = nextc[p]; cnt++; } while (p != d && cnt <= N);
        if (!__syncthreads_or(hit)) { ok = true; break; }
    }
    int ret = -1;
    if (ok) {
        part = 0.0;
        int p = d, cnt = 0;
        do { const int pn = nextc[p];
            for (int j = tid; j < M; j += nt) {
                const int q = j == M - 1 ? pn : p, jn = j == M - 1 ? 0 : j + 1;
                const double x = d_teleport_fast(S.r[RIDX(S, c, p, 0, j)] + dx, L, twoL, inv2L), y = dim > 1 ? d_teleport_fast(S.r[RIDX(S, c, p, 1, j)] + dy, L, twoL, inv2L) : 0.0;
                const double xn = d_teleport_fast(S.r[RIDX(S, c, q, 0, jn)] + dx, L, twoL, inv2L), yn = dim > 1 ? d_teleport_fast(S.r[RIDX(S, c, q, 1, jn)] + dy, L, twoL, inv2L) : 0.0;
                const double lk = mht * (d_pot(S.pot, x, y, dim) + d_pot(S.pot, xn, yn, dim));
                part += lk;
                S.prop[RIDX(S, c, p, 0, j)] = x; if (dim > 1) S.prop[RIDX(S, c, p, 1, j)] = y;
                S.propV[VIDX(S, c, p, j)] = lk;
            }
            p = pn; cnt++; } while (p != d && cnt <= N);
        const double w_updated = bsum(part);
        const pimc_u4 dm = pimc_draw_rk(st, &P.sp.rk, (uint32_t)d, PIMC_K_TASK, 0, 1);
        ret = d_metropolis(pimc_exp(w_updated - w_initial), pimc_u01_co(dm.w[0], dm.w[1])) ? 1 : 0;
    }
    if (npol_out) *npol_out = npol;
    return ret;
}

template <int KM>
__global__ void __launch_bounds__(ISW_THREADS, 4) k_isweep_com(const __grid_constant__ DevSys S, const __grid_constant__ ISweepParams P)
{
    extern __shared__ double smd[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = ISW_THREADS / 32;
    const int M = S.M, N = S.N, dim = S.dim;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.sp.iter);
    const pimc_u4 di = pimc_draw_rk(st, &P.sp.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P.sp, di);
    const int kind = P.sp.kind[pick];
    if (kind != PIMC_UPD_SINGLE_COM && kind != PIMC_UPD_POLYMER_COM) return;
    const bool polymer = kind == PIMC_UPD_POLYMER_COM;
    const UpdDev &U = P.upd[pick];
    unsigned char *flag = (unsigned char *)smd;
    unsigned *stat = (unsigned *)(flag + ((N + 15) & ~15));
    unsigned short *mlen = (unsigned short *)(stat + N), *prev = mlen + N;
    int *lead = (int *)(prev + N);
    double *red = (double *)(((size_t)(lead + N) + 15) & ~(size_t)15);
    __shared__ BookPre s_pre; __shared__ unsigned long long s_bead; __shared__ int s_nrep, s_anyacc;
    unsigned long long pacc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; long long tlast = P.prof ? clock64() : 0;
    if (tid == 0) { s_pre = d_book_prefetch(U, c); s_bead = 0; s_nrep = 0; s_anyacc = 0; }
    const double maxd = U.var[c];
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int *nextc = S.next + (size_t)c * N;
    // proposals: SingleCOM -- every particle with next == self (com.jl:139-164); PolymerCOM -- every cycle once, slot = its smallest index
    for (int n = tid; n < N; n += ISW_THREADS) {
        int ld = n, p = nextc[n], cnt = 0;
        while (p != n && cnt <= N) { ld = p < ld ? p : ld; p = nextc[p]; cnt++; }
        const bool single = nextc[n] == n;
        lead[n] = (single || polymer) ? ld : -1;
        const bool proposes = single || (polymer && ld == n);
        stat[n] = proposes ? (ISW_PROP | (single ? 0u : ISW_DIRTY)) : 0u;              // exchange cycles take the sequential path
        flag[n] = proposes ? 0 : 2;
    }
    __syncthreads();
    IsCtx X; X.first = 0; X.M = M; X.N = N; X.mlen = mlen; X.prev = prev; X.lead = lead; X.stat = stat;
    ISW_TICK(0);
    unsigned long long my_beads = 0;
    // ---- P1: single worldlines against S0, one warp per proposal, the worldline in registers ----
    for (int n = warp; n < N; n += NW) {
        if ((stat[n] & (ISW_PROP | ISW_DIRTY)) != ISW_PROP) continue;
        const pimc_u4 w = pimc_draw_rk(st, &P.sp.rk, (uint32_t)n, PIMC_K_COM, 0, 0);
        const double dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5), dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        const double *rx = S.r + RIDX(S, c, n, 0, 0), *ry = rx + M, *vl = S.Vl + VIDX(S, c, n, 0);
        double x[KM], y[KM], v[KM], wi = 0.0, wu = 0.0;
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            x[k] = j < M ? rx[j] : 0.0; y[k] = (dim > 1 && j < M) ? ry[j] : 0.0; v[k] = j < M ? vl[j] : 0.0;
        }
        bool hit = false;
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            wi += j < M ? v[k] : 0.0;
            x[k] = d_teleport_fast(x[k] + dx, L, twoL, inv2L);
            if (dim > 1) y[k] = d_teleport_fast(y[k] + dy, L, twoL, inv2L);
        }
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            if (j < M && d_hardcore_hit(S, c, x[k], y[k], j, n)) hit = true;
        }
        if (lane == 0) my_beads += (unsigned long long)M;
        if (__any_sync(0xffffffffu, hit)) { if (lane == 0) stat[n] = ISW_PROP | ISW_DIRTY; __syncwarp(); continue; }
#pragma unroll
        for (int k = 0; k < KM; ++k) v[k] = (lane + 32 * k < M) ? d_pot(S.pot, x[k], y[k], dim) : 0.0;
        const double v00 = __shfl_sync(0xffffffffu, v[0], 0);
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            double up = __shfl_down_sync(0xffffffffu, v[k], 1);
            const double nextreg = __shfl_sync(0xffffffffu, v[(k + 1 < KM) ? k + 1 : k], 0);
            if (lane == 31) up = nextreg;
            if (j == M - 1) up = v00;
            const double lk = mht * (v[k] + up);
            v[k] = lk;
            if (j < M) wu += lk;
        }
        wi = 0.0 + warp_sum(wi); wu = 0.0 + warp_sum(wu);
        int acc = 0;
        if (lane == 0) {
            const double dw = wu - wi;
            if (dw >= 0.0) acc = 1;
            else {
                const double delta = pimc_exp(dw);
                if (delta >= 1.0) acc = 1;
                else { const pimc_u4 dm = pimc_draw_rk(st, &P.sp.rk, (uint32_t)n, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
            }
            stat[n] = ISW_PROP | ISW_INS | (acc ? ISW_ACC : 0u);
        }
        double *qx = S.prop + RIDX(S, c, n, 0, 0), *qy = qx + M, *qv = S.propV + VIDX(S, c, n, 0);
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            if (j < M) { qx[j] = x[k]; if (dim > 1) qy[j] = y[k]; qv[j] = v[k]; d_nw_insert(S, P, c, n, j, x[k], y[k]); }
        }
        __syncwarp();
    }
    __threadfence_block();
    __syncthreads();
    ISW_TICK(1);
    // ---- P3 ----
    for (int n = warp; n < N; n += NW) {
        if ((stat[n] & (ISW_PROP | ISW_DIRTY)) != ISW_PROP) continue;
        bool bad = false;
        for (int j = lane; j < M; j += 32)
            if (d_nw_near<true, 0>(S, P, X, c, d_prop_x(S, c, n, j), d_prop_y(S, c, n, j), j, n)) bad = true;
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&stat[n], ISW_DIRTY);
    }
    ISW_TICK(2);
    // ---- P4: the whole CTA on one dirty proposal at a time, in index order ----
    for (int cur = 0;;) {
        __syncthreads();
        int d = cur;
        while (d < N && (stat[d] & (ISW_PROP | ISW_DIRTY | ISW_DONE)) != (ISW_PROP | ISW_DIRTY)) ++d;
        if (d >= N) break;
        if (stat[d] & ISW_INS) {                                                          // (only single worldlines were inserted)
            for (int j = tid; j < M; j += ISW_THREADS) d_nw_remove(S, P, c, d, j, d_prop_x(S, c, d, j), d_prop_y(S, c, d, j));
            __threadfence_block();
            __syncthreads();
        }
        int npol = 1;
        const int r = d_isw_com_replay(S, P, X, st, c, d, maxd, red, &npol);
        __threadfence_block();
        __syncthreads();
        if (r == 1) {
            int p = d, cnt = 0;
            do { for (int j = tid; j < M; j += ISW_THREADS) d_nw_insert(S, P, c, p, j, d_prop_x(S, c, p, j), d_prop_y(S, c, p, j));
                 p = nextc[p]; cnt++; } while (p != d && cnt <= N);
            __threadfence_block();
            __syncthreads();
            p = d; cnt = 0;
            do { for (int j = tid; j < M; j += ISW_THREADS) d_nw_near<true, 1>(S, P, X, c, d_prop_x(S, c, p, j), d_prop_y(S, c, p, j), j, d);
                 p = nextc[p]; cnt++; } while (p != d && cnt <= N);
        }
        __syncthreads();
        if (tid == 0) {
            stat[d] = ISW_PROP | ISW_DIRTY | ISW_DONE | (r == 1 ? (ISW_ACC | ISW_INS) : 0u);
            s_nrep += 1;
            if (!(nextc[d] == d)) my_beads += (unsigned long long)M * npol;               // cycles were not counted in P1
        }
        cur = d + 1;
    }
    ISW_TICK(3);
    // ---- P5 ----
    int any = 0;
    for (int n = tid; n < N; n += ISW_THREADS) any |= (stat[n] & ISW_ACC) != 0;
    any = __syncthreads_or(any);
    if (any) for (int sl = warp; sl < M; sl += NW) d_clear_slice_heads(S, c, sl);
    __syncthreads();
    for (int p = warp; p < N; p += NW) {
        const int k = lead[p];
        if (k < 0) continue;
        const unsigned s = stat[k];
        if (!(s & ISW_INS)) continue;
        const bool acc = (s & ISW_ACC) != 0;
        for (int j = lane; j < M; j += 32) {
            const double x = d_prop_x(S, c, p, j), y = d_prop_y(S, c, p, j);
            const int b = d_bin(S, x, y);
            P.nw_head[HIDX(S, c, j, b)] = -1;
            if (acc) {
                S.r[RIDX(S, c, p, 0, j)] = x; if (dim > 1) S.r[RIDX(S, c, p, 1, j)] = y;
                S.Vl[VIDX(S, c, p, j)] = S.propV[VIDX(S, c, p, j)];
                S.bins[VIDX(S, c, p, j)] = b; S.mult[VIDX(S, c, p, j)] = 1;
            }
        }
        if (acc && lane == 0 && k == p) flag[k] = 1;
    }
    __threadfence_block();
    __syncthreads();
    if (any) d_isw_commit_rebuild(S, c, 0, M);
    ISW_TICK(4);
    { const unsigned long long b = my_beads; if (b) atomicAdd(&s_bead, b); }
    __syncthreads();
    if (warp == 0) d_bookkeep_sweep_warp(U, c, flag, N, s_bead, P.sp.stats, s_pre);
    ISW_TICK(5);
    if (P.prof && tid == 0) for (int i = 0; i < 6; ++i) atomicAdd(P.prof + 10 + i, pacc[i]);
    if (tid == 0) { atomicAdd(P.sp.stats + 12, (unsigned long long)s_nrep); atomicAdd(P.sp.stats + 13, (unsigned long long)N); }
}
