/*
 * pimc_rng.h -- the random-number SPEC of the pimc-b200 engine.
 *
 * The reference (oameye/PIMC.jl) draws from Julia's global RNG
 * (`rand`, `randn`, `StatsBase.sample`: src/updates/helper.jl:4,134,165,207,225,378;
 * src/updates/reshape.jl:37-38,45,129-130; src/updates/com.jl:42,163;
 * src/simulation.jl:34-36) and never seeds it, so no stream of the reference can be
 * reproduced.  This header therefore DEFINES the stream: counter-based Philox4x32-10
 * (Salmon et al., SC'11) addressed by (seed, chain, iteration, slot, kind, retry, bead),
 * and a bit-reproducible uniform -> Gaussian / exp map built from IEEE-754 basic operations
 * (+ - * / sqrt, explicit fma) only.  The same header is compiled by nvcc for the sm_100a
 * kernels and by gcc for the CPU oracle, so a CPU trajectory and a GPU trajectory can be
 * compared bit for bit.  Nothing in here restates reference arithmetic; it only replaces
 * Julia's RNG, which has no reproducible behaviour to be faithful to.
 *
 * Build flags that make it bit-reproducible: nvcc -fmad=false ; gcc -ffp-contract=off -mfma.
 */
#ifndef PIMC_RNG_H
#define PIMC_RNG_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PIMC_HD __host__ __device__ __forceinline__
#else
#define PIMC_HD static inline
#endif

/* ---- polynomial coefficients of the Gaussian transform.  On the device they live in the constant bank and reach the fp64
 * pipe as c[bank][offset] operands: as literals every coefficient costs two extra instructions (an fp64 immediate is moved
 * into a register pair half by half) in front of each DFMA of the hot loop.  Host and device read the same values. */
#define PIMC_KLIST(X) \
    X(PIMC_KI_LN2HI, 6.93147180369123816490e-01) X(PIMC_KI_LN2LO, 1.90821492927058770002e-10) \
    X(PIMC_KI_L7, 1.0 / 7.0) X(PIMC_KI_L6, -1.0 / 6.0) X(PIMC_KI_L5, 0.2) X(PIMC_KI_L3, 1.0 / 3.0) \
    X(PIMC_KI_TWOPI, 6.283185307179586476925) \
    X(PIMC_KI_S0, 1.58969099521155010221e-10) X(PIMC_KI_S1, -2.50507602534068634195e-08) X(PIMC_KI_S2, 2.75573137070700676789e-06) \
    X(PIMC_KI_S3, -1.98412698298579493134e-04) X(PIMC_KI_S4, 8.33333333332248946124e-03) X(PIMC_KI_S5, -1.66666666666666324348e-01) \
    X(PIMC_KI_C0, -1.13596475577881948265e-11) X(PIMC_KI_C1, 2.08757232129817482790e-09) X(PIMC_KI_C2, -2.75573143513906633035e-07) \
    X(PIMC_KI_C3, 2.48015872894767294178e-05) X(PIMC_KI_C4, -1.38888888888741095749e-03) X(PIMC_KI_C5, 4.16666666666666019037e-02)
#define PIMC_KENUM(name, value) name,
enum { PIMC_KLIST(PIMC_KENUM) PIMC_KI_COUNT };
#undef PIMC_KENUM
#if defined(__CUDACC__)
#define PIMC_KVAL(name, value) value,
static __constant__ double pimc_kc_dev[PIMC_KI_COUNT] = { PIMC_KLIST(PIMC_KVAL) };
#undef PIMC_KVAL
#endif
#define PIMC_KHOST(name, value) value,
static const double pimc_kc_host[PIMC_KI_COUNT] = { PIMC_KLIST(PIMC_KHOST) };
#undef PIMC_KHOST
#if defined(__CUDA_ARCH__)
#define PIMC_K(name) pimc_kc_dev[name]
#else
#define PIMC_K(name) pimc_kc_host[name]
#endif

/* ---- draw kinds (bits 28..31 of counter word 0) ------------------------------------ */
#define PIMC_K_ITER     0u  /* per-iteration, per-chain: update pick, sweep window j0    */
#define PIMC_K_TASK     1u  /* per-move choices: n, j0, m ; bead=1: Metropolis uniform   */
#define PIMC_K_BRIDGE   2u  /* Gaussian pair of bridge bead t (retry = hard-core redraw) */
#define PIMC_K_BRIDGE2  3u  /* second bridge of the swap move                            */
#define PIMC_K_COM      4u  /* centre-of-mass displacement (retry = hard-core redraw)    */
#define PIMC_K_SWAP     5u  /* swap move: n1, n2 table draw                              */
#define PIMC_K_INIT     6u  /* init_world: ring Gaussians (retry = levy! call index)     */
#define PIMC_K_INIT0    7u  /* init_world: uniform start point (retry = attempt index)   */

#define PIMC_SLOT_CHAIN 0xFFFFu /* slot id of chain-level draws */

typedef struct { uint32_t w[4]; } pimc_u4;

typedef struct {
    uint32_t seed_lo, seed_hi; /* Philox key                                  */
    uint32_t chain;            /* global chain id  -> counter word 3          */
    uint32_t iter_lo;          /* iteration index  -> counter word 2          */
    uint32_t iter_hi16;        /* high 16 bits of the 48-bit iteration index  */
} pimc_stream;

PIMC_HD void pimc_mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo)
{
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

/* Philox4x32-10, Random123 round function and key schedule. */
PIMC_HD pimc_u4 pimc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                   uint32_t k0, uint32_t k1)
{
    pimc_u4 o;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        pimc_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        pimc_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

/* The 10 round keys depend on the seed only: precomputed once, they reach the kernels as constant-bank operands. */
typedef struct { uint32_t k0[10], k1[10]; } pimc_roundkeys;
PIMC_HD void pimc_roundkeys_make(uint64_t seed, pimc_roundkeys *rk)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk->k0[r] = k0; rk->k1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
PIMC_HD pimc_u4 pimc_philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const pimc_roundkeys *rk)
{
    pimc_u4 o;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        pimc_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        pimc_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ rk->k0[r];
        uint32_t n2 = hi0 ^ c3 ^ rk->k1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

PIMC_HD pimc_stream pimc_stream_make(uint64_t seed, uint32_t chain, uint64_t iter)
{
    pimc_stream s;
    s.seed_lo = (uint32_t)seed;
    s.seed_hi = (uint32_t)(seed >> 32);
    s.chain = chain;
    s.iter_lo = (uint32_t)iter;
    s.iter_hi16 = (uint32_t)((iter >> 32) & 0xFFFFu);
    return s;
}

/* One addressed 128-bit draw.  slot < 2^16, kind < 16, retry < 2^14, bead < 2^14. */
PIMC_HD pimc_u4 pimc_draw(pimc_stream s, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead)
{
    uint32_t c0 = (kind << 28) | ((retry & 0x3FFFu) << 14) | (bead & 0x3FFFu);
    uint32_t c1 = (slot & 0xFFFFu) | (s.iter_hi16 << 16);
    return pimc_philox4x32_10(c0, c1, s.iter_lo, s.chain, s.seed_lo, s.seed_hi);
}

PIMC_HD pimc_u4 pimc_draw_rk(pimc_stream s, const pimc_roundkeys *rk, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead)
{
    uint32_t c0 = (kind << 28) | ((retry & 0x3FFFu) << 14) | (bead & 0x3FFFu);
    uint32_t c1 = (slot & 0xFFFFu) | (s.iter_hi16 << 16);
    return pimc_philox4x32_10_rk(c0, c1, s.iter_lo, s.chain, rk);
}

/* words (0,1) -> uniform in (0,1] ; words (2,3) -> uniform in [0,1) ; both 53-bit. */
PIMC_HD double pimc_u01_oc(uint32_t lo, uint32_t hi)
{
    uint64_t x = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return ((double)x + 1.0) * 0x1p-53;
}
PIMC_HD double pimc_u01_co(uint32_t lo, uint32_t hi)
{
    uint64_t x = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return (double)x * 0x1p-53;
}
/* integer in [0, n) from one 32-bit word (multiply-shift). */
PIMC_HD uint32_t pimc_index(uint32_t w, uint32_t n)
{
    return (uint32_t)(((uint64_t)w * (uint64_t)n) >> 32);
}

PIMC_HD double pimc_bits2d(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { uint64_t u; double d; } c; c.u = b; return c.d;
#endif
}
PIMC_HD uint64_t pimc_d2bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { uint64_t u; double d; } c; c.d = d; return c.u;
#endif
}

/* natural log for normal positive x; atanh series with one IEEE division, < 1 ulp typ.  Reference implementation: it
 * generates the table of pimc_log_tab and serves as its accuracy yardstick. */
PIMC_HD double pimc_log(double x)
{
    uint64_t b = pimc_d2bits(x);
    int k = (int)(b >> 52) - 1023;
    uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
    /* m in [sqrt(1/2), sqrt(2)) : mantissa above sqrt(2) goes down one binade */
    if (mant > 0x6A09E667F3BCCull) { k += 1; b = mant | 0x3FE0000000000000ull; }
    else                           {          b = mant | 0x3FF0000000000000ull; }
    double m = pimc_bits2d(b);
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double R = 1.479819860511658591e-01;
    R = fma(R, z, 1.531383769920937332e-01);
    R = fma(R, z, 1.818357216161805012e-01);
    R = fma(R, z, 2.222219843214978396e-01);
    R = fma(R, z, 2.857142874366239149e-01);
    R = fma(R, z, 3.999999999940941908e-01);
    R = fma(R, z, 6.666666666666735130e-01);
    R = R * z;
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    return dk * 6.93147180369123816490e-01 -
           ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

/* Division-free log for the Gaussian transform: x = 2^k z, z in [1,2); the top 7 mantissa bits pick c_i (interval midpoint;
 * 1 and 2 exactly for the first and last interval so that results near x = 1 keep their relative accuracy);
 * e = z/c_i - 1 through the tabulated 1/c_i (a 10-bit number, so the fma is exact; |e| < 2^-7), log(1+e) by a degree-7 Taylor polynomial, and
 * log(c_i) [minus ln 2 and k+1 when c_i >= sqrt 2, which keeps the pieces small near x = 1] from the table.
 * tab[2i] = 1/c_i, tab[2i+1] = log(c_i) adjusted; filled by pimc_logtab_fill with IEEE operations only. */
#define PIMC_LOGTAB_N 128
PIMC_HD double pimc_logtab_center(int i)
{
    if (i == 0) return 1.0;
    if (i == PIMC_LOGTAB_N - 1) return 2.0;
    return 1.0 + ((double)i + 0.5) / (double)PIMC_LOGTAB_N;
}
PIMC_HD void pimc_logtab_fill(double *tab)
{
    for (int i = 0; i < PIMC_LOGTAB_N; ++i) {
        double c = pimc_logtab_center(i);
        int up = i >= 53;                              /* centre >= sqrt 2: fold one factor 2 into k */
        double invc = floor(512.0 / c + 0.5) / 512.0;  /* 1/c_i rounded to 9 fractional bits: z * invc - 1 is then exact in one fma */
        tab[2 * i] = invc;
        tab[2 * i + 1] = -pimc_log(up ? 2.0 * invc : invc);
    }
}
PIMC_HD double pimc_log_tab(double x, const double *tab)
{
    uint64_t b = pimc_d2bits(x);
    int k = (int)(b >> 52) - 1023;
    int i = (int)((b >> 45) & 0x7Fu);
    double z = pimc_bits2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    double invc = tab[2 * i], lc = tab[2 * i + 1];
    k += (i >= 53);                        /* c_53 = 1.41796875 is the first centre >= sqrt 2 */
    double e = fma(z, invc, -1.0);
    double p = PIMC_K(PIMC_KI_L7);
    p = fma(p, e, PIMC_K(PIMC_KI_L6));
    p = fma(p, e, PIMC_K(PIMC_KI_L5));
    p = fma(p, e, -0.25);
    p = fma(p, e, PIMC_K(PIMC_KI_L3));
    p = fma(p, e, -0.5);
    p = fma(p, e, 1.0);
    double dk = (double)k;
    return fma(dk, PIMC_K(PIMC_KI_LN2HI), lc) + fma(p, e, dk * PIMC_K(PIMC_KI_LN2LO));
}

/* sin and cos of 2*pi*u for u in [0,1): exact quadrant reduction in u, minimax kernels. */
PIMC_HD void pimc_sincos2pi(double u, double *sn, double *cs)
{
    double q = floor(4.0 * u + 0.5);          /* 0..4, exact                    */
    double t = u - 0.25 * q;                  /* exact, |t| <= 1/8              */
    double x = t * PIMC_K(PIMC_KI_TWOPI);     /* |x| <= pi/4                    */
    double z = x * x;
    double ps = PIMC_K(PIMC_KI_S0);
    ps = fma(ps, z, PIMC_K(PIMC_KI_S1));
    ps = fma(ps, z, PIMC_K(PIMC_KI_S2));
    ps = fma(ps, z, PIMC_K(PIMC_KI_S3));
    ps = fma(ps, z, PIMC_K(PIMC_KI_S4));
    ps = fma(ps, z, PIMC_K(PIMC_KI_S5));
    double s0 = fma(x * z, ps, x);
    double pc = PIMC_K(PIMC_KI_C0);
    pc = fma(pc, z, PIMC_K(PIMC_KI_C1));
    pc = fma(pc, z, PIMC_K(PIMC_KI_C2));
    pc = fma(pc, z, PIMC_K(PIMC_KI_C3));
    pc = fma(pc, z, PIMC_K(PIMC_KI_C4));
    pc = fma(pc, z, PIMC_K(PIMC_KI_C5));
    double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
    int iq = (int)q & 3;
    double s1 = (iq & 1) ? c0 : s0;
    double c1 = (iq & 1) ? s0 : c0;
    *sn = (iq == 2 || iq == 3) ? -s1 : s1;
    *cs = (iq == 1 || iq == 2) ? -c1 : c1;
}

/* exp(x) for the Metropolis ratio; +-inf/NaN/overflow/underflow handled explicitly. */
PIMC_HD double pimc_exp(double x)
{
    if (!(x == x)) return x;
    if (x > 709.782712893384) return pimc_bits2d(0x7FF0000000000000ull);
    if (x < -745.2) return 0.0;
    double kf = floor(x * 1.44269504088896338700e+00 + 0.5);
    double hi = x - kf * 6.93147180369123816490e-01;
    double lo = kf * 1.90821492927058770002e-10;
    double r = hi - lo;
    double t = r * r;
    double p = 4.13813679705723846039e-08;
    p = fma(p, t, -1.65339022054652515390e-06);
    p = fma(p, t, 6.61375632143793436117e-05);
    p = fma(p, t, -2.77777777770155933842e-03);
    p = fma(p, t, 1.66666666666666019037e-01);
    double c = r - t * p;
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    int k = (int)kf;
    /* scale by 2^k in two exact steps so that k down to -1075 stays correct */
    int k1 = k / 2, k2 = k - k1;
    y = y * pimc_bits2d((uint64_t)(1023 + k1) << 52);
    y = y * pimc_bits2d((uint64_t)(1023 + k2) << 52);
    return y;
}

/* Box-Muller: 128 random bits -> two independent N(0,1). g0 -> dim 1 (x), g1 -> dim 2 (y).
 * Uniforms carry 52 random bits: d = 1.mantissa in [1,2); u1 = 2 - d in (0,1], u2 = d' - 1 in [0,1). */
PIMC_HD void pimc_gauss_pair_t(pimc_u4 d, const double *tab, double *g0, double *g1)
{
    double d1 = pimc_bits2d(((((uint64_t)d.w[1] << 32) | (uint64_t)d.w[0]) >> 12) | 0x3FF0000000000000ull);
    double d2 = pimc_bits2d(((((uint64_t)d.w[3] << 32) | (uint64_t)d.w[2]) >> 12) | 0x3FF0000000000000ull);
    double u1 = 2.0 - d1, u2 = d2 - 1.0;
    double rad = sqrt(-2.0 * pimc_log_tab(u1, tab));
    double sn, cs;
    pimc_sincos2pi(u2, &sn, &cs);
    *g0 = rad * cs;
    *g1 = rad * sn;
}

#if !defined(__CUDACC__)
/* host table, filled once at library load (see the constructor in the oracle) or lazily */
static double pimc_logtab_host[2 * PIMC_LOGTAB_N];
static int pimc_logtab_host_ready = 0;
static inline const double *pimc_logtab_get(void)
{
    if (!pimc_logtab_host_ready) { pimc_logtab_fill(pimc_logtab_host); pimc_logtab_host_ready = 1; }
    return pimc_logtab_host;
}
static inline void pimc_gauss_pair(pimc_u4 d, double *g0, double *g1) { pimc_gauss_pair_t(d, pimc_logtab_get(), g0, g1); }
#endif

#endif /* PIMC_RNG_H */
