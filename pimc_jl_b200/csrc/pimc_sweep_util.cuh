// pimc_sweep_util.cuh -- helpers shared by the per-iteration sweep kernels (pimc_sweep.cuh: independent worldlines; pimc_isweep.cuh:
// interacting worldlines): division-free teleport, templated potential, update pick, apply! bookkeeping of a whole sweep, async copies.
#pragma once
#include "pimc_moves.cuh"
#include "pimc_launch.h"

// teleport (propagator.jl:30-32) without the IEEE division on the fast path: q = x * (1/2L) differs from x / 2L by
// <= 1 ulp, so floor(q + 0.5) can differ only when q + 0.5 sits within a few ulp of an integer; that case takes the exact path.
__device__ __forceinline__ double d_teleport_fast(double x, double L, double twoL, double inv2L)
{
    double s = x * inv2L + 0.5;
    double f = floor(s);
    double d = s - f;                       // in [0, 1): distance of q + 0.5 to the integer below
    // |q_fast - q_exact| < 1e-9 for |s| < 1e6, so the floors agree unless d is within 1e-9 of 0 or 1 (NaN also lands here)
    if (!(fabs(d - 0.5) <= 0.5 - 1e-9) || !(fabs(s) < 1e6)) f = floor(x / twoL + 0.5);
    return ((x + L) - f * twoL) - L;
}

template <int POT> __device__ __forceinline__ double d_pot_t(const PotDev &p, double x, double y, int dim)
{
    if (POT == PIMC_POT_ZERO) return 0.0;
    if (POT == PIMC_POT_HARMONIC) { double s = x * x; if (dim > 1) s = s + y * y; return (0.5 * p.k) * s; }
    return d_pot(p, x, y, dim);
}

__device__ __forceinline__ int d_pick_update(const SweepParams &P, const pimc_u4 &di)
{
    return d_sample_weighted(P.w, P.nupd, pimc_u01_co(di.w[0], di.w[1]));
}

// apply! bookkeeping of one sweep (simulation.jl:19-27), one thread.  Same final state as d_ring_push per proposal, but the
// ring words are buffered in registers (one global read/write per 32 proposals instead of a dependent RMW per proposal).
__device__ __forceinline__ void d_bookkeep_sweep(const UpdDev &U, int c, const unsigned char *flag, int ntask, unsigned long long beads,
                                                 unsigned long long *stats)
{
    unsigned *ring = U.ring + (size_t)c * U.ring_words;
    const int cap = (int)U.range + 1, range = (int)U.range;
    int head = U.ring_head[c], len = U.ring_len[c], sum = U.ring_sum[c];
    long long tries = U.tries_var[c];
    const long long tries0 = tries;
    long long tr = U.tries[c], ac = U.accepted[c];
    int cnt = 0, tail = head + len; if (tail >= cap) tail -= cap;
    int cw = -1, ew = -1; unsigned cword = 0, eword = 0;
    for (int slot = 0; slot < ntask; ++slot) {
        const int f = flag[slot];
        if (f == 2) continue;
        cnt += 1; tr += 1;
        if (f == 3) continue;
        ac += f; tries += 1;
        const int w = tail >> 5;
        if (w != cw) { if (cw >= 0) ring[cw] = cword; cw = w; cword = ring[w]; }
        const unsigned bit = 1u << (tail & 31);
        cword = f ? (cword | bit) : (cword & ~bit);
        tail = tail + 1 == cap ? 0 : tail + 1;
        len += 1; sum += f;
        if (len > range) {
            const int hw = head >> 5; unsigned eb;
            if (hw == cw) eb = (cword >> (head & 31)) & 1u;
            else { if (hw != ew) { ew = hw; eword = ring[hw]; } eb = (eword >> (head & 31)) & 1u; }
            sum -= (int)eb; head = head + 1 == cap ? 0 : head + 1; len -= 1;
        }
    }
    if (cw >= 0) ring[cw] = cword;
    RingReg R; R.head = head; R.len = len; R.sum = sum; R.tries = tries;
    bool adj = cnt > 0 && (tries / U.adj) != (tries0 / U.adj);
    U.ring_head[c] = head; U.ring_len[c] = len; U.ring_sum[c] = sum; U.tries_var[c] = tries;
    U.tries[c] = tr; U.accepted[c] = ac; U.bead_moves[c] += (long long)beads;
    if (adj) d_adjust(U, c, R);
    if (stats) { atomicAdd(stats + 0, (unsigned long long)cnt); atomicAdd(stats + 2, beads); }
}

// ---- warp-cooperative variant: 32 outcomes per step are gathered with ballots and appended to the ring with word operations.
// Final state (head, len, sum, tries, ring bits, adaptive variable) is identical to pushing the outcomes one by one.
__device__ __forceinline__ unsigned d_ring_read_bits(const unsigned *ring, int cap, int pos, int n) // n <= 32 bits starting at pos (wraps at cap)
{
    const int n1 = n < cap - pos ? n : cap - pos;
    const int w = pos >> 5, o = pos & 31;
    unsigned long long two = (unsigned long long)ring[w];
    if (o + n1 > 32) two |= (unsigned long long)ring[w + 1] << 32;
    unsigned out = (unsigned)((two >> o) & ((n1 >= 32) ? 0xFFFFFFFFull : ((1ull << n1) - 1ull)));
    if (n1 < n) out |= (ring[0] & ((1u << (n - n1)) - 1u)) << n1;
    return out;
}
__device__ __forceinline__ void d_ring_write_seg(unsigned *ring, int pos, int n, unsigned bits) // n <= 32 bits, no wrap inside
{
    const int w = pos >> 5, o = pos & 31;
    const unsigned long long m = ((n >= 32) ? 0xFFFFFFFFull : ((1ull << n) - 1ull)) << o;
    const unsigned long long v = ((unsigned long long)bits << o) & m;
    ring[w] = (ring[w] & ~(unsigned)m) | (unsigned)v;
    if (o + n > 32) ring[w + 1] = (ring[w + 1] & ~(unsigned)(m >> 32)) | (unsigned)(v >> 32);
}
// per-chain counters of one update object, fetched at the START of a sweep kernel (thread 0) so that the global-memory
// latency is hidden behind the moves instead of sitting on the CTA's tail
#define BOOK_PW 6
struct BookPre { int head, len, sum; long long tries, tr, ac; double var; long long range, adj; int tw0, hw0; unsigned tw[BOOK_PW], hw[BOOK_PW]; };
__device__ __forceinline__ BookPre d_book_prefetch(const UpdDev &U, int c)
{
    BookPre b; b.head = U.ring_head[c]; b.len = U.ring_len[c]; b.sum = U.ring_sum[c]; b.tries = U.tries_var[c];
    b.tr = U.tries[c]; b.ac = U.accepted[c]; b.var = U.var[c]; b.range = U.range; b.adj = U.adj;
    // the ring words this sweep can touch: BOOK_PW words from the tail (appends) and from the head (evictions)
    const unsigned *ring = U.ring + (size_t)c * U.ring_words;
    const int cap = (int)b.range + 1;
    int tail = b.head + b.len; if (tail >= cap) tail -= cap;
    b.tw0 = tail >> 5; b.hw0 = b.head >> 5;
    for (int i = 0; i < BOOK_PW; ++i) {
        const int wt = (b.tw0 + i) % U.ring_words, wh = (b.hw0 + i) % U.ring_words;   // small windows wrap more than once
        b.tw[i] = ring[wt]; b.hw[i] = ring[wh];
    }
    return b;
}
// ring word `w` through the prefetched windows (tail window wins: it holds this sweep's own appends)
struct RingCache {
    unsigned *ring; int words; BookPre *pre;
    __device__ __forceinline__ int slot(int w, int w0) const { int d = w - w0; if (d < 0) d += words; return d; }
    __device__ __forceinline__ unsigned get(int w) const {
        int d = slot(w, pre->tw0); if (d < BOOK_PW) return pre->tw[d];
        d = slot(w, pre->hw0); if (d < BOOK_PW) return pre->hw[d];
        return ring[w];
    }
    __device__ __forceinline__ void put(int w, unsigned v) {
        int d = slot(w, pre->tw0); if (d < BOOK_PW) pre->tw[d] = v;
        d = slot(w, pre->hw0); if (d < BOOK_PW) pre->hw[d] = v;
        ring[w] = v;
    }
};
__device__ __forceinline__ unsigned d_ringc_read_bits(const RingCache &rc, int cap, int pos, int n)
{
    const int n1 = n < cap - pos ? n : cap - pos;
    const int w = pos >> 5, o = pos & 31;
    unsigned long long two = (unsigned long long)rc.get(w);
    if (o + n1 > 32) two |= (unsigned long long)rc.get(w + 1) << 32;
    unsigned out = (unsigned)((two >> o) & ((n1 >= 32) ? 0xFFFFFFFFull : ((1ull << n1) - 1ull)));
    if (n1 < n) out |= (rc.get(0) & ((1u << (n - n1)) - 1u)) << n1;
    return out;
}
__device__ __forceinline__ void d_ringc_write_seg(RingCache &rc, int pos, int n, unsigned bits)
{
    const int w = pos >> 5, o = pos & 31;
    const unsigned long long m = ((n >= 32) ? 0xFFFFFFFFull : ((1ull << n) - 1ull)) << o;
    const unsigned long long v = ((unsigned long long)bits << o) & m;
    rc.put(w, (rc.get(w) & ~(unsigned)m) | (unsigned)v);
    if (o + n > 32) rc.put(w + 1, (rc.get(w + 1) & ~(unsigned)(m >> 32)) | (unsigned)(v >> 32));
}
__device__ __forceinline__ void d_bookkeep_sweep_warp(const UpdDev &U, int c, const unsigned char *flag, int ntask, unsigned long long beads,
                                                      unsigned long long *stats, BookPre &pre)
{
    const int lane = threadIdx.x & 31;
    const long long range_ll = pre.range;
    if (range_ll < 64) { if (lane == 0) d_bookkeep_sweep(U, c, flag, ntask, beads, stats); return; }
    RingCache rcache; rcache.ring = U.ring + (size_t)c * U.ring_words; rcache.words = U.ring_words; rcache.pre = &pre;
    const int cap = (int)range_ll + 1, range = (int)range_ll;
    int head = pre.head, len = pre.len, sum = pre.sum;
    long long tries = pre.tries;
    const long long tries0 = tries;
    long long tr = pre.tr, ac = pre.ac;
    int cnt = 0;
    for (int base = 0; base < ntask; base += 32) {
        const int f = base + lane < ntask ? flag[base + lane] : 2;
        const unsigned valid = __ballot_sync(0xffffffffu, f < 2), accb = __ballot_sync(0xffffffffu, f == 1), early = __ballot_sync(0xffffffffu, f == 3);
        if (lane != 0) continue;
        cnt += __popc(valid | early); tr += __popc(valid | early);
        unsigned todo = valid;
        while (todo) {
            // longest run of consecutive valid outcomes starting at the lowest set bit
            const int lo = __ffs(todo) - 1;
            const unsigned run = todo >> lo;
            const int k = (run == 0xFFFFFFFFu) ? 32 : (__ffs(~run) - 1);
            const unsigned bits = (accb >> lo) & ((k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u));
            todo &= (k + lo >= 32) ? 0u : ~((1u << (k + lo)) - 1u);
            const int e = len + k - range;
            if (e > 0) { sum -= __popc(d_ringc_read_bits(rcache, cap, head, e)); head += e; if (head >= cap) head -= cap; len -= e; }
            int tail = head + len; if (tail >= cap) tail -= cap;
            const int k1 = k < cap - tail ? k : cap - tail;
            d_ringc_write_seg(rcache, tail, k1, bits);
            if (k1 < k) d_ringc_write_seg(rcache, 0, k - k1, bits >> k1);
            len += k; sum += __popc(bits); tries += k; ac += __popc(bits);
        }
    }
    if (lane != 0) return;
    bool adj = cnt > 0 && (tries / pre.adj) != (tries0 / pre.adj);
    U.ring_head[c] = head; U.ring_len[c] = len; U.ring_sum[c] = sum; U.tries_var[c] = tries;
    U.tries[c] = tr; U.accepted[c] = ac; atomicAdd((unsigned long long *)&U.bead_moves[c], beads);
    if (adj) { // adjust! (helper.jl:22-52) on the prefetched variable
        double acc = (double)sum / (double)len, v = pre.var;
        if (U.kind == PIMC_UPD_RESHAPE_LINEAR || U.kind == PIMC_UPD_RESHAPE_SWAP) { if (acc < U.minacc) v -= 1; else if (acc > U.maxacc) v += 1; }
        else { if (acc < U.minacc) v *= 0.9; else if (acc > U.maxacc) v *= 1.1; }
        v = U.vmin > v ? U.vmin : v; v = U.vmax < v ? U.vmax : v;
        U.var[c] = v;
    }
    if (stats) { atomicAdd(stats + 0, (unsigned long long)cnt); atomicAdd(stats + 2, beads); }
}

__device__ __forceinline__ uint32_t d_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d_cp_async16(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d_smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void d_cp_async8(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d_smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void d_cp_async4(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d_smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void d_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// apply! counters of (update, chain) fetched asynchronously (cp.async: no registers, no stall) in two steps: the scalars at
// kernel start, the ring words they point at once the scalars have landed.  Same contents as d_book_prefetch.
__device__ __forceinline__ void d_book_prefetch_async1(const UpdDev &U, int c, BookPre *b)
{
    d_cp_async4(&b->head, U.ring_head + c); d_cp_async4(&b->len, U.ring_len + c); d_cp_async4(&b->sum, U.ring_sum + c);
    d_cp_async8(&b->tries, U.tries_var + c); d_cp_async8(&b->tr, U.tries + c); d_cp_async8(&b->ac, U.accepted + c);
    d_cp_async8(&b->var, U.var + c);
    b->range = U.range; b->adj = U.adj;
}
__device__ __forceinline__ void d_book_prefetch_async2(const UpdDev &U, int c, BookPre *b)   // after cp.async.wait_all of step 1
{
    const unsigned *ring = U.ring + (size_t)c * U.ring_words;
    const int cap = (int)b->range + 1;
    int tail = b->head + b->len; if (tail >= cap) tail -= cap;
    b->tw0 = tail >> 5; b->hw0 = b->head >> 5;
    for (int i = 0; i < BOOK_PW; ++i) {
        d_cp_async4(&b->tw[i], ring + (b->tw0 + i) % U.ring_words);
        d_cp_async4(&b->hw[i], ring + (b->hw0 + i) % U.ring_words);
    }
}

