// Pathwise alignment (modes 4 / 5: pathwise_alignment.rs:5-340, pathwise_alignment_semiglobal.rs:6-277, traceback
// pathwise_alignment_output.rs:7-184) and recombination alignment (modes 8 / 9:
// pathwise_alignment_recombination.rs:23-897, tracebacks recombination_output.rs:12-782) on sm_100a —
// the SCORE-TRANSPORT kernel.
//
// What the reference computes: per row and per incoming edge ("group") the LEADER path does a
// linear-gap DP over the read, and every other member path applies the leader's move (D / U / L) to its own scores.
// Consequence used here: on a row with a single group, S[i][j][q] - S[i][j][q'] is COPIED from the source cell of the
// leader's move for every pair of paths. So the scores of all P paths on such a row are
//        S[i][j][q] = base_i[j] + T_k[q][org_i[j]]
// where T_k is the table of absolute scores written by the last row that had several incoming edges (a MATERIALISING
// row), org_i[j] is a column of that table and base_i[j] accumulates the increments of the moves since then.  A
// TRANSPORT row therefore costs one leader DP plus an origin index per column — O(L) instead of O(L * P) — and only
// the rows where paths from different edges meet pay the per-path work.  This is the reference's delta encoding read
// backwards: the deltas never change between merges, so they are kept once per merge instead of once per row.
//
// One CTA of 256 threads per read in flight; thread t owns the CPT contiguous columns [t*CPT, (t+1)*CPT) and keeps the
// previous row's frame (leader score, base, origin) in registers.  The horizontal dependency is an in-thread chain plus
// one CTA-wide max-plus scan; the origin of an L run is found by a second scan that carries (column, origin) of the
// nearest cell that is not an L move.  Two barriers per transport row, no global traffic except 2 bits of leader move
// per column (and, in modes 8/9, the per-(row, column) maxima best_alignment needs).
//
// Traceback: the reference re-derives the arg-max of the CHOSEN path from its own scores.  The chosen path is only
// known at the end, so its scores are replayed from the stored leader moves (one path, O(n * L)), its own arg-max
// codes are written as 2 bits per cell and thread 0 walks them.
// This file is compiled twice (pathwise_tr.cu: 256 threads per CTA, reads of up to 8 191 bases; pathwise_tr_wide.cu: 384
// threads, up to 12 287 bases): PWT_NT = threads per CTA, PWT_FN(name) = the exported symbols of the instance.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

namespace {

constexpr int NT = PWT_NT;      // threads per CTA
constexpr int NWP = NT / 32;   // warps per CTA
enum { MV_D = 1, MV_U = 2, MV_L = 3 };

struct LastRec {   // last-column cell of a finished row, consumed one barrier later by the last warp
    int base;
    unsigned org, row, tid;
};

struct PwtShared {
    int32_t sc[48];
    __align__(16) int totA[NWP];
    __align__(16) long long totK[2][NWP];
    __align__(16) int totA2[2][NWP];
    int zref[2];
    int zconv;
    int dlt[2][NT];
    int val[2][NT];
    __align__(16) unsigned tot2[2][NWP];
    LastRec last[2];
    int best_val, best_set;
    uint32_t best_row, best_path;
    unsigned long long ticket;
    // best_alignment
    float redv[NWP];
    unsigned long long redk1[NWP], redk2[NWP];
    int red_i[NWP];
    float rb_v;
    unsigned long long rb_k1, rb_k2;
    int nsurv;
};

__device__ __forceinline__ int block_excl_max_i(int z, int* tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_max(z, lane);
    if (lane == 31) tot[w] = inc;
    __syncthreads();
    int base = NEG_INF;
#pragma unroll
    for (int k = 0; k < NWP; k++)
        if (k < w) base = max(base, tot[k]);
    int exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = NEG_INF;
    return max(base, exc);
}
// 32-bit non-negative keys; the warp totals are combined with one load + one warp reduction
__device__ __forceinline__ int block_excl_max_r(int z, int* tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_max(z, lane);
    if (lane == 31) tot[w] = inc;
    __syncthreads();
    const int mine = (lane < w) ? tot[lane] : -1;
    const int base = __reduce_max_sync(FULL, mine);
    int exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = -1;
    return max(base, exc);
}
// 64-bit keys (score << 32 | column << 16 | origin): ONE scan gives the incoming chain value and where it comes from
__device__ __forceinline__ long long block_excl_max_k(long long z, long long* tot, long long none) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = z;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) tot[w] = inc;
    __syncthreads();
    long long base = none;
#pragma unroll
    for (int k = 0; k < NWP; k += 2) {
        const longlong2 v = *reinterpret_cast<const longlong2*>(tot + k);
        if (k < w) base = max(base, v.x);
        if (k + 1 < w) base = max(base, v.y);
    }
    long long exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = none;
    return max(base, exc);
}
__device__ __forceinline__ unsigned block_excl_max_u(unsigned z, unsigned* tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = z;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) tot[w] = inc;
    __syncthreads();
    unsigned base = 0;
#pragma unroll
    for (int k = 0; k < NWP; k++)
        if (k < w) base = max(base, tot[k]);
    unsigned exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = 0;
    return max(base, exc);
}

// Two bit planes of my CPT columns (a = low plane, b = high plane) -> bytes [t * CPT / 4, (t + 1) * CPT / 4) of a row of
// LP / 4 bytes. Leader moves: a = L move, b = D move (else U). Own codes of the replayed path: a = D, b = U (else L).
template <int CPT>
__device__ __forceinline__ void store_planes2(uint8_t* rowp, int t, unsigned a, unsigned b) {
    uint8_t* p = rowp + (size_t)t * (CPT / 4);
    if constexpr (CPT == 4)
        *p = (uint8_t)(a | (b << 4));
    else if constexpr (CPT == 8)
        *reinterpret_cast<uint16_t*>(p) = (uint16_t)(a | (b << 8));
    else if constexpr (CPT == 16)
        *reinterpret_cast<uint32_t*>(p) = a | (b << 16);
    else
        *reinterpret_cast<uint2*>(p) = make_uint2(a, b);
}
template <int CPT>
__device__ __forceinline__ void load_planes2(const uint8_t* rowp, int t, unsigned& a, unsigned& b) {
    const uint8_t* p = rowp + (size_t)t * (CPT / 4);
    if constexpr (CPT == 4) {
        const unsigned v = *p;
        a = v & 15u, b = v >> 4;
    } else if constexpr (CPT == 8) {
        const unsigned v = *reinterpret_cast<const uint16_t*>(p);
        a = v & 255u, b = v >> 8;
    } else if constexpr (CPT == 16) {
        const unsigned v = *reinterpret_cast<const uint32_t*>(p);
        a = v & 0xffffu, b = v >> 16;
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        a = v.x, b = v.y;
    }
}

// The leader's linear-gap DP over my columns (pathwise_alignment_semiglobal.rs:38-60): A = leader scores of the
// predecessor row, Am1 = the one left of my first column. Equality tests in the order d, u, l. Thread 0's first cell is
// column 0 (value m0, never an L move). Contains ONE barrier (the scan).
template <int CPT>
__device__ __forceinline__ void leader_dp(const int (&A)[CPT], int Am1, const int (&sv)[CPT], int g, int m0, int j0, int* tot,
                                          int (&nl)[CPT], unsigned& dbits, unsigned& lbits, int& lead_left) {
    int du[CPT];
    dbits = 0;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int d = ((k == 0) ? Am1 : A[k - 1]) + sv[k];
        const int u = A[k] + g;
        du[k] = max(d, u);
        if (d >= u) dbits |= 1u << k;
    }
    if (threadIdx.x == 0) {
        du[0] = m0;
        dbits &= ~1u;
    }
    int v = du[0];
#pragma unroll
    for (int k = 1; k < CPT; k++) v = max(v + g, du[k]);
    const int z = v - (j0 + CPT - 1) * g;
    const int wexc = block_excl_max_i(z, tot);
    int lc = (threadIdx.x == 0) ? NEG_INF : wexc + j0 * g;   // lead[j0 - 1] + g
    lead_left = lc - g;
    lbits = 0;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        int m = du[k];
        if (lc > m) {
            m = lc;
            lbits |= 1u << k;
        }
        nl[k] = m;
        lc = m + g;
    }
}

// CTA-wide maximum of a 64-bit key, returned to every thread (two barriers).
__device__ __forceinline__ long long block_max_ll(long long key, PwtShared& sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int dl = 16; dl >= 1; dl >>= 1) key = max(key, __shfl_xor_sync(FULL, key, dl));
    __syncthreads();
    if (lane == 0) sh.redk1[w] = (unsigned long long)key;
    __syncthreads();
    long long m = (long long)sh.redk1[0];
#pragma unroll
    for (int k = 1; k < NWP; k++) m = max(m, (long long)sh.redk1[k]);
    __syncthreads();
    return m;
}

struct PwtDirBufs {
    uint8_t* mv;     // leader moves of this direction: [group][LP / 4]
    int2* cb;        // modes 8/9 per-(row, column) maxima, or nullptr
    int32_t* lastcol;// forward pass of modes 8/9: [row][Pp], or nullptr
    int32_t* colmax; // modes 8/9: [LP] column maxima of the member entries of cb, or nullptr
};

struct PwtCtx {
    int32_t* tables;
    int32_t* ring_lead;
    uint16_t* ring_org;
    uint4* ring_meta;
    // dynamic shared arrays
    int32_t* s_wb;
    int32_t* s_A;
    uint16_t* s_org;
    uint8_t* s_mv;
    int2* s_mx;
    int32_t* s_res;
    uint32_t* s_end;
    uint32_t LP, LT, Pp, TRmax;
    long long* mcyc;   // diagnostics: cycles spent in materialising rows (thread 0)
};

// Substitution scores of my columns against graph base `lnz`. SIMPLE (match / mismatch tables of score_matrix.rs:35-66): a
// bit test on per-base equality masks built once per read; otherwise a shared-memory table look-up.
template <int CPT, bool SIMPLE>
struct SubScores {
    unsigned eqm[4];                 // SIMPLE: bit k = my column k holds base b
    unsigned rcw[(CPT + 3) / 4];     // read codes of my columns, one byte each
    int s_match, s_mis;
    __device__ __forceinline__ void init(const uint8_t* read, int L, int j0, bool rev, const int32_t* sc) {
#pragma unroll
        for (int q = 0; q < (CPT + 3) / 4; q++) rcw[q] = 0;
        eqm[0] = eqm[1] = eqm[2] = eqm[3] = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = j0 + k;
            // column jj aligns read[jj-1] in the forward pass, read[L-1-jj] in the reverse pass
            const unsigned code = (j >= 1 && j < L) ? (rev ? read[L - 1 - j] : read[j - 1]) : 4u;
            rcw[k / 4] |= code << (8 * (k % 4));
#pragma unroll
            for (unsigned bse = 0; bse < 4; bse++) eqm[bse] |= (code == bse ? 1u : 0u) << k;
        }
        s_match = sc[0];
        s_mis = sc[1];
    }
    __device__ __forceinline__ void get(int lnz, const int32_t* sc, int (&sv)[CPT]) const {
        if constexpr (SIMPLE) {
            const unsigned em = lnz == 0 ? eqm[0] : (lnz == 1 ? eqm[1] : (lnz == 2 ? eqm[2] : (lnz == 3 ? eqm[3] : 0u)));
            const int dm = s_match - s_mis;
#pragma unroll
            for (int k = 0; k < CPT; k++) sv[k] = s_mis + (int)((em >> k) & 1u) * dm;
        } else {
            const int32_t* srow = sc + lnz * 8;
#pragma unroll
            for (int k = 0; k < CPT; k++) sv[k] = srow[(rcw[k / 4] >> (8 * (k % 4))) & 0xffu];
        }
    }
};

// Per-origin maxima over the member paths of row i of table T, RELATIVE to the frame's leader path lam:
//   s_mx[c] = {max_q T[q][c] - T[lam][c], highest q with the maximum | lowest q with it << 16}
// so that the best score over the row's paths at a cell with leader score v and origin c is v + s_mx[c].x
// (modes 8/9 take the highest path id on ties, …_recombination.rs:809-830; the best end of modes 5/9 the lowest).
template <int CPT>
__device__ __forceinline__ void build_mx(const DevPathGraph& g, const PwtCtx& cx, const int32_t* T, uint32_t i, uint32_t lam, int j0,
                                         unsigned ZC) {
    int bv[CPT];
    unsigned bq[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) bv[k] = INT_MIN, bq[k] = 0;
    unsigned lo = 0xffffu, hi = 0;
    for (uint32_t q = 0; q < g.P; q++) {
        if (!((g.node_bits[(size_t)i * g.PW + q / 32] >> (q % 32)) & 1u)) continue;
        if (lo == 0xffffu) lo = q;
        hi = q;
        const int32_t* Tq = T + (size_t)q * cx.LT + j0;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            const int4 v = *reinterpret_cast<const int4*>(Tq + k);
            const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (vv[e] > bv[k + e])
                    bv[k + e] = vv[e], bq[k + e] = q | (q << 16);
                else if (vv[e] == bv[k + e])
                    bq[k + e] = (bq[k + e] & 0xffff0000u) | q;
            }
        }
    }
    const int32_t* Tl = T + (size_t)lam * cx.LT + j0;
#pragma unroll
    for (int k = 0; k < CPT; k += 4) {
        const int4 v = *reinterpret_cast<const int4*>(Tl + k);
        cx.s_mx[j0 + k] = make_int2(bv[k] - v.x, (int)bq[k]);
        cx.s_mx[j0 + k + 1] = make_int2(bv[k + 1] - v.y, (int)bq[k + 1]);
        cx.s_mx[j0 + k + 2] = make_int2(bv[k + 2] - v.z, (int)bq[k + 2]);
        cx.s_mx[j0 + k + 3] = make_int2(bv[k + 3] - v.w, (int)bq[k + 3]);
    }
    if (threadIdx.x == 0) cx.s_mx[ZC] = make_int2(0, (int)(hi | (lo << 16)));
}

// One DP pass over all rows of one direction.
// A row's FRAME is (leader scores `lead`, origins `org`) + (lam = the path the leader scores belong to, tid = table):
//        S[i][j][q] = lead[j] + T_tid[q][org[j]] - T_tid[lam][org[j]]          for every path q of the row.
// K32: the CTA-wide scan of a transport row runs on 32-bit keys (score relative to the previous row's first column, biased,
// 22 / 21 bits | thread 8 / 9 bits) — the host enables it when 3 * max|score| * (LP + 2) < 2^22 (2^21), which bounds every key (a cell of
// row i differs from the row's first column by at most one substitution / gap score per read character).
template <int CPT, bool SIMPLE, bool K32>
__device__ void pwt_pass(const DevPathGraph& g, const PwtCtx& cx, const PwtDirBufs& d, PwtShared& sh, const uint8_t* read,
                         int L, bool rev, bool free_border, bool track_best, bool track_results, bool fpred_rows, int gap) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = g.n, P = g.P, PW = g.PW, RM = g.ring - 1, TM = g.TR - 1;
    int cmax[CPT];   // modes 8/9: per column, the maximum over the rows of the member entries written to d.cb
#pragma unroll
    for (int k = 0; k < CPT; k++) cmax[k] = NEG_INF;
    const uint32_t LP = cx.LP, LT = cx.LT, Pp = cx.Pp;
    const unsigned ZC = LP;                       // the all-zero column every table has
    const int j0 = tid * CPT;
    const size_t tstride = (size_t)Pp * LT;
    const uint32_t base_row = rev ? n - 1 : 0;
    const long long KNONE = (long long)NEG_INF << 32;
    constexpr int TIDBITS = NT > 256 ? 9 : 8;
    constexpr int KBIAS = 1 << (30 - TIDBITS);
    constexpr unsigned ALL = (CPT >= 32) ? 0xffffffffu : ((1u << (CPT % 32)) - 1u);
    const bool need_mx = track_best || d.cb != nullptr;
    SubScores<CPT, SIMPLE> ss;
    ss.init(read, L, j0, rev, sh.sc);

    // ---- frame of the previously processed row, in registers
    int pl[CPT];
    unsigned po[CPT];
    int plm1;
    unsigned pom1;
    uint32_t lam, tidp, prev_row;
    uint32_t mx_lam = 0;   // leader path the per-origin maxima in shared memory are relative to
    int zref = 0;          // K32: leader score of the previous row's first column

    // ---- base row: every path carries the accumulated read gaps (pathwise_alignment_semiglobal.rs:26-32,
    //      pathwise_alignment_recombination.rs:148-155); it creates table 0
    {
        int32_t* Tn = cx.tables;   // table id 0 -> slot 0
        for (uint32_t idx = tid; idx < P * LT; idx += NT) {
            const uint32_t j = idx % LT;
            Tn[(size_t)(idx / LT) * LT + j] = (j < LP) ? (int)j * gap : 0;
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            pl[k] = (j0 + k) * gap;
            po[k] = (unsigned)(j0 + k);
        }
        plm1 = (j0 - 1) * gap;
        pom1 = (unsigned)(j0 - 1);
        lam = g.alphas[base_row];
        tidp = 0;
        prev_row = base_row;
        if (need_mx) {   // every path has the same scores on the base row
#pragma unroll
            for (int k = 0; k < CPT; k++) cx.s_mx[j0 + k] = make_int2(0, (int)P - 1);
            if (tid == 0) cx.s_mx[ZC] = make_int2(0, (int)P - 1);
            mx_lam = lam;
        }
    }
    // running best end cell (mode 5: …_semiglobal.rs:244-277; mode 9 baseline: …_recombination.rs:790-799, which also
    // scans row 0): kept by the thread that owns column L-1
    const bool own_last = j0 <= L - 1 && L - 1 < j0 + CPT;
    bool bset = false;
    int bval = 0;
    uint32_t brow = 0, bpath = 0;
    if (track_best && d.cb) {
        bset = true;
        bval = (L - 1) * gap;
    }
    auto ring_store = [&](uint32_t row) {
        const size_t o = (size_t)(row & RM) * LP + j0;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            *reinterpret_cast<int4*>(cx.ring_lead + o + k) = make_int4(pl[k], pl[k + 1], pl[k + 2], pl[k + 3]);
            *reinterpret_cast<uint2*>(cx.ring_org + o + k) = make_uint2(po[k] | (po[k + 1] << 16), po[k + 2] | (po[k + 3] << 16));
        }
        if (tid == 0) cx.ring_meta[row & RM] = make_uint4(lam, tidp, 0u, 0u);
    };
    if (g.rows[base_row].kind & PWT_RING) ring_store(base_row);
    __syncthreads();

    int par = 0;
    uint4 rnext = reinterpret_cast<const uint4*>(g.rows)[rev ? n - 2 : 1];
    for (uint32_t t = 1; t + 1 < n; t++) {
        const uint32_t i = rev ? n - 1 - t : t;
        PwtRow r;
        r.pred = rnext.x, r.g0 = rnext.y, r.tid = rnext.z;
        r.leader = (uint8_t)(rnext.w & 0xffu), r.nmh = (uint8_t)((rnext.w >> 8) & 0xffu), r.lnz = (uint8_t)((rnext.w >> 16) & 0xffu),
        r.kind = (uint8_t)(rnext.w >> 24);
        if (t + 2 < n) rnext = reinterpret_cast<const uint4*>(g.rows)[rev ? i - 1 : i + 1];
        const int32_t nmh = r.nmh == 255 ? -1 : (int32_t)r.nmh;
        const long long tm0 = (cx.mcyc && !(r.kind & PWT_T)) ? clock64() : 0;
        if (r.kind & PWT_T) {
            // ================= transport row =================
            if (r.pred != prev_row) {
                const size_t o = (size_t)(r.pred & RM) * LP + j0;
#pragma unroll
                for (int k = 0; k < CPT; k += 4) {
                    const int4 a = *reinterpret_cast<const int4*>(cx.ring_lead + o + k);
                    const uint2 oo = *reinterpret_cast<const uint2*>(cx.ring_org + o + k);
                    pl[k] = a.x, pl[k + 1] = a.y, pl[k + 2] = a.z, pl[k + 3] = a.w;
                    po[k] = oo.x & 0xffffu, po[k + 1] = oo.x >> 16, po[k + 2] = oo.y & 0xffffu, po[k + 3] = oo.y >> 16;
                }
                if (tid > 0) {
                    plm1 = cx.ring_lead[o - 1];
                    pom1 = cx.ring_org[o - 1];
                }
                const uint4 mt = cx.ring_meta[r.pred & RM];
                lam = mt.x;
                tidp = mt.y;
                if (K32) zref = cx.ring_lead[(size_t)(r.pred & RM) * LP];
            }
            const int32_t* Tp = cx.tables + (size_t)(tidp & TM) * tstride;
            if ((uint32_t)r.leader != lam) {
                // the leader of this row is not the path the frame's leader scores belong to:
                // S[p][j][a] = lead[j] + T[a][org[j]] - T[lam][org[j]]
                const int32_t* Ta = Tp + (size_t)r.leader * LT;
                const int32_t* Tl = Tp + (size_t)lam * LT;
#pragma unroll
                for (int k = 0; k < CPT; k++) pl[k] += Ta[po[k]] - Tl[po[k]];
                if (tid > 0) plm1 += Ta[pom1] - Tl[pom1];
                if (K32) {   // every thread must use the same reference: thread 0 publishes its converted first column
                    if (tid == 0) sh.zconv = pl[0];
                    __syncthreads();
                    zref = sh.zconv;
                }
                lam = r.leader;
            }
            if (need_mx && ((r.kind & PWT_MXREBUILD) || mx_lam != lam)) {
                __syncthreads();   // the previous row may still be reading s_mx
                if (r.kind & PWT_MXREBUILD) {
                    // per-origin maxima over the paths of THIS row (its path set or table differs from the previous row's)
                    build_mx<CPT>(g, cx, Tp, i, lam, j0, ZC);
                } else {
                    // same table and path set, other leader path: re-express the maxima relative to it
                    const int32_t* Tl = Tp + (size_t)lam * LT + j0;
                    const int32_t* Tm = Tp + (size_t)mx_lam * LT + j0;
#pragma unroll
                    for (int k = 0; k < CPT; k++) cx.s_mx[j0 + k].x += Tm[k] - Tl[k];
                }
                mx_lam = lam;
            }
            // ---- phase A: candidates and the chain inside my columns as if nothing came in from the left; every cell
            // copies its origin from the diagonal / vertical source, or from its left neighbour on an L move
            int nl[CPT];
            unsigned dbits = 0;
            {
                int sv[CPT];
                ss.get(r.lnz, sh.sc, sv);
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const int dd = ((k == 0) ? plm1 : pl[k - 1]) + sv[k];
                    const int uu = pl[k] + gap;
                    nl[k] = max(dd, uu);
                    if (dd >= uu) dbits |= 1u << k;   // equality tests in the order d, u, l (…_semiglobal.rs:46-57)
                }
            }
            if (tid == 0) {
                nl[0] = free_border ? 0 : pl[0] + gap;
                dbits &= ~1u;
            }
            unsigned lbits = 0;
            {
                unsigned old_o = pom1, cur_o = 0;
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const unsigned o_k = po[k];
                    unsigned no = ((dbits >> k) & 1u) ? old_o : o_k;
                    if (k == 0 && tid == 0 && free_border) no = ZC;
                    if (k > 0) {
                        const int c = nl[k - 1] + gap;
                        if (c > nl[k]) {
                            no = cur_o;
                            lbits |= 1u << k;
                        }
                        nl[k] = max(c, nl[k]);
                    }
                    po[k] = no;
                    cur_o = no;
                    old_o = o_k;
                }
            }
            // ---- one scan: the best source left of my columns (normalised chain value; ties go to the right-most source,
            // which is exactly the cell where the reference's strict `l > max(d, u)` test stops an L run) and its origin.
            // All cells after a thread's last source copy that source's origin, so the origin to publish is po[CPT-1].
            int lc_in;
            unsigned org_in;
            if constexpr (K32) {
                if (tid == 0) sh.zref[par] = nl[0];
                const int zrel = nl[CPT - 1] - (j0 + CPT - 1) * gap - zref;
                sh.dlt[par][tid] = (int)po[CPT - 1];
                const int exc = block_excl_max_r((((zrel + KBIAS) << TIDBITS) | tid), sh.totA2[par]);
                lc_in = (tid == 0) ? NEG_INF : (exc >> TIDBITS) - KBIAS + zref + j0 * gap;   // lead[j0 - 1] + gap
                org_in = (unsigned)sh.dlt[par][exc & ((1 << TIDBITS) - 1)];
                zref = sh.zref[par];
            } else {
                const unsigned nonl = ~lbits & ALL;
                const long long key = ((long long)(nl[CPT - 1] - (j0 + CPT - 1) * gap) << 32) |
                                      (long long)(((unsigned)(j0 + 31 - __clz(nonl)) << 16) | po[CPT - 1]);
                const long long exc = block_excl_max_k(key, sh.totK[par], KNONE);
                lc_in = (tid == 0) ? NEG_INF : (int)(exc >> 32) + j0 * gap;
                org_in = (unsigned)exc & 0xffffu;
            }
            // ---- phase B: the incoming chain overrides a prefix of my cells (it loses `gap` per column like every chain)
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                const int nv = max(nl[k], lc_in + k * gap);
                if (nv != nl[k]) {
                    po[k] = org_in;
                    lbits |= 1u << k;
                }
                pl[k] = nv;
            }
            store_planes2<CPT>(d.mv + (size_t)r.g0 * (LP / 4), tid, lbits, dbits);
            plm1 = lc_in - gap;
            pom1 = org_in;
        } else {
            // ================= materialising row: several incoming edges =================
            const uint32_t g0 = r.g0, g1 = g.grp_off[i + 1];
            int32_t* Tn = cx.tables + (size_t)(r.tid & TM) * tstride;
            __syncthreads();   // the previous row's ring copy (frame + meta, written in its tail) must be visible to every thread
            for (uint32_t gi = g0; gi < g1; gi++) {
                const PwGroup gr = g.grp[gi];
                const uint32_t a = gr.leader;
                // frame of the group's predecessor row (always in the ring: the host flags every such row)
                int A[CPT];
                uint32_t ftid;
                {
                    const size_t o = (size_t)(gr.pred & RM) * LP + j0;
                    const uint4 mt = cx.ring_meta[gr.pred & RM];
                    ftid = mt.y;
                    const int32_t* Tf = cx.tables + (size_t)(ftid & TM) * tstride;
                    const int32_t* Ta = Tf + (size_t)a * LT;
                    const int32_t* Tl = Tf + (size_t)mt.x * LT;
#pragma unroll
                    for (int k = 0; k < CPT; k += 4) {
                        const int4 fl = *reinterpret_cast<const int4*>(cx.ring_lead + o + k);
                        const uint2 oo = *reinterpret_cast<const uint2*>(cx.ring_org + o + k);
                        const int fls[4] = {fl.x, fl.y, fl.z, fl.w};
                        const unsigned fos[4] = {oo.x & 0xffffu, oo.x >> 16, oo.y & 0xffffu, oo.y >> 16};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int fb = fls[e] - Tl[fos[e]];   // base: S[p][j][q] = base + T[q][org]
                            A[k + e] = fb + Ta[fos[e]];
                            cx.s_A[j0 + k + e] = A[k + e];
                            cx.s_wb[j0 + k + e] = fb;
                            cx.s_org[j0 + k + e] = (uint16_t)fos[e];
                        }
                    }
                }
                __syncthreads();
                const int Am1 = tid ? cx.s_A[j0 - 1] : NEG_INF;
                int sv[CPT], nl[CPT];
                ss.get(r.lnz, sh.sc, sv);
                unsigned dbits, lbits;
                int lead_left;
                const int m0 = free_border ? 0 : A[0] + gap;
                leader_dp<CPT>(A, Am1, sv, gap, m0, j0, sh.totA, nl, dbits, lbits, lead_left);
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    cx.s_mv[j0 + k] = (uint8_t)(((lbits >> k) & 1u) ? (unsigned)MV_L : (((dbits >> k) & 1u) ? (unsigned)MV_D : (unsigned)MV_U));
                    cx.s_A[j0 + k] = sv[k];   // the member pass reads the substitution scores from here
                }
                store_planes2<CPT>(d.mv + (size_t)gi * (LP / 4), tid, lbits, dbits);
                __syncthreads();
                // ---- members apply the leader's move: one warp per path, lane owns 8 consecutive columns of a 256-column
                // tile. A cell whose move is D or U depends only on the predecessor row; an L run inside the lane is a serial
                // chain; a lane whose leading cells are L moves takes the last value of the nearest lane below that is not
                // all-L (one ballot, one shuffle) plus the gaps in between.
                {
                    constexpr int CW = 8;
                    const int ntile8 = (int)(LP / (32 * CW));
                    const int32_t* Tpb = cx.tables + (size_t)(ftid & TM) * tstride;
                    for (uint32_t q = warp; q < P; q += NWP) {
                        if (!((g.grp_mask[(size_t)gi * PW + q / 32] >> (q % 32)) & 1u)) continue;
                        const int32_t* Tpq = Tpb + (size_t)q * LT;
                        int32_t* Tnq = Tn + (size_t)q * LT;
                        const int col0 = free_border ? 0 : cx.s_wb[0] + Tpq[cx.s_org[0]] + gap;
                        int carry_last = 0, carry_sp = 0;
                        for (int tile = 0; tile < ntile8; tile++) {
                            const int jb = (tile * 32 + lane) * CW;
                            int sp[CW], sjv[CW];
                            unsigned mvw[2];
                            {
                                const int4 w0 = *reinterpret_cast<const int4*>(cx.s_wb + jb), w1 = *reinterpret_cast<const int4*>(cx.s_wb + jb + 4);
                                const uint4 ov = *reinterpret_cast<const uint4*>(cx.s_org + jb);
                                const int wb[CW] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                                const unsigned og[CW] = {ov.x & 0xffffu, ov.x >> 16, ov.y & 0xffffu, ov.y >> 16,
                                                         ov.z & 0xffffu, ov.z >> 16, ov.w & 0xffffu, ov.w >> 16};
#pragma unroll
                                for (int k = 0; k < CW; k++) sp[k] = wb[k] + Tpq[og[k]];
                                const uint2 mw = *reinterpret_cast<const uint2*>(cx.s_mv + jb);
                                mvw[0] = mw.x, mvw[1] = mw.y;
                                const int4 s0 = *reinterpret_cast<const int4*>(cx.s_A + jb), s1 = *reinterpret_cast<const int4*>(cx.s_A + jb + 4);
                                sjv[0] = s0.x, sjv[1] = s0.y, sjv[2] = s0.z, sjv[3] = s0.w;
                                sjv[4] = s1.x, sjv[5] = s1.y, sjv[6] = s1.z, sjv[7] = s1.w;
                            }
                            int spm = __shfl_up_sync(FULL, sp[CW - 1], 1);
                            if (lane == 0) spm = carry_sp;
                            int nv[CW];
                            unsigned lb = 0;
#pragma unroll
                            for (int k = 0; k < CW; k++) {
                                const int j = jb + k;
                                const unsigned m = (mvw[k / 4] >> (8 * (k % 4))) & 0xffu;
                                if (j >= 1 && m == MV_L) lb |= 1u << k;
                                const int spl = (k == 0) ? spm : sp[k - 1];
                                nv[k] = (j == 0) ? col0 : ((m == MV_D) ? spl + sjv[k] : sp[k] + gap);
                            }
                            const int firstn = __ffs(~lb) - 1;   // 0..8 (8: every cell of the lane is an L move)
                            const bool transparent = firstn >= CW;
#pragma unroll
                            for (int k = 1; k < CW; k++)
                                if (((lb >> k) & 1u) && k > firstn) nv[k] = nv[k - 1] + gap;
                            const unsigned fixedm = __ballot_sync(FULL, !transparent);
                            const unsigned below = fixedm & ((1u << lane) - 1u);
                            const int srcl = below ? 31 - __clz(below) : 0;
                            const int lastv = __shfl_sync(FULL, nv[CW - 1], srcl);
                            const int X = below ? lastv + (lane - 1 - srcl) * CW * gap : carry_last + lane * CW * gap;
#pragma unroll
                            for (int k = 0; k < CW; k++)
                                if (k < firstn) nv[k] = X + (k + 1) * gap;
                            *reinterpret_cast<int4*>(Tnq + jb) = make_int4(nv[0], nv[1], nv[2], nv[3]);
                            *reinterpret_cast<int4*>(Tnq + jb + 4) = make_int4(nv[4], nv[5], nv[6], nv[7]);
                            carry_last = __shfl_sync(FULL, nv[CW - 1], 31);
                            carry_sp = __shfl_sync(FULL, sp[CW - 1], 31);
                        }
                        if (lane == 0) Tnq[ZC] = 0;
                    }
                }
                __syncthreads();
            }
            // ---- the row's new frame: origin = own column, leader scores of the row's alpha path
            lam = g.alphas[i];
            tidp = r.tid;
            {
                const int32_t* Tl = Tn + (size_t)lam * LT + j0;
#pragma unroll
                for (int k = 0; k < CPT; k += 4) {
                    const int4 v = *reinterpret_cast<const int4*>(Tl + k);
                    pl[k] = v.x, pl[k + 1] = v.y, pl[k + 2] = v.z, pl[k + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < CPT; k++) po[k] = (unsigned)(j0 + k);
                if (tid > 0) plm1 = Tl[-1];
                pom1 = (unsigned)(j0 - 1);
                if (K32) zref = Tn[(size_t)lam * LT];
            }
            if (need_mx) {
                build_mx<CPT>(g, cx, Tn, i, lam, j0, ZC);
                mx_lam = lam;
                __syncthreads();
            }
        }
        // ---- common tail of a row: per-(row, column) maxima (modes 8/9), ring copy, last-column record
        if (d.cb) {
            int2* out = d.cb + (size_t)i * LP + j0;
#pragma unroll
            for (int k = 0; k < CPT; k += 2) {
                int2 e[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int2 mx = cx.s_mx[po[k + h]];
                    int val = pl[k + h] + mx.x, path = mx.y & 0xffff;
                    bool memb = true;
                    // slots of paths that do not go through the row hold 0 (as in the reference); highest path id wins ties
                    if (nmh >= 0 && (0 > val || (0 == val && nmh > path))) {
                        val = 0;
                        path = nmh;
                        memb = false;
                    }
                    e[h] = make_int2(val, (int)((unsigned)path | (memb ? 0x80000000u : 0u)));
                    if (memb) cmax[k + h] = max(cmax[k + h], val);
                }
                *reinterpret_cast<int4*>(out + k) = make_int4(e[0].x, e[0].y, e[1].x, e[1].y);
            }
        }
        if (cx.mcyc && !(r.kind & PWT_T)) *cx.mcyc += clock64() - tm0;
        if (r.kind & PWT_RING) ring_store(i);
        if ((track_best || (fpred_rows && (r.kind & PWT_FPRED))) && own_last) {
            int llead = pl[0];
            unsigned lorg = po[0];
#pragma unroll
            for (int k = 1; k < CPT; k++)
                if (j0 + k == L - 1) llead = pl[k], lorg = po[k];
            if (track_best) {
                // best member path of the row = leader score + per-origin maximum; first strict maximum in path order (lowest
                // id); a row replaces the incumbent only if strictly better (…_semiglobal.rs:256-273)
                const int2 mx = cx.s_mx[lorg];
                const int v = llead + mx.x;
                if (!bset || v > bval) {
                    bset = true;
                    bval = v;
                    brow = i;
                    bpath = (uint32_t)mx.y >> 16;
                }
            }
            if (fpred_rows && (r.kind & PWT_FPRED)) {
                LastRec lr;
                lr.base = llead, lr.org = lorg, lr.row = lam, lr.tid = tidp;
                sh.last[0] = lr;
            }
        }
        if (fpred_rows && (r.kind & PWT_FPRED)) {
            // predecessor of the end row: per-path scores of its last column (mode 4 results, pathwise_alignment.rs:305-319;
            // mode 8 baseline, …_recombination.rs:777-788). Rare rows: two extra barriers.
            __syncthreads();
            if (warp == 0) {
                const LastRec lr = sh.last[0];   // .row holds the frame's leader path here
                const int32_t* T = cx.tables + (size_t)(lr.tid & TM) * tstride;
                const int lbase = lr.base - T[(size_t)lr.row * LT + lr.org];
                for (uint32_t q = lane; q < Pp; q += 32) {
                    const bool memb = q < P && ((g.node_bits[(size_t)i * PW + q / 32] >> (q % 32)) & 1u);
                    const int v = memb ? lbase + T[(size_t)q * LT + lr.org] : 0;
                    if (d.lastcol) d.lastcol[(size_t)i * Pp + q] = v;
                    if (memb && track_results)
                        for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                            if (g.grp[fg].pred == i && ((g.grp_mask[(size_t)fg * PW + q / 32] >> (q % 32)) & 1u)) {
                                cx.s_res[q] = v;
                                cx.s_end[q] = i;
                            }
                }
            }
            __syncthreads();
        }
        prev_row = i;
        par ^= 1;
    }
    if (track_best && own_last) {
        sh.best_set = bset ? 1 : 0;
        sh.best_val = bval;
        sh.best_row = brow;
        sh.best_path = bpath;
    }
    if (d.colmax) {
#pragma unroll
        for (int k = 0; k < CPT; k++) d.colmax[j0 + k] = cmax[k];
    }
    __syncthreads();
}

// Scores of ONE path replayed from the stored leader moves, rows up to `limit` (inclusive) in processing order; writes the
// path's own arg-max code (build_alignment's order: d, then u, else l — pathwise_alignment_output.rs:80-109) per cell and
// the path's predecessor per row.
template <int CPT, bool SIMPLE>
__device__ void pwt_replay(const DevPathGraph& g, const PwtCtx& cx, const uint8_t* mv, uint8_t* own, uint32_t* own_pred,
                           PwtShared& sh, const uint8_t* read, int L, bool rev, bool free_border, uint32_t q, uint32_t limit, int gap,
                           bool find_end = false) {
    const int tid = threadIdx.x;
    const uint32_t n = g.n, PW = g.PW, LP = cx.LP;
    const int j0 = tid * CPT;
    constexpr unsigned ALL = (CPT >= 32) ? 0xffffffffu : ((1u << (CPT % 32)) - 1u);
    SubScores<CPT, SIMPLE> ss;
    ss.init(read, L, j0, rev, sh.sc);
    unsigned inread = 0;   // columns 1 .. L-1 among mine
#pragma unroll
    for (int k = 0; k < CPT; k++)
        if (j0 + k >= 1 && j0 + k < L) inread |= 1u << k;
    int prev[CPT], prevm1 = (j0 - 1) * gap;
#pragma unroll
    for (int k = 0; k < CPT; k++) prev[k] = (j0 + k) * gap;
    const bool own_last = j0 <= L - 1 && L - 1 < j0 + CPT;
    bool eset = false;
    int eval = 0;
    uint32_t erow = 0;
    int par = 0;
    for (uint32_t t = 1; t + 1 < n; t++) {
        const uint32_t i = rev ? n - 1 - t : t;
        if (rev ? i < limit : i > limit) break;
        if (!((g.node_bits[(size_t)i * PW + q / 32] >> (q % 32)) & 1u)) continue;
        const PwtRow r = g.rows[i];
        uint32_t gi = r.g0, pred = r.pred;
        if (!(r.kind & PWT_T)) {
            for (uint32_t gg = r.g0; gg < g.grp_off[i + 1]; gg++)
                if ((g.grp_mask[(size_t)gg * PW + q / 32] >> (q % 32)) & 1u) gi = gg;
            pred = g.grp[gi].pred;
        }
        unsigned lbits, dbits;
        load_planes2<CPT>(mv + (size_t)gi * (LP / 4), tid, lbits, dbits);
        if (tid == 0) lbits &= ~1u;
        // reverse pass: the 'F' row is never made absolute by the reference (absolute_scores stops before it,
        // pathwise_alignment_recombination.rs:748), so its traceback sees 0 for every path but path 0
        const bool quirk = rev && pred == n - 1 && q != 0;
        int sv[CPT], cur[CPT];
        ss.get(r.lnz, sh.sc, sv);
        int lastnorm = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            if ((lbits >> k) & 1u) {
                cur[k] = (k == 0) ? 0 : cur[k - 1] + gap;
            } else {
                cur[k] = ((dbits >> k) & 1u) ? ((k == 0) ? prevm1 : prev[k - 1]) + sv[k] : prev[k] + gap;
                if (k == 0 && tid == 0) cur[0] = free_border ? 0 : prev[0] + gap;
                lastnorm = cur[k] - (j0 + k) * gap;
            }
        }
        const unsigned nonl = ~lbits & ALL;
        const int first = nonl ? __ffs(nonl) - 1 : CPT;
        sh.val[par][tid] = lastnorm;
        const unsigned key = nonl ? (unsigned)(j0 + 31 - __clz(nonl) + 1) : 0u;
        const unsigned exc = block_excl_max_u(key, sh.tot2[par]);
        const int vin = (tid == 0) ? 0 : sh.val[par][(exc - 1) / CPT];
        const int cur_left = vin + (j0 - 1) * gap;   // S[i][j0-1][q]
#pragma unroll
        for (int k = 0; k < CPT; k++)
            if (k < first) cur[k] = vin + (j0 + k) * gap;
        unsigned od = 0, ou = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int spl = (k == 0) ? prevm1 : prev[k - 1];
            const int lq = ((k == 0) ? cur_left : cur[k - 1]) + gap;
            const int dq = (quirk ? 0 : spl) + sv[k], uq = (quirk ? 0 : prev[k]) + gap;
            const int bq = max(dq, max(uq, lq));
            if (bq == dq)
                od |= 1u << k;
            else if (bq == uq)
                ou |= 1u << k;
        }
        store_planes2<CPT>(own + (size_t)i * (LP / 4), tid, od & inread, ou & inread);
        if (tid == 0) own_pred[i] = pred;
        if (find_end && own_last) {
            // ending_node (…_recombination.rs:885-897): first strict maximum of the path's last-column scores
            int v = cur[0];
#pragma unroll
            for (int k = 1; k < CPT; k++)
                if (j0 + k == L - 1) v = cur[k];
            if (!eset || v > eval) {
                eset = true;
                eval = v;
                erow = i;
            }
        }
        prevm1 = cur_left;
#pragma unroll
        for (int k = 0; k < CPT; k++) prev[k] = cur[k];
        par ^= 1;
    }
    if (find_end && own_last) {
        sh.best_row = erow;
        sh.best_val = eval;
    }
    __syncthreads();
}

// own arg-max of the replayed path at (row, col): MV_D / MV_U / MV_L
template <int CPT>
__device__ __forceinline__ unsigned own_code(const uint8_t* own, uint32_t LP, uint32_t row, int col) {
    unsigned a, b;
    load_planes2<CPT>(own + (size_t)row * (LP / 4), col / CPT, a, b);
    const int k = col % CPT;
    return ((a >> k) & 1u) ? (unsigned)MV_D : (((b >> k) & 1u) ? (unsigned)MV_U : (unsigned)MV_L);
}

// Forward-direction walk shared by build_alignment (pathwise_alignment_output.rs:7-184), the no_rec builders
// (recombination_output.rs:239-361,633-782) and the forward half of the rec builders (:100-163,472-557).
template <int CPT>
__device__ void pwt_walk_fwd(const DevPathGraph& g, const uint8_t* own, const uint32_t* own_pred, uint32_t LP, const uint8_t* read,
                             uint32_t& ii, int& j, bool pad_global, RunEmitter& em) {
    while (ii > 0 && j > 0) {
        const unsigned code = own_code<CPT>(own, LP, ii, j);
        if (code == MV_D) {
            em.step(g.lnz[ii] != read[j - 1] ? RG_OP_d : RG_OP_D, ii, 0);
            ii = own_pred[ii];
            j--;
        } else if (code == MV_U) {
            em.step(RG_OP_U, ii, 0);
            ii = own_pred[ii];
        } else {
            em.step(RG_OP_L, ii, 0);
            j--;
        }
    }
    while (j > 0) {
        em.step(RG_OP_L, ii, 0);
        j--;
    }
    if (pad_global)
        while (ii > 0) {
            em.step(RG_OP_U, ii, 0);
            ii = own_pred[ii];
        }
}

struct RecBest {   // state of best_alignment's reduction
    float v;            // maximum candidate score
    unsigned long long k1;  // first (j,i,ri) with score v
    unsigned long long k2;  // first (j,i,ri) with score v on a segment edge
};
__device__ __forceinline__ void rec_merge(RecBest& a, float v, unsigned long long key, bool edge) {
    if (v > a.v) {
        a.v = v;
        a.k1 = key;
        a.k2 = edge ? key : ~0ull;
    } else if (v == a.v) {
        a.k1 = min(a.k1, key);
        if (edge) a.k2 = min(a.k2, key);
    }
}
__device__ __forceinline__ void rec_merge2(RecBest& a, const RecBest& o) {
    if (o.v > a.v)
        a = o;
    else if (o.v == a.v) {
        a.k1 = min(a.k1, o.k1);
        a.k2 = min(a.k2, o.k2);
    }
}

template <int CPT, bool SIMPLE, bool K32>
__global__ void __launch_bounds__(NT, (NT > 256) ? 1 : ((CPT <= 8) ? 4 : ((CPT <= 16) ? 2 : 1)))
    k_pathwise_tr(DevPathGraph g, DevPathGraph rg_, DevScoring sc, PwtWorkspace ws, PoaBatch b, int mode) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ PwtShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slot = blockIdx.x;
    if (tid < 48) sh.sc[tid] = (&sc.sc[0][0])[tid];
    const uint32_t n = g.n, P = g.P, PW = g.PW, LP = ws.LP, LT = ws.LT, Pp = ws.Pp;
    const bool rec_mode = mode == RG_MODE_REC_GLOBAL || mode == RG_MODE_REC_SEMIGLOBAL;
    const bool global_mode = mode == RG_MODE_PATHWISE_GLOBAL || mode == RG_MODE_REC_GLOBAL;
    PwtCtx cx;
    cx.LP = LP, cx.LT = LT, cx.Pp = Pp, cx.TRmax = ws.TRmax;
    cx.tables = ws.tables + (size_t)slot * ws.TRmax * Pp * LT;
    cx.ring_lead = ws.ring_lead + (size_t)slot * ws.ringmax * LP;
    cx.ring_org = ws.ring_org + (size_t)slot * ws.ringmax * LP;
    cx.ring_meta = ws.ring_meta + (size_t)slot * ws.ringmax;
    {
        size_t off = 0;
        if (ws.stage) {   // wide instance: the two 32-bit staging rows of the (rare) materialising rows live in HBM / L2
            cx.s_wb = ws.stage + (size_t)slot * 2 * LP;
            cx.s_A = cx.s_wb + LP;
        } else {
            cx.s_wb = reinterpret_cast<int32_t*>(s_dyn + off);
            off += (size_t)LP * 4;
            cx.s_A = reinterpret_cast<int32_t*>(s_dyn + off);
            off += (size_t)LP * 4;
        }
        cx.s_res = reinterpret_cast<int32_t*>(s_dyn + off);
        off += (size_t)Pp * 4;
        cx.s_end = reinterpret_cast<uint32_t*>(s_dyn + off);
        off += (size_t)Pp * 4;
        cx.s_org = reinterpret_cast<uint16_t*>(s_dyn + off);
        off += (size_t)LP * 2;
        cx.s_mv = s_dyn + off;
        off += (size_t)LP;
        cx.s_mx = reinterpret_cast<int2*>(s_dyn + off);   // modes 8/9 only: [LT]
    }
    uint32_t* s_surv = reinterpret_cast<uint32_t*>(cx.s_mx);   // best_alignment: survivors in the per-origin maxima's space (2 * LT words)
    PwtDirBufs fwd, rvd;
    fwd.mv = ws.mv_f + (size_t)slot * g.n_groups * (LP / 4);
    fwd.cb = rec_mode ? ws.cb_f + (size_t)slot * n * LP : nullptr;
    fwd.lastcol = rec_mode ? ws.lastcol + (size_t)slot * n * Pp : nullptr;
    rvd.mv = rec_mode ? ws.mv_r + (size_t)slot * rg_.n_groups * (LP / 4) : nullptr;
    rvd.cb = rec_mode ? ws.cb_r + (size_t)slot * n * LP : nullptr;
    rvd.lastcol = nullptr;
    fwd.colmax = rec_mode ? ws.colmax + (size_t)slot * 2 * LP : nullptr;
    rvd.colmax = rec_mode ? ws.colmax + (size_t)slot * 2 * LP + LP : nullptr;
    uint8_t* own = ws.own + (size_t)slot * n * (LP / 4);
    uint32_t* own_pred = ws.own_pred + (size_t)slot * n;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const int gap = sc.sc[0][5];  // uniform gap score (checked on the host)
    __syncthreads();

    for (;;) {
        if (tid == 0) sh.ticket = atomicAdd(&b.counters[0], 1ull);
        __syncthreads();
        const unsigned long long ticket = sh.ticket;
        __syncthreads();
        if (ticket >= (unsigned long long)b.n_reads) break;
        const int ridx = b.order ? b.order[ticket] : (int)ticket;
        const uint8_t* read = b.reads + b.read_off[ridx];
        const int L = (int)(b.read_off[ridx + 1] - b.read_off[ridx]) + 1;
        rg_read_result res;
        res.status = 0;
        res.score = 0;
        res.score_f32 = 0.f;
        res.displacement = 0;
        res.end_row = res.end_col = res.start_row = res.start_col = 0;
        res.best_path = res.rev_best_path = 0;
        res.fen = res.rsn = res.rec_col = res.rev_end_row = 0;
        res.cells = (uint64_t)(n - 2) * (uint64_t)(L - 1) * (rec_mode ? 2 : 1);
        res.run_off = 0;
        res.n_runs = 0;
        res.n_runs_rev = 0;
        if ((uint32_t)L > LP) {
            res.status = RG_READ_TRACE_OVERFLOW;
            if (tid == 0) b.results[ridx] = res;
            continue;
        }
        if (tid == 0) {
            sh.best_set = 0;
            sh.best_val = 0;
            sh.best_row = 0;
            sh.best_path = 0;
        }
        for (uint32_t q = tid; q < Pp; q += NT) {
            cx.s_res[q] = 0;
            cx.s_end[q] = 0;
        }
        __syncthreads();
        long long mcyc = 0;
        cx.mcyc = ws.diag ? &mcyc : nullptr;
        const long long tc0 = clock64();
        pwt_pass<CPT, SIMPLE, K32>(g, cx, fwd, sh, read, L, false, !global_mode, !global_mode, mode == RG_MODE_PATHWISE_GLOBAL,
                              global_mode, gap);
        const long long tc1 = clock64();
        if (rec_mode) pwt_pass<CPT, SIMPLE, K32>(rg_, cx, rvd, sh, read, L, true, !global_mode, false, false, false, gap);
        const long long tc2 = clock64();
        long long tc3 = tc2;

        if (!rec_mode) {
            // ================= modes 4 / 5: end cell, replay of the chosen path, traceback (thread 0), publish =================
            uint32_t best_path, ending;
            int score;
            if (global_mode) {
                // max of (score, path): highest path id wins ties (pathwise_alignment.rs:320-325)
                best_path = 0;
                for (uint32_t q = 1; q < P; q++)
                    if (cx.s_res[q] >= cx.s_res[best_path]) best_path = q;
                ending = cx.s_end[best_path];
                score = cx.s_res[best_path];
            } else {
                best_path = sh.best_path;
                ending = sh.best_row;
                score = sh.best_val;
            }
            __syncthreads();
            pwt_replay<CPT, SIMPLE>(g, cx, fwd.mv, own, own_pred, sh, read, L, false, !global_mode, best_path, ending, gap);
            if (tid == 0) {
                res.score = score;
                res.best_path = best_path;
                res.end_row = ending;
                res.end_col = (uint32_t)(L - 1);
                RunEmitter em;
                em.init(runs, ws.run_cap);
                uint32_t ii = ending;
                int j = L - 1;
                pwt_walk_fwd<CPT>(g, own, own_pred, LP, read, ii, j, global_mode, em);
                em.flush(0);
                res.start_row = ii;
                if (em.overflow) res.status |= RG_READ_TRACE_OVERFLOW;
                uint32_t nr = em.overflow ? 0 : em.n;
                unsigned long long ro = atomicAdd(&b.counters[1], (unsigned long long)nr);
                if (ro + nr > b.out_run_cap) {
                    res.status |= RG_READ_TRACE_OVERFLOW;
                    nr = 0;
                }
                for (uint32_t k = 0; k < nr; k++) b.out_runs[ro + k] = runs[k];
                res.run_off = ro;
                res.n_runs = nr;
                if (ws.diag) {   // kilo-cycles: forward pass, materialising rows, replay + walk
                    res.fen = (uint32_t)((tc1 - tc0) >> 10);
                    res.rsn = (uint32_t)(mcyc >> 10);
                    res.rev_end_row = (uint32_t)((clock64() - tc2) >> 10);
                }
                b.results[ridx] = res;
            }
            __syncthreads();
            continue;
        }

        // ================= modes 8 / 9: best_alignment (…_recombination.rs:759-873) =================
        // 1. baseline (no recombination)
        if (tid == 0) {
            bool has = false;
            int mx = 0;
            uint32_t bp = 0;
            if (mode == RG_MODE_REC_GLOBAL) {
                for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++) {
                    const uint32_t pred = g.grp[fg].pred;
                    for (uint32_t q = 0; q < P; q++)
                        if ((g.grp_mask[(size_t)fg * PW + q / 32] >> (q % 32)) & 1u) {
                            const int v = fwd.lastcol[(size_t)pred * Pp + q];
                            if (!has || mx < v) {
                                has = true;
                                mx = v;
                                bp = q;
                            }
                        }
                }
            }
            if (mode == RG_MODE_REC_GLOBAL) {
                sh.best_val = mx;
                sh.best_path = bp;
            }
        }
        __syncthreads();
        // mode 9: the forward pass already tracked the first strict maximum in (row, path) order over the member slots of rows
        // 0..n-2 (…_recombination.rs:790-799) in sh.best_val / sh.best_path
        if (tid == 0) {
            sh.rb_v = (float)sh.best_val;
            sh.rb_k1 = ~0ull;
            sh.rb_k2 = ~0ull;
        }
        __syncthreads();
        const int base_score = sh.best_val;
        const uint32_t base_path = sh.best_path;
        // 2. candidates. out_of_band = max((L * (1 - B) / 2) as i32, 1)
        int oob;
        {
            const float t1 = __fmul_rn((float)L, __fsub_rn(1.0f, sc.rbw));
            const float t2 = __fdiv_rn(t1, 2.0f);
            int oi = (t2 != t2) ? 0 : (t2 >= 2147483648.0f ? 2147483647 : (t2 <= -2147483648.0f ? (-2147483647 - 1) : (int)t2));
            oob = max(oi, 1);
        }
        const float Rf = (float)sc.R;
        const bool prune = sc.r >= 0.0f;
        RecBest mine;
        mine.v = (float)base_score;
        mine.k1 = ~0ull;
        mine.k2 = ~0ull;
        const uint32_t surv_cap = min((uint32_t)REC_SURV, 2 * LT);
        // Column bounds: no pair of column j can score above (max_i m[i][j] + max_ri w[ri][j]) - R when the displacement
        // multiplier is not negative. The maxima were collected by the two passes; columns whose bound is below the running
        // maximum are skipped without touching their entries, and the column with the best bound goes first so that the
        // running maximum is close to final from the start. (The acceptance rule is order-free, see step 4.)
        int32_t* s_cmf = cx.s_wb;   // staging arrays of the DP are free now
        int32_t* s_cmr = cx.s_A;
        for (uint32_t j = tid; j < LP; j += NT) {
            s_cmf[j] = fwd.colmax[j];
            s_cmr[j] = rvd.colmax[j];
        }
        __syncthreads();
        int first_col = -1;
        if (prune) {
            long long bk = LLONG_MIN;
            for (int j = oob + tid; j < L - oob; j += NT) {
                const int a = s_cmf[j], bb = s_cmr[L - 1 - j];
                if (a > NEG_INF / 2 && bb > NEG_INF / 2) bk = max(bk, ((long long)(a + bb) << 32) | (long long)(0xffffffffu - (uint32_t)j));
            }
            bk = block_max_ll(bk, sh);
            if (bk != LLONG_MIN) first_col = (int)(0xffffffffu - (uint32_t)(bk & 0xffffffffll));
        }
        for (int jt = (first_col >= 0 ? oob - 1 : oob); jt < L - oob; jt++) {
            const int j = (jt < oob) ? first_col : jt;
            const int jj = L - 1 - j;  // column of w in the reverse pass's coordinates
            const int wmax = s_cmr[jj];
            const float cur = sh.rb_v;
            if (wmax <= NEG_INF / 2 || s_cmf[j] <= NEG_INF / 2) continue;
            if (prune && __fsub_rn((float)(s_cmf[j] + wmax), Rf) < cur) continue;
            if (tid == 0) sh.nsurv = 0;
            __syncthreads();
            // survivors: forward nodes whose best case can still reach the running maximum
            for (uint32_t i = 1 + tid; i + 1 < n; i += NT) {
                const int2 e = fwd.cb[(size_t)i * LP + j];
                if (e.y >= 0) continue;  // arg-max slot is not a member path (…_recombination.rs:833)
                const float ub = __fsub_rn((float)(e.x + wmax), Rf);
                if (!prune || ub >= cur) {
                    const int pos = atomicAdd(&sh.nsurv, 1);
                    if ((uint32_t)pos < surv_cap) s_surv[pos] = i;
                }
            }
            __syncthreads();
            const int nsurv = sh.nsurv;
            if ((uint32_t)nsurv > surv_cap) {
                // too many to stage: every thread walks all forward nodes itself (exact, just slower)
                for (uint32_t i = 1; i + 1 < n; i++) {
                    const int2 fe = fwd.cb[(size_t)i * LP + j];
                    if (fe.y >= 0) continue;
                    const uint32_t fp = (uint32_t)fe.y & 0x7fffffffu;
                    const uint64_t seg_i = g.seg[i];
                    const bool iedge = seg_i != g.seg[i + 1];
                    for (uint32_t ri = 1 + tid; ri + 1 < n; ri += NT) {
                        const int2 we = rvd.cb[(size_t)ri * LP + jj];
                        if (we.y >= 0) continue;
                        const uint32_t rp = (uint32_t)we.y & 0x7fffffffu;
                        if (g.seg[ri] == seg_i || fp == rp) continue;
                        const int dd = abs(g.dfs[i] - g.dfs[ri]) + abs(g.dfe[i] - g.dfe[ri]);
                        const float pen = __fadd_rn(Rf, __fmul_rn(sc.r, (float)dd));
                        const float nv = __fsub_rn((float)(fe.x + we.x), pen);
                        const bool edge = iedge && g.seg[ri] != g.seg[ri - 1];
                        rec_merge(mine, nv, ((unsigned long long)j << 42) | ((unsigned long long)i << 21) | ri, edge);
                    }
                }
            } else {
                for (int s = 0; s < nsurv; s++) {
                    const uint32_t i = s_surv[s];
                    const int2 fe = fwd.cb[(size_t)i * LP + j];
                    const uint32_t fp = (uint32_t)fe.y & 0x7fffffffu;
                    const uint64_t seg_i = g.seg[i];
                    const bool iedge = seg_i != g.seg[i + 1];
                    const int dfs_i = g.dfs[i], dfe_i = g.dfe[i];
                    for (uint32_t ri = 1 + tid; ri + 1 < n; ri += NT) {
                        const int2 we = rvd.cb[(size_t)ri * LP + jj];
                        if (we.y >= 0) continue;
                        const uint32_t rp = (uint32_t)we.y & 0x7fffffffu;
                        if (g.seg[ri] == seg_i || fp == rp) continue;
                        const int dd = abs(dfs_i - g.dfs[ri]) + abs(dfe_i - g.dfe[ri]);
                        const float pen = __fadd_rn(Rf, __fmul_rn(sc.r, (float)dd));
                        const float nv = __fsub_rn((float)(fe.x + we.x), pen);
                        const bool edge = iedge && g.seg[ri] != g.seg[ri - 1];
                        rec_merge(mine, nv, ((unsigned long long)j << 42) | ((unsigned long long)i << 21) | ri, edge);
                    }
                }
            }
            // publish the running maximum so that later columns prune against it
            float wv = mine.v;
#pragma unroll
            for (int dlt = 16; dlt >= 1; dlt >>= 1) wv = fmaxf(wv, __shfl_xor_sync(FULL, wv, dlt));
            if (lane == 0) sh.redv[warp] = wv;
            __syncthreads();
            if (tid == 0) {
                float m = sh.rb_v;
                for (int k = 0; k < NWP; k++) m = fmaxf(m, sh.redv[k]);
                sh.rb_v = m;
            }
            __syncthreads();
        }
        // 3. reduce (max score; first key with it; first edge key with it)
        {
            RecBest r = mine;
#pragma unroll
            for (int dlt = 16; dlt >= 1; dlt >>= 1) {
                RecBest o;
                o.v = __shfl_xor_sync(FULL, r.v, dlt);
                o.k1 = __shfl_xor_sync(FULL, r.k1, dlt);
                o.k2 = __shfl_xor_sync(FULL, r.k2, dlt);
                rec_merge2(r, o);
            }
            if (lane == 0) {
                sh.redv[warp] = r.v;
                sh.redk1[warp] = r.k1;
                sh.redk2[warp] = r.k2;
            }
            __syncthreads();
            if (tid == 0) {
                RecBest t;
                t.v = sh.redv[0];
                t.k1 = sh.redk1[0];
                t.k2 = sh.redk2[0];
                for (int k = 1; k < NWP; k++) {
                    RecBest o;
                    o.v = sh.redv[k];
                    o.k1 = sh.redk1[k];
                    o.k2 = sh.redk2[k];
                    rec_merge2(t, o);
                }
                sh.rb_v = t.v;
                sh.rb_k1 = t.k1;
                sh.rb_k2 = t.k2;
            }
            __syncthreads();
        }
        tc3 = clock64();
        // 4. outcome (every thread derives it: the replays below are CTA-wide)
        const float basef = (float)base_score;
        // sequential acceptance rule restated: a candidate is taken if it beats the incumbent, or ties it while
        // the incumbent is not on a segment edge and the candidate is (…_recombination.rs:844-851)
        // If the maximum beats the baseline: the first candidate with it wins unless it is off-edge and an on-edge
        // tie follows (k2, the first on-edge tie, equals k1 when k1 itself is on an edge). If the maximum only
        // ties the baseline: the first on-edge candidate with it, if any.
        unsigned long long key = ~0ull;
        {
            const float tv = sh.rb_v;
            const unsigned long long k1 = sh.rb_k1, k2 = sh.rb_k2;
            if (k1 != ~0ull) {
                if (tv > basef)
                    key = (k2 != ~0ull) ? k2 : k1;
                else if (tv == basef)
                    key = k2;
            }
        }
        RunEmitter em;
        em.init(runs, ws.run_cap);
        res.score = base_score;
        if (key == ~0ull) {
            // no recombination: gaf_output_{global,semiglobal}_no_rec
            uint32_t ending = 0;
            if (mode == RG_MODE_REC_GLOBAL) {
                for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                    if ((g.grp_mask[(size_t)fg * PW + base_path / 32] >> (base_path % 32)) & 1u) ending = g.grp[fg].pred;
            }
            int end_score = 0;
            if (mode == RG_MODE_REC_GLOBAL) {
                pwt_replay<CPT, SIMPLE>(g, cx, fwd.mv, own, own_pred, sh, read, L, false, false, base_path, ending, gap);
                end_score = fwd.lastcol[(size_t)ending * Pp + base_path];
            } else {
                // ending_node (…_recombination.rs:885-897) needs the path's last-column score on every row: the replay finds it
                pwt_replay<CPT, SIMPLE>(g, cx, fwd.mv, own, own_pred, sh, read, L, false, true, base_path, n - 2, gap, true);
                ending = sh.best_row;
                end_score = sh.best_val;
            }
            if (tid == 0) {
                res.best_path = res.rev_best_path = base_path;
                res.end_row = ending;
                res.end_col = (uint32_t)(L - 1);
                res.score = end_score;
                res.score_f32 = (float)base_score;
                uint32_t ii = ending;
                int j = L - 1;
                pwt_walk_fwd<CPT>(g, own, own_pred, LP, read, ii, j, mode == RG_MODE_REC_GLOBAL, em);
                em.flush(0);
                res.start_row = ii;
                res.n_runs = em.n;
            }
        } else {
            const uint32_t rcol = (uint32_t)(key >> 42), fen = (uint32_t)((key >> 21) & 0x1fffffu), rsn = (uint32_t)(key & 0x1fffffu);
            const uint32_t fp = (uint32_t)fwd.cb[(size_t)fen * LP + rcol].y & 0x7fffffffu;
            const uint32_t rp = (uint32_t)rvd.cb[(size_t)rsn * LP + (L - 1 - rcol)].y & 0x7fffffffu;
            pwt_replay<CPT, SIMPLE>(g, cx, fwd.mv, own, own_pred, sh, read, L, false, !global_mode, fp, fen, gap);
            uint32_t ii = fen;
            if (tid == 0) {
                res.status |= RG_READ_RECOMBINATION;
                res.best_path = fp;
                res.rev_best_path = rp;
                res.fen = fen;
                res.rsn = rsn;
                res.rec_col = rcol;
                res.score_f32 = sh.rb_v;
                res.displacement = abs(g.dfs[fen] - g.dfs[rsn]) + abs(g.dfe[fen] - g.dfe[rsn]);
                // forward half, traceback order (recombination_output.rs:100-163 / 472-557)
                int j = (int)rcol;
                pwt_walk_fwd<CPT>(g, own, own_pred, LP, read, ii, j, mode == RG_MODE_REC_GLOBAL, em);
                em.flush(0);
                res.start_row = ii;
                res.n_runs = em.n;
            }
            __syncthreads();   // the walk is done with the own-code buffer before the reverse replay rewrites it
            pwt_replay<CPT, SIMPLE>(rg_, cx, rvd.mv, own, own_pred, sh, read, L, true, !global_mode, rp, rsn, gap);
            if (tid == 0) {
                // reverse half, forward order (:38-98 / 389-470): rows ascend
                em.ascending = true;
                uint32_t ri = rsn;
                int cj = (int)rcol;
                uint32_t rev_end = ri;
                while (ri > 0 && ri < n - 1 && cj < L - 1) {
                    const unsigned code = own_code<CPT>(own, LP, ri, L - 1 - cj);
                    rev_end = ri;
                    if (code == MV_D) {
                        em.step(g.lnz[ri] != read[cj] ? RG_OP_d : RG_OP_D, ri, 0);  // r_seq[j] = seq[j+1]
                        ri = own_pred[ri];
                        cj++;
                    } else if (code == MV_U) {
                        em.step(RG_OP_U, ri, 0);
                        ri = own_pred[ri];
                    } else {
                        em.step(RG_OP_L, ri, 0);
                        cj++;
                    }
                }
                while (cj < L - 1) {
                    em.step(RG_OP_L, ri, 0);
                    cj++;
                }
                if (mode == RG_MODE_REC_GLOBAL)
                    while (ri < n - 1) {
                        em.step(RG_OP_U, ri, 0);
                        ri = own_pred[ri];
                    }
                em.flush(0);
                res.rev_end_row = rev_end;
                res.n_runs_rev = em.n - res.n_runs;
                res.end_row = fen;
                res.end_col = rcol;
            }
        }
        if (tid == 0) {
            if (em.overflow) res.status |= RG_READ_TRACE_OVERFLOW;
            uint32_t nr = em.overflow ? 0 : em.n;
            unsigned long long ro = atomicAdd(&b.counters[1], (unsigned long long)nr);
            if (ro + nr > b.out_run_cap) {
                res.status |= RG_READ_TRACE_OVERFLOW;
                nr = 0;
            }
            for (uint32_t k = 0; k < nr; k++) b.out_runs[ro + k] = runs[k];
            res.run_off = ro;
            if (nr == 0) res.n_runs = res.n_runs_rev = 0;
            if (ws.diag) {   // kilo-cycles: forward pass, reverse pass, pair reduction, replay + walk; materialising rows
                res.fen = (uint32_t)((tc1 - tc0) >> 10);
                res.rsn = (uint32_t)((tc2 - tc1) >> 10);
                res.rec_col = (uint32_t)((tc3 - tc2) >> 10);
                res.rev_end_row = (uint32_t)((clock64() - tc3) >> 10);
                res.displacement = (int32_t)(mcyc >> 10);
            }
            b.results[ridx] = res;
        }
        __syncthreads();
    }
}

size_t pwt_smem_bytes(const PwtWorkspace& ws, bool mx) {
    return (size_t)ws.LP * (ws.stage ? 3 : 11) + (size_t)ws.Pp * 8 + (mx ? (size_t)ws.LT * 8 : 0) + 16;
}

bool pwt_simple(const DevScoring& s) {   // match/mismatch table of score_matrix.rs:35-66
    for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++)
            if (s.sc[a][b] != ((a == b && a < 4) ? s.sc[0][0] : s.sc[0][1])) return false;
    return true;
}
// 32-bit scan keys are exact when every score difference inside a row fits 22 bits (see pwt_pass)
bool pwt_k32(const DevScoring& s, uint32_t LP) {
    long long mx = 1;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) mx = std::max<long long>(mx, std::llabs((long long)s.sc[a][b]));
    return 3 * mx * ((long long)LP + 2) < (1ll << (NT > 256 ? 21 : 22));
}
template <int CPT>
const void* pwt_kernel_c(bool simple, bool k32) {
    if (simple) return k32 ? (const void*)k_pathwise_tr<CPT, true, true> : (const void*)k_pathwise_tr<CPT, true, false>;
    return k32 ? (const void*)k_pathwise_tr<CPT, false, true> : (const void*)k_pathwise_tr<CPT, false, false>;
}
const void* pwt_kernel(uint32_t cpt, const DevScoring& s, uint32_t LP) {
    const bool simple = pwt_simple(s), k32 = pwt_k32(s, LP) && !getenv("RG_PW_K64");
    switch (cpt) {
#if PWT_NT == 256
        case 4: return pwt_kernel_c<4>(simple, k32);
        case 8: return pwt_kernel_c<8>(simple, k32);
        case 16: return pwt_kernel_c<16>(simple, k32);
#endif
        case 32: return pwt_kernel_c<32>(simple, k32);
        default: return nullptr;
    }
}

}  // namespace

int PWT_FN(pathwise_tr_cpt)(uint32_t Lmax) {
#if PWT_NT == 256
    for (int c : {4, 8, 16, 32})
        if ((uint32_t)(NT * c) >= Lmax) return c;
#else
    if ((uint32_t)(NT * 32) >= Lmax) return 32;
#endif
    return 0;
}

int PWT_FN(launch_pathwise_tr)(int mode, const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwtWorkspace& ws,
                       const PoaBatch& b, int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = pwt_smem_bytes(ws, mode != RG_MODE_PATHWISE_GLOBAL);   // every mode but 4 keeps the per-origin maxima
    const void* k = pwt_kernel(ws.CPT, s, ws.LP);
    if (smem > 220 * 1024 || !k) return -3;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    DevPathGraph ga = g, gb = rg_;
    DevScoring sa = s;
    PwtWorkspace wa = ws;
    PoaBatch ba = b;
    int ma = mode;
    void* args[] = {&ga, &gb, &sa, &wa, &ba, &ma};
    if (cudaLaunchKernel(k, dim3(blocks), dim3(NT), args, smem, st) != cudaSuccess) return -1;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int PWT_FN(pathwise_tr_blocks_per_sm)(const DevPathGraph&, const DevPathGraph&, const DevScoring& s, const PwtWorkspace& ws, bool mx, int* nb) {
    const size_t smem = pwt_smem_bytes(ws, mx);
    const void* k = pwt_kernel(ws.CPT, s, ws.LP);
    if (smem > 220 * 1024 || !k) return -3;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, NT, smem) == cudaSuccess ? 0 : -1;
}

}  // namespace rg
