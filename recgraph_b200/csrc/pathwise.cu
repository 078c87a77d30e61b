// Pathwise alignment (modes 4 / 5: pathwise_alignment.rs:5-340, pathwise_alignment_semiglobal.rs:6-277, traceback
// pathwise_alignment_output.rs:7-184) and recombination alignment (modes 8 / 9:
// pathwise_alignment_recombination.rs:23-897, tracebacks recombination_output.rs:12-782) on sm_100a.
//
// The DP is the exact ABSOLUTE-score form of SURVEY §3.4: per row and per incoming edge ("group") the LEADER path
// does a linear-gap DP over the read, every other member path copies the leader's move (D / U / L) applied to its own
// scores. This is what the reference's delta-encoded tensor computes (the parity tests check it against a literal CPU
// restatement). The reverse pass of modes 8/9 (`rev_align`) is the same routine on the reverse graph, rows
// descending, with the read reversed (column jj = L-1-j).
//
// This is the PER-PATH kernel of round 1: every path of every row is computed. Since round 2 the product runs the
// score-transport kernel (pathwise_tr_impl.cuh) for reads of up to 12 287 bases; this one is kept as its A/B reference
// (RG_PW_V1=1, tools/ab_pathwise.py) and for longer reads.
//
// One CTA per read. Scores live in an L2-resident ring of rows, layout [row & RM][path][column] (path-major: a warp
// works on ONE path, its lanes own 8 consecutive columns each and move 128-bit vectors). Per row:
//   phase 1 (per group, all threads over column blocks): leader candidates, CTA-wide max-plus scan for the
//           horizontal dependency, one move byte per column in shared memory;
//   phase 2 (one warp per member path, lanes over columns): the path applies the leader's moves to its own scores
//           (L runs inside a warp are resolved with one ballot + shuffle) and records its OWN arg-max (2 bits per
//           path-cell, two bit planes) — what build_alignment re-derives from the stored scores when it walks a path
//           back. Modes 8/9 also keep, per (row, column), the arg-max over ALL path slots (non-member slots hold 0 as
//           in the reference) for best_alignment, collected with shared-memory atomics.
// best_alignment (pathwise_alignment_recombination.rs:759-873) is an exact pruned reduction: a (column, node) pair
// is expanded over the second node only if its upper bound m + max_w - R can still reach the running maximum;
// the f32 arithmetic uses separately rounded mul / add / sub exactly as the reference, and the sequential
// acceptance rule is restated as (max score; first index with it; first index with it that is on a segment edge).
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

#ifndef RG_PW_THREADS
#define RG_PW_THREADS 512
#endif
constexpr int PT = RG_PW_THREADS;    // threads per CTA
constexpr int MAXPW = 4;   // up to 128 paths (kernels are instantiated for 1..4 words of 32 paths)
enum { MV_D = 1, MV_U = 2, MV_L = 3 };

__device__ __forceinline__ int block_excl_max(int z, int tid, int* s_w) {
    const int lane = tid & 31, w = tid >> 5;
    int inc = warp_incl_max(z, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    int base = NEG_INF;
    for (int k = 0; k < w; k++) base = max(base, s_w[k]);
    int exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = NEG_INF;
    __syncthreads();
    return max(base, exc);
}

#ifdef RG_PWSTATS
__device__ unsigned long long g_pw_cyc[4];
#endif
struct PwSmem {
    unsigned char* mv;  // [max_groups][Lp]
    int32_t* du;        // [Lp]
    int32_t* sj;        // [Lp] substitution score of the current row against every read column
    int32_t* res;       // [Pp]
    uint32_t* end;      // [Pp]
    long long* cb;      // [Lp] modes 8/9: per column max of (score << 8 | path) over all path slots
    long long* rowbest; // [1]  best last-column (score << 8 | 255 - path) among the member paths of the row
    int* w;             // [PT/32]
};

struct PwDir {  // buffers of one DP direction for the read in flight
    int32_t* S;        // ring of rows, [row & RM][path][column]
    int32_t* lead;
    uint32_t* trace;   // [row][path]{plane 0: Lp/8 bytes, plane 1: Lp/8 bytes}, bit = column % 8
    int2* colbest;     // modes 8/9: per (row, column): {max over all slots, path | member << 31}
    int32_t* lastcol;  // forward: [row][Pp] scores of the last column
};

// One DP pass over all rows. Returns nothing; mode-5 best end and mode-4 results are left in shared memory.
template <int TPW>
__device__ void pw_dp(const DevPathGraph& g, const PwDir& d, const PwSmem& sm, const int32_t* s_sc,
                      const uint8_t* read, int L, uint32_t Lp, uint32_t Pp, bool rev, bool free_border, bool track_best,
                      bool track_results, int g_gr, int g_rd, int* s_best_val, int* s_best_set, uint32_t* s_best_row,
                      uint32_t* s_best_path) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = PT / 32;
    const uint32_t n = g.n, P = g.P, PW = g.PW, RM = g.ring - 1;
    const uint32_t LT = Lp / 32;  // column tiles of 32
    const size_t rowsz = (size_t)Pp * Lp;
    const uint32_t base_row = rev ? n - 1 : 0;
    const long long KEY_MIN = -(1ll << 62);
    // per-column (score, path) maxima fit one 32-bit word when |score| < 2^23 for every reachable cell
    int maxabs = 1;
    for (int k = 0; k < 48; k++) maxabs = max(maxabs, abs(s_sc[k]));
    const bool cb32 = (long long)maxabs * ((long long)n + Lp) < (1ll << 23);
    auto rcode = [&](int jj) -> unsigned { return rev ? read[L - 1 - jj] : read[jj - 1]; };
    // ---- base row: every path carries the accumulated read gaps (pathwise_alignment_semiglobal.rs:26-32,
    //      pathwise_alignment_recombination.rs:148-155)
    {
        int32_t* Sb = d.S + (size_t)(base_row & RM) * rowsz;
        for (uint32_t q = warp; q < Pp; q += NWARP)
            for (int j = lane; j < L; j += 32) Sb[(size_t)q * Lp + j] = (q < P) ? j * g_rd : 0;
        int32_t* lb = d.lead + (size_t)(base_row & RM) * Lp;
        for (int j = tid; j < L; j += PT) lb[j] = j * g_rd;
        if (d.lastcol && !rev)
            for (uint32_t q = tid; q < Pp; q += PT) d.lastcol[q] = (q < P) ? (L - 1) * g_rd : 0;
    }
    __syncthreads();
    const int Cc = (L - 1 + PT - 1) / PT;
    const int ntile = (L + 31) / 32;

    for (uint32_t t = 1; t + 1 < n; t++) {
        const uint32_t i = rev ? n - 1 - t : t;
        const uint32_t g0 = g.grp_off[i], g1 = g.grp_off[i + 1];
        const int li = g.lnz[i];
        const int32_t* srow = s_sc + li * 8;
        const uint32_t alpha_i = g.alphas[i];
        int32_t* Si = d.S + (size_t)(i & RM) * rowsz;
        int32_t* lead_i = d.lead + (size_t)(i & RM) * Lp;
#ifdef RG_PWSTATS
        long long pt0 = clock64();
#endif
        // ================= phase 0: per-column data shared by every path of the row =================
        for (int j = tid; j < L; j += PT) sm.sj[j] = (j >= 1) ? srow[rcode(j)] : 0;
        if (d.colbest) {
            // slots of paths that do not go through the row hold 0 (as in the reference): start every column's
            // (score, path) maximum from the highest such slot
            long long init = KEY_MIN;
            for (uint32_t q = 0; q < P; q++)
                if (!((g.node_bits[(size_t)i * PW + q / 32] >> (q % 32)) & 1u)) init = (long long)q;  // (0 << 8) | q
            if (cb32) {
                const int init32 = init == KEY_MIN ? INT_MIN : (int)init;   // (0 << 7) | q
                for (int j = tid; j < L; j += PT) reinterpret_cast<int*>(sm.cb)[j] = init32;
            } else {
                for (int j = tid; j < L; j += PT) sm.cb[j] = init;
            }
        }
        if (d.lastcol && !rev)
            for (uint32_t q = tid; q < Pp; q += PT) d.lastcol[(size_t)i * Pp + q] = 0;
        if (tid == 0) *sm.rowbest = KEY_MIN;
        __syncthreads();
#ifdef RG_PWSTATS
        long long pt1 = clock64();
#endif
        // ================= phase 1: the leader's DP of every group =================
        for (uint32_t gi = g0; gi < g1; gi++) {
            const PwGroup gr = g.grp[gi];
            unsigned char* mv = sm.mv + (size_t)(gi - g0) * Lp;
            const int32_t* lp = gr.lead_is_alpha_of_pred ? d.lead + (size_t)(gr.pred & RM) * Lp : nullptr;
            const int32_t* Sp = d.S + (size_t)(gr.pred & RM) * rowsz + (size_t)gr.leader * Lp;
            const int32_t* src = lp ? lp : Sp;
            const int m0 = free_border ? 0 : src[0] + g_gr;
            const int jb = 1 + tid * Cc, je = min(L, jb + Cc);
            int v = NEG_INF;
            int pl = (jb < L) ? src[jb - 1] : 0;
            for (int j = jb; j < je; j++) {
                const int pc = src[j];
                const int dd = pl + sm.sj[j];
                const int u = pc + g_gr;
                const int du = max(dd, u);
                sm.du[j] = du;
                mv[j] = (dd >= u) ? MV_D : MV_U;  // equality tests in the order d, u (…_semiglobal.rs:46-57)
                int gen = du;
                if (j == 1) gen = max(du, m0 + g_rd);
                v = max(v + g_rd, gen);
                pl = pc;
            }
            const int z = (jb < je && v > NEG_INF / 2) ? v - (je - 1) * g_rd : NEG_INF;
            const int wexc = block_excl_max(z, tid, sm.w);
            int lcand = (wexc > NEG_INF / 2) ? wexc + jb * g_rd : NEG_INF;  // m[jb-1] + g_rd
            const bool own_alpha = gr.leader == alpha_i;
            for (int j = jb; j < je; j++) {
                if (j == 1) lcand = m0 + g_rd;
                const int du = sm.du[j];
                int m = du;
                if (lcand > du) {
                    m = lcand;
                    mv[j] = MV_L;
                }
                if (own_alpha) lead_i[j] = m;
                lcand = m + g_rd;
            }
            if (own_alpha && tid == 0) lead_i[0] = m0;
        }
        __syncthreads();
#ifdef RG_PWSTATS
        long long pt2 = clock64();
#endif
        // ================= phase 2: members apply their leader's move =================
        // One warp per path; lane t of a tile owns CW = 8 consecutive columns (128-bit loads and stores, the chain of an
        // L run inside the lane is a plain serial dependency). A cell whose move is D or U depends only on the
        // predecessor row. Only a lane whose leading cells are L moves needs the value left of its first column: the
        // last value of the nearest lane below that is not made of L moves only (found with a ballot, fetched with a
        // shuffle) plus the gaps in between.
        constexpr int CW = 8;
        const int ntile8 = (L + 32 * CW - 1) / (32 * CW);
        const uint32_t LB = Lp / 8;  // bytes per trace plane of one (row, path)
        for (uint32_t q = warp; q < P; q += NWARP) {
            int gq = -1;
            uint32_t pq = 0;
            for (uint32_t gi = g0; gi < g1; gi++)
                if ((g.grp_mask[(size_t)gi * PW + q / 32] >> (q % 32)) & 1u) {
                    gq = (int)(gi - g0);
                    pq = g.grp[gi].pred;
                }
            if (gq < 0) continue;
            const int32_t* Spq = d.S + (size_t)(pq & RM) * rowsz + (size_t)q * Lp;
            int32_t* Siq = Si + (size_t)q * Lp;
            const unsigned char* mvq = sm.mv + (size_t)gq * Lp;
            unsigned char* trq = reinterpret_cast<unsigned char*>(d.trace) + ((size_t)i * Pp + q) * LB * 2;
            // reverse pass: the 'F' row is never made absolute by the reference (absolute_scores stops before it,
            // pathwise_alignment_recombination.rs:748), so its traceback sees 0 for every path but path 0
            const bool quirk = rev && pq == n - 1 && q != 0;
            const int col0 = free_border ? 0 : Spq[0] + g_gr;
            int carry_last = 0;   // new value of the column left of the tile
            int carry_sp = 0;     // predecessor-row value of the column left of the tile
            int4 nx0 = make_int4(0, 0, 0, 0), nx1 = nx0;
            if ((uint32_t)(lane * CW) < Lp) {
                nx0 = *reinterpret_cast<const int4*>(Spq + lane * CW);
                nx1 = *reinterpret_cast<const int4*>(Spq + lane * CW + 4);
            }
            for (int tile = 0; tile < ntile8; tile++) {
                const int j0 = (tile * 32 + lane) * CW;
                const int sp[CW] = {nx0.x, nx0.y, nx0.z, nx0.w, nx1.x, nx1.y, nx1.z, nx1.w};
                {   // next tile's predecessor values are in flight while this one is computed
                    const int jn = j0 + 32 * CW;
                    if (tile + 1 < ntile8 && (uint32_t)jn < Lp) {
                        nx0 = *reinterpret_cast<const int4*>(Spq + jn);
                        nx1 = *reinterpret_cast<const int4*>(Spq + jn + 4);
                    }
                }
                int spm = __shfl_up_sync(FULL, sp[CW - 1], 1);
                if (lane == 0) spm = carry_sp;
                unsigned mvw[2] = {0x01010101u * MV_D, 0x01010101u * MV_D};
                int sjv[CW] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (j0 < L) {
                    const uint2 mw = *reinterpret_cast<const uint2*>(mvq + j0);
                    mvw[0] = mw.x;
                    mvw[1] = mw.y;
                    const int4 s0 = *reinterpret_cast<const int4*>(sm.sj + j0), s1 = *reinterpret_cast<const int4*>(sm.sj + j0 + 4);
                    sjv[0] = s0.x, sjv[1] = s0.y, sjv[2] = s0.z, sjv[3] = s0.w;
                    sjv[4] = s1.x, sjv[5] = s1.y, sjv[6] = s1.z, sjv[7] = s1.w;
                }
                // values that do not depend on the left neighbour, L flags, position of my first non-L cell
                int nv[CW];
                unsigned lbits = 0;
#pragma unroll
                for (int k = 0; k < CW; k++) {
                    const int j = j0 + k;
                    const unsigned m = (mvw[k / 4] >> (8 * (k % 4))) & 0xffu;
                    const bool isL = j >= 1 && j < L && m == MV_L;
                    if (isL) lbits |= 1u << k;
                    const int spl = (k == 0) ? spm : sp[k - 1];
                    nv[k] = (j == 0) ? col0 : ((m == MV_D) ? spl + sjv[k] : sp[k] + g_gr);
                }
                const int first = __ffs(~lbits) - 1;            // 0..8 (8: every cell of the lane is an L move)
                const bool transparent = first >= CW;
                // chain inside the lane from my first non-L cell on
#pragma unroll
                for (int k = 1; k < CW; k++)
                    if (((lbits >> k) & 1u) && k > first) nv[k] = nv[k - 1] + g_rd;
                // value left of my first column
                const unsigned fixedm = __ballot_sync(FULL, !transparent);
                const unsigned below = fixedm & ((1u << lane) - 1u);
                const int srcl = below ? 31 - __clz(below) : 0;
                const int lastv = __shfl_sync(FULL, nv[CW - 1], srcl);
                const int X = below ? lastv + (lane - 1 - srcl) * CW * g_rd : carry_last + lane * CW * g_rd;
#pragma unroll
                for (int k = 0; k < CW; k++)
                    if (k < first) nv[k] = X + (k + 1) * g_rd;
                // own arg-max in build_alignment's order (d, then u, else l) as the two planes of its code:
                // D = 1, U = 2, L = 3  ->  plane 0 = not U, plane 1 = not D
                unsigned b0 = 0, b1 = 0;
#pragma unroll
                for (int k = 0; k < CW; k++) {
                    const int j = j0 + k;
                    if (j >= 1 && j < L) {
                        const int spl = (k == 0) ? spm : sp[k - 1];
                        const int lq = ((k == 0) ? X : nv[k - 1]) + g_rd;
                        const int dq = (quirk ? 0 : spl) + sjv[k], uq = (quirk ? 0 : sp[k]) + g_gr;
                        const int bq = max(dq, max(uq, lq));
                        const bool isD = bq == dq;
                        if (!isD) b1 |= 1u << k;
                        if (isD || bq != uq) b0 |= 1u << k;
                    }
                }
                if (j0 < L) {
                    *reinterpret_cast<int4*>(Siq + j0) = make_int4(nv[0], nv[1], nv[2], nv[3]);
                    *reinterpret_cast<int4*>(Siq + j0 + 4) = make_int4(nv[4], nv[5], nv[6], nv[7]);
                    trq[j0 / CW] = (unsigned char)b0;
                    trq[LB + j0 / CW] = (unsigned char)b1;
                }
                carry_last = __shfl_sync(FULL, nv[CW - 1], 31);
                carry_sp = __shfl_sync(FULL, sp[CW - 1], 31);
                // max of (score, path) over ALL slots; the highest path id wins ties (…_recombination.rs:809-830)
                if (d.colbest) {
#pragma unroll
                    for (int k = 0; k < CW; k++) {
                        const int j = j0 + k;
                        if (j < L) {
                            if (cb32)
                                atomicMax(reinterpret_cast<int*>(sm.cb) + j, (nv[k] << 7) | (int)q);   // one ATOMS.MAX, no CAS loop
                            else
                                atomicMax(&sm.cb[j], ((long long)nv[k] << 8) | (long long)q);
                        }
                    }
                }
                if (!rev && j0 <= L - 1 && L - 1 < j0 + CW) {
                    int lastval = nv[0];
#pragma unroll
                    for (int k = 1; k < CW; k++)
                        if (j0 + k == L - 1) lastval = nv[k];
                    if (d.lastcol) d.lastcol[(size_t)i * Pp + q] = lastval;
                    // first strict maximum in path order among the member paths
                    atomicMax(sm.rowbest, ((long long)lastval << 8) | (long long)(255 - q));
                    if (track_results)
                        for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                            if (g.grp[fg].pred == i && ((g.grp_mask[(size_t)fg * PW + q / 32] >> (q % 32)) & 1u)) {
                                sm.res[q] = lastval;   // pathwise_alignment.rs:305-319
                                sm.end[q] = i;
                            }
                }
            }
        }
        __syncthreads();
#ifdef RG_PWSTATS
        long long pt3 = clock64();
#endif
        // ================= phase 3: per-row results =================
        if (d.colbest)
            for (int j = tid; j < L; j += PT) {
                uint32_t cb_path;
                int cb_val;
                if (cb32) {
                    const int key = reinterpret_cast<int*>(sm.cb)[j];
                    cb_path = (uint32_t)(key & 0x7f);
                    cb_val = key >> 7;
                } else {
                    const long long key = sm.cb[j];
                    cb_path = (uint32_t)(key & 0xff);
                    cb_val = (int)(key >> 8);
                }
                const bool memb = (g.node_bits[(size_t)i * PW + cb_path / 32] >> (cb_path % 32)) & 1u;
                d.colbest[(size_t)i * Lp + j] = make_int2(cb_val, (int)(cb_path | (memb ? 0x80000000u : 0u)));
            }
        if (track_best && tid == 0 && *sm.rowbest != KEY_MIN) {
            // …_semiglobal.rs:269-273: a row replaces the incumbent only if strictly better
            const long long key = *sm.rowbest;
            const int row_best = (int)(key >> 8);
            if (!*s_best_set || row_best > *s_best_val) {
                *s_best_set = 1;
                *s_best_val = row_best;
                *s_best_row = i;
                *s_best_path = 255u - (uint32_t)(key & 0xff);
            }
        }
        __syncthreads();
#ifdef RG_PWSTATS
        if (tid == 0 && blockIdx.x == 0) {
            const long long pt4 = clock64();
            atomicAdd(&g_pw_cyc[0], (unsigned long long)(pt1 - pt0));
            atomicAdd(&g_pw_cyc[1], (unsigned long long)(pt2 - pt1));
            atomicAdd(&g_pw_cyc[2], (unsigned long long)(pt3 - pt2));
            atomicAdd(&g_pw_cyc[3], (unsigned long long)(pt4 - pt3));
        }
#endif
    }
}

// own-argmax code of path q at (row, col) in one direction's trace
__device__ __forceinline__ unsigned pw_code(const uint32_t* trace, uint32_t Lp, uint32_t PW, uint32_t row, int col, uint32_t q) {
    const uint32_t LB = Lp / 8;
    const unsigned char* t = reinterpret_cast<const unsigned char*>(trace) + ((size_t)row * (PW * 32) + q) * LB * 2;
    const unsigned b0 = t[(uint32_t)col / 8], b1 = t[LB + (uint32_t)col / 8];
    return ((b0 >> (col % 8)) & 1u) | (((b1 >> (col % 8)) & 1u) << 1);
}
__device__ __forceinline__ uint32_t pw_pred(const DevPathGraph& g, uint32_t row, uint32_t q, uint32_t dflt) {
    uint32_t pred = dflt;
    for (uint32_t gi = g.grp_off[row]; gi < g.grp_off[row + 1]; gi++)
        if ((g.grp_mask[(size_t)gi * g.PW + q / 32] >> (q % 32)) & 1u) pred = g.grp[gi].pred;
    return pred;
}

// Forward-direction walk shared by build_alignment (pathwise_alignment_output.rs:7-184), the no_rec builders
// (recombination_output.rs:239-361,633-782) and the forward half of the rec builders (:100-163,472-557).
__device__ void pw_walk_fwd(const DevPathGraph& g, const uint32_t* trace, uint32_t Lp, const uint8_t* read, uint32_t path,
                            uint32_t& ii, int& j, bool pad_global, RunEmitter& em) {
    while (ii > 0 && j > 0) {
        const unsigned code = pw_code(trace, Lp, g.PW, ii, j, path);
        if (code == MV_D) {
            em.step(g.lnz[ii] != read[j - 1] ? RG_OP_d : RG_OP_D, ii, 0);
            ii = pw_pred(g, ii, path, ii - 1);
            j--;
        } else if (code == MV_U) {
            em.step(RG_OP_U, ii, 0);
            ii = pw_pred(g, ii, path, ii - 1);
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
            ii = pw_pred(g, ii, path, g.nwp[ii] ? 0u : ii - 1);
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

template <int TPW>
__global__ void __launch_bounds__(PT, PT <= 256 ? 2 : 1) k_pathwise(DevPathGraph g, DevPathGraph rg_, DevScoring sc, PwWorkspace ws,
                                                  PwRecWorkspace rw, PoaBatch b, int mode) {
    extern __shared__ unsigned char s_dyn[];
    __shared__ int32_t s_sc[48];
    __shared__ int s_w[PT / 32];
    __shared__ unsigned long long s_ticket;
    __shared__ int s_best_val, s_best_set;
    __shared__ uint32_t s_best_row, s_best_path;
    __shared__ float s_redv[PT / 32];
    __shared__ unsigned long long s_redk1[PT / 32], s_redk2[PT / 32];
    __shared__ int s_red_i[PT / 32];
    __shared__ RecBest s_rb;
    __shared__ int s_nsurv;
    __shared__ long long s_rowbest;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slot = blockIdx.x;
    if (tid < 48) s_sc[tid] = (&sc.sc[0][0])[tid];
    const uint32_t n = g.n, P = g.P, PW = g.PW, Lp = ws.Lp, Pp = ws.Pp;
    const uint32_t mg = max(g.max_groups, rg_.max_groups);
    PwSmem sm;
    sm.mv = s_dyn;
    {
        size_t off = (((size_t)mg * Lp + 15) & ~(size_t)15);
        sm.cb = reinterpret_cast<long long*>(s_dyn + off);       // 8-byte aligned (off is a multiple of 16)
        off += (size_t)Lp * 8;
        sm.du = reinterpret_cast<int32_t*>(s_dyn + off);
        sm.sj = sm.du + Lp;
        sm.res = sm.sj + Lp;
        sm.end = reinterpret_cast<uint32_t*>(sm.res + Pp);
    }
    sm.w = s_w;
    sm.rowbest = &s_rowbest;
    uint32_t* s_surv = sm.end + Pp;  // [REC_SURV] survivors of one column (modes 8/9)
    const bool rec_mode = mode == RG_MODE_REC_GLOBAL || mode == RG_MODE_REC_SEMIGLOBAL;
    const bool global_mode = mode == RG_MODE_PATHWISE_GLOBAL || mode == RG_MODE_REC_GLOBAL;
    PwDir fwd, rvd;
    fwd.S = ws.S + (size_t)slot * g.ring * Lp * Pp;
    fwd.lead = ws.lead + (size_t)slot * g.ring * Lp;
    fwd.trace = ws.trace + (size_t)slot * n * Lp * PW * 2;
    fwd.colbest = rec_mode ? rw.fm + (size_t)slot * n * Lp : nullptr;
    fwd.lastcol = rec_mode ? rw.lastcol + (size_t)slot * n * Pp : nullptr;
    if (rec_mode) {
        rvd.S = rw.S + (size_t)slot * rg_.ring * Lp * Pp;
        rvd.lead = rw.lead + (size_t)slot * rg_.ring * Lp;
        rvd.trace = rw.trace + (size_t)slot * n * Lp * PW * 2;
        rvd.colbest = rw.rw + (size_t)slot * n * Lp;
        rvd.lastcol = nullptr;
    }
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const int g_gr = sc.sc[0][5], g_rd = sc.sc[0][5];  // uniform gap score (checked on the host)
    __syncthreads();

    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(&b.counters[0], 1ull);
        __syncthreads();
        const unsigned long long ticket = s_ticket;
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
        if ((uint32_t)L > Lp) {
            res.status = RG_READ_TRACE_OVERFLOW;
            if (tid == 0) b.results[ridx] = res;
            continue;
        }
        if (tid == 0) {
            s_best_set = 0;
            s_best_val = 0;
            s_best_row = 0;
            s_best_path = 0;
        }
        for (uint32_t q = tid; q < Pp; q += PT) {
            sm.res[q] = 0;
            sm.end[q] = 0;
        }
        __syncthreads();
        pw_dp<TPW>(g, fwd, sm, s_sc, read, L, Lp, Pp, false, !global_mode, mode == RG_MODE_PATHWISE_SEMIGLOBAL,
              mode == RG_MODE_PATHWISE_GLOBAL, g_gr, g_rd, &s_best_val, &s_best_set, &s_best_row, &s_best_path);
        if (rec_mode)
            pw_dp<TPW>(rg_, rvd, sm, s_sc, read, L, Lp, Pp, true, !global_mode, false, false, g_gr, g_rd, &s_best_val,
                  &s_best_set, &s_best_row, &s_best_path);

        if (!rec_mode) {
            // ================= modes 4 / 5: end cell, traceback (thread 0), publish =================
            if (tid == 0) {
                uint32_t best_path, ending;
                int score;
                if (global_mode) {
                    // max of (score, path): highest path id wins ties (pathwise_alignment.rs:320-325)
                    best_path = 0;
                    for (uint32_t q = 1; q < P; q++)
                        if (sm.res[q] >= sm.res[best_path]) best_path = q;
                    ending = sm.end[best_path];
                    score = sm.res[best_path];
                } else {
                    best_path = s_best_path;
                    ending = s_best_row;
                    score = s_best_val;
                }
                res.score = score;
                res.best_path = best_path;
                res.end_row = ending;
                res.end_col = (uint32_t)(L - 1);
                RunEmitter em;
                em.init(runs, ws.run_cap);
                uint32_t ii = ending;
                int j = L - 1;
                pw_walk_fwd(g, fwd.trace, Lp, read, best_path, ii, j, global_mode, em);
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
            } else {
                for (uint32_t i = 0; i + 1 < n; i++)
                    for (uint32_t q = 0; q < P; q++)
                        if ((g.node_bits[(size_t)i * PW + q / 32] >> (q % 32)) & 1u) {
                            const int v = fwd.lastcol[(size_t)i * Pp + q];
                            if (!has || mx < v) {
                                has = true;
                                mx = v;
                                bp = q;
                            }
                        }
            }
            s_best_val = mx;
            s_best_path = bp;
            s_rb.v = (float)mx;
            s_rb.k1 = ~0ull;
            s_rb.k2 = ~0ull;
        }
        __syncthreads();
        const int base_score = s_best_val;
        const uint32_t base_path = s_best_path;
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
        for (int j = oob; j < L - oob; j++) {
            const int jj = L - 1 - j;  // column of w in the reverse pass's coordinates
            // upper bounds of this column
            int wmax = NEG_INF;
            for (uint32_t ri = 1 + tid; ri + 1 < n; ri += PT) {
                const int2 e = rvd.colbest[(size_t)ri * Lp + jj];
                if (e.y < 0) wmax = max(wmax, e.x);
            }
            wmax = __reduce_max_sync(FULL, wmax);
            if (lane == 0) s_red_i[warp] = wmax;
            if (tid == 0) s_nsurv = 0;
            __syncthreads();
            wmax = s_red_i[0];
            for (int k = 1; k < PT / 32; k++) wmax = max(wmax, s_red_i[k]);
            const float cur = s_rb.v;
            __syncthreads();
            if (wmax <= NEG_INF / 2) continue;
            // survivors: forward nodes whose best case can still reach the running maximum
            for (uint32_t i = 1 + tid; i + 1 < n; i += PT) {
                const int2 e = fwd.colbest[(size_t)i * Lp + j];
                if (e.y >= 0) continue;  // arg-max slot is not a member path (…_recombination.rs:833)
                const float ub = __fsub_rn((float)(e.x + wmax), Rf);
                if (!prune || ub >= cur) {
                    const int pos = atomicAdd(&s_nsurv, 1);
                    if (pos < REC_SURV) s_surv[pos] = i;
                }
            }
            __syncthreads();
            const int nsurv = s_nsurv;
            if (nsurv > REC_SURV) {
                // too many to stage: every thread walks all forward nodes itself (exact, just slower)
                for (uint32_t i = 1; i + 1 < n; i++) {
                    const int2 fe = fwd.colbest[(size_t)i * Lp + j];
                    if (fe.y >= 0) continue;
                    const uint32_t fp = (uint32_t)fe.y & 0x7fffffffu;
                    const uint32_t seg_i = g.seg[i];
                    const bool iedge = seg_i != g.seg[i + 1];
                    for (uint32_t ri = 1 + tid; ri + 1 < n; ri += PT) {
                        const int2 we = rvd.colbest[(size_t)ri * Lp + jj];
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
                    const int2 fe = fwd.colbest[(size_t)i * Lp + j];
                    const uint32_t fp = (uint32_t)fe.y & 0x7fffffffu;
                    const uint32_t seg_i = g.seg[i];
                    const bool iedge = seg_i != g.seg[i + 1];
                    const int dfs_i = g.dfs[i], dfe_i = g.dfe[i];
                    for (uint32_t ri = 1 + tid; ri + 1 < n; ri += PT) {
                        const int2 we = rvd.colbest[(size_t)ri * Lp + jj];
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
            if (lane == 0) s_redv[warp] = wv;
            __syncthreads();
            if (tid == 0) {
                float m = s_rb.v;
                for (int k = 0; k < PT / 32; k++) m = fmaxf(m, s_redv[k]);
                s_rb.v = m;
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
                s_redv[warp] = r.v;
                s_redk1[warp] = r.k1;
                s_redk2[warp] = r.k2;
            }
            __syncthreads();
            if (tid == 0) {
                RecBest t;
                t.v = s_redv[0];
                t.k1 = s_redk1[0];
                t.k2 = s_redk2[0];
                for (int k = 1; k < PT / 32; k++) {
                    RecBest o;
                    o.v = s_redv[k];
                    o.k1 = s_redk1[k];
                    o.k2 = s_redk2[k];
                    rec_merge2(t, o);
                }
                s_rb = t;
            }
            __syncthreads();
        }
        // 4. outcome + traceback (thread 0)
        if (tid == 0) {
            const RecBest t = s_rb;
            const float basef = (float)base_score;
            // sequential acceptance rule restated: a candidate is taken if it beats the incumbent, or ties it while
            // the incumbent is not on a segment edge and the candidate is (…_recombination.rs:844-851)
            // If the maximum beats the baseline: the first candidate with it wins unless it is off-edge and an on-edge
            // tie follows (k2, the first on-edge tie, equals k1 when k1 itself is on an edge). If the maximum only
            // ties the baseline: the first on-edge candidate with it, if any.
            unsigned long long key = ~0ull;
            if (t.k1 != ~0ull) {
                if (t.v > basef)
                    key = (t.k2 != ~0ull) ? t.k2 : t.k1;
                else if (t.v == basef)
                    key = t.k2;
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
                } else {
                    // ending_node (…_recombination.rs:885-897): first strict maximum over the rows of the path
                    bool has = false;
                    int bs = 0;
                    for (uint32_t i = 1; i + 1 < n; i++)
                        if ((g.node_bits[(size_t)i * PW + base_path / 32] >> (base_path % 32)) & 1u) {
                            const int v = fwd.lastcol[(size_t)i * Pp + base_path];
                            if (!has || v > bs) {
                                has = true;
                                bs = v;
                                ending = i;
                            }
                        }
                }
                res.best_path = res.rev_best_path = base_path;
                res.end_row = ending;
                res.end_col = (uint32_t)(L - 1);
                res.score = fwd.lastcol[(size_t)ending * Pp + base_path];
                res.score_f32 = (float)base_score;
                uint32_t ii = ending;
                int j = L - 1;
                pw_walk_fwd(g, fwd.trace, Lp, read, base_path, ii, j, mode == RG_MODE_REC_GLOBAL, em);
                em.flush(0);
                res.start_row = ii;
                res.n_runs = em.n;
            } else {
                const uint32_t rcol = (uint32_t)(key >> 42), fen = (uint32_t)((key >> 21) & 0x1fffffu), rsn = (uint32_t)(key & 0x1fffffu);
                const uint32_t fp = (uint32_t)fwd.colbest[(size_t)fen * Lp + rcol].y & 0x7fffffffu;
                const uint32_t rp = (uint32_t)rvd.colbest[(size_t)rsn * Lp + (L - 1 - rcol)].y & 0x7fffffffu;
                res.status |= RG_READ_RECOMBINATION;
                res.best_path = fp;
                res.rev_best_path = rp;
                res.fen = fen;
                res.rsn = rsn;
                res.rec_col = rcol;
                res.score_f32 = t.v;
                res.displacement = abs(g.dfs[fen] - g.dfs[rsn]) + abs(g.dfe[fen] - g.dfe[rsn]);
                // forward half, traceback order (recombination_output.rs:100-163 / 472-557)
                uint32_t ii = fen;
                int j = (int)rcol;
                pw_walk_fwd(g, fwd.trace, Lp, read, fp, ii, j, mode == RG_MODE_REC_GLOBAL, em);
                em.flush(0);
                res.start_row = ii;
                res.n_runs = em.n;
                // reverse half, forward order (:38-98 / 389-470): rows ascend
                em.ascending = true;
                uint32_t ri = rsn;
                int cj = (int)rcol;
                uint32_t rev_end = ri;
                while (ri > 0 && ri < n - 1 && cj < L - 1) {
                    const unsigned code = pw_code(rvd.trace, Lp, PW, ri, L - 1 - cj, rp);
                    rev_end = ri;
                    if (code == MV_D) {
                        em.step(g.lnz[ri] != read[cj] ? RG_OP_d : RG_OP_D, ri, 0);  // r_seq[j] = seq[j+1]
                        ri = pw_pred(rg_, ri, rp, ri + 1);
                        cj++;
                    } else if (code == MV_U) {
                        em.step(RG_OP_U, ri, 0);
                        ri = pw_pred(rg_, ri, rp, ri + 1);
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
                        ri = rg_.nwp[ri] ? pw_pred(rg_, ri, rp, ri + 1) : ri + 1;
                    }
                em.flush(0);
                res.rev_end_row = rev_end;
                res.n_runs_rev = em.n - res.n_runs;
                res.end_row = fen;
                res.end_col = rcol;
            }
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
            b.results[ridx] = res;
        }
        __syncthreads();
    }
}

size_t pathwise_smem_bytes(const DevPathGraph& g, const DevPathGraph& rg_, const PwWorkspace& ws, bool rec) {
    const uint32_t mg = rec ? (g.max_groups > rg_.max_groups ? g.max_groups : rg_.max_groups) : g.max_groups;
    return (((size_t)mg * ws.Lp + 15) & ~(size_t)15) + (size_t)ws.Lp * 16 + (size_t)ws.Pp * 8 + (rec ? REC_SURV * 4 : 0);
}

static const void* pw_kernel(uint32_t PW) {
    switch (PW) {
        case 1: return (const void*)k_pathwise<1>;
        case 2: return (const void*)k_pathwise<2>;
        case 3: return (const void*)k_pathwise<3>;
        case 4: return (const void*)k_pathwise<4>;
        default: return nullptr;
    }
}

int launch_pathwise(int mode, const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwWorkspace& ws,
                    const PwRecWorkspace& rw, const PoaBatch& b, int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const bool rec = mode == RG_MODE_REC_GLOBAL || mode == RG_MODE_REC_SEMIGLOBAL;
    const size_t smem = pathwise_smem_bytes(g, rg_, ws, rec);
    const void* k = pw_kernel(g.PW);
    if (smem > 200 * 1024 || !k) return -3;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    switch (g.PW) {
        case 1: k_pathwise<1><<<blocks, PT, smem, st>>>(g, rg_, s, ws, rw, b, mode); break;
        case 2: k_pathwise<2><<<blocks, PT, smem, st>>>(g, rg_, s, ws, rw, b, mode); break;
        case 3: k_pathwise<3><<<blocks, PT, smem, st>>>(g, rg_, s, ws, rw, b, mode); break;
        default: k_pathwise<4><<<blocks, PT, smem, st>>>(g, rg_, s, ws, rw, b, mode); break;
    }
#ifdef RG_PWSTATS
    {
        cudaStreamSynchronize(st);
        unsigned long long h[4];
        cudaMemcpyFromSymbol(h, g_pw_cyc, sizeof h);
        double t = (double)(h[0] + h[1] + h[2] + h[3]);
        fprintf(stderr, "pathwise phases (CTA 0, cumulative): p0 %.1f%% p1 %.1f%% p2 %.1f%% p3 %.1f%%\n", 100 * h[0] / t, 100 * h[1] / t, 100 * h[2] / t, 100 * h[3] / t);
    }
#endif
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int pathwise_blocks_per_sm(const DevPathGraph& g, const DevPathGraph& rg_, const PwWorkspace& ws, bool rec, int* nb) {
    const size_t smem = pathwise_smem_bytes(g, rg_, ws, rec);
    const void* k = pw_kernel(g.PW);
    if (smem > 200 * 1024 || !k) return -3;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, PT, smem) == cudaSuccess ? 0 : -1;
}

}  // namespace rg
