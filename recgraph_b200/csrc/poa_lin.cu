// Modes 0, 1, 3 on sm_100a: global banded POA with the reference's AVX2 semantics (global_abpoa.rs:10-257),
// local POA with AVX2 semantics (local_poa.rs:10-179) and affine-gap local POA (gap_local_poa.rs:8-187), with their
// tracebacks (gaf_output.rs:753-865, 639-751, 502-637).
//
// Same decomposition as the mode-2 kernel: one warp per read, lane t owns the C contiguous columns [tC,(t+1)C),
// the horizontal dependency is a max-plus chain resolved in-lane plus one cross-lane scan per row. These modes keep
// the reference's FULL-matrix semantics: every row holds a value for every column (cells outside the band keep the
// f32 sentinel `min_score` / the local-mode zero), so predecessor rows need no availability logic at all.
// Loops are rolled and the per-lane row lives in local memory (L1): compact code, not the tuned path (mode 2 is).
#include <cuda_runtime.h>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

// trace codes of this family: bits 0-2 kind, bit 3 x flag, bit 4 y flag, then d slot and u slot
enum { K_UNSET = 0, K_D = 1, K_U = 2, K_L = 3, K_STOP = 4 };

// utils.rs:74-98
__device__ __forceinline__ void left_right_x64(uint32_t& left, uint32_t& right, uint32_t seq_len) {
    uint32_t nl = left, nr = right;
    while ((nr - nl) % 8 != 0) {
        if ((nr - nl) % 2 == 0 && nr < seq_len)
            nr += 1;
        else if (nl > 0)
            nl -= 1;
        else
            break;
    }
    if (nl == 0)
        while ((nr - 1) % 8 != 0 && nr < seq_len) nr += 1;
    if (nr == seq_len)
        while ((nr - nl) % 8 != 0 && nl > 1) nl -= 1;
    left = nl;
    right = nr;
}

template <int MODE, int C, typename TC, int SB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    k_poa_lin(DevGraph g, DevScoring sc, PoaWorkspace ws, PoaBatch b) {
    constexpr int STRIDE = 32 * C;
    constexpr unsigned SMASK = (1u << SB) - 1;
    constexpr unsigned SLOT_ROW0 = SMASK;  // mode 3: "predecessor row 0" default of gap_local_poa.rs:131-187
    constexpr int DSH = 5, USH = 5 + SB;
    constexpr bool GLOBAL = MODE == RG_MODE_GLOBAL;
    // RG_MODE_LOCAL_SCALAR (local_poa::exec, local_poa.rs:181-293) is the affine routine with o = 0 and e = the gap score:
    // x[c] = m[c-1] + gap and y = max over predecessors of m + gap (y <= m always, so neither chain flag can be set), same
    // `first = false` restart, same get_max_d_u_l tie order, same clamp and end cell — except that its vertical source is the
    // best predecessor by m alone (get_best_u, local_poa.rs:277-293), never the y arg-max of gap_local_poa.rs:150-187.
    constexpr bool NOY = MODE == RG_MODE_LOCAL_SCALAR;
    constexpr bool AFFINE = MODE == RG_MODE_GAP_LOCAL || NOY;
    __shared__ int32_t s_sc[48];   // [graph][read]
    __shared__ int32_t s_sct[48];  // [read][graph]
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (threadIdx.x < 48) {
        s_sc[threadIdx.x] = (&sc.sc[0][0])[threadIdx.x];
        s_sct[threadIdx.x] = sc.sc[threadIdx.x % 8 < 6 ? threadIdx.x % 8 : 0][threadIdx.x / 8];
    }
    __syncthreads();
    if (slot >= ws.slots) return;

    const uint32_t n = g.n;
    const uint32_t RM = g.ring - 1;
    RowMeta* rowmeta = ws.rowmeta + (size_t)slot * n;
    int32_t* ring_m = ws.ring_m + (size_t)slot * g.ring * STRIDE;
    int32_t* ring_y = ws.ring_y + (size_t)slot * g.ring * STRIDE;
    TC* trace = reinterpret_cast<TC*>(ws.trace) + (size_t)slot * ws.trace_cap;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const int g_gr = sc.sc[0][5];  // score(graph char, '-'), uniform over the alphabet (checked on the host)
    const int o = NOY ? 0 : sc.o, e = NOY ? g_gr : sc.e;
    const int g_rd = sc.sc[0][5];  // score(read char, '-')
    const int cbase = lane * C;

    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&b.counters[0], 1ull);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= (unsigned long long)b.n_reads) break;
        const int ridx = b.order ? b.order[ticket] : (int)ticket;
        const uint8_t* read = b.reads + b.read_off[ridx];
        const int32_t L = (int32_t)(b.read_off[ridx + 1] - b.read_off[ridx]) + 1;
        int32_t bta = 0;
        if (GLOBAL) {
            if (sc.fixed_bta >= 0)
                bta = sc.fixed_bta;
            else {
                float v = __fadd_rn(sc.b, __fmul_rn(sc.f, (float)L));
                bta = !(v > 0.0f) ? 0 : (v >= 536870912.0f ? (1 << 29) : (int32_t)v);
            }
            bta = min(bta, 1 << 29);
        }
        rg_read_result res;
        res.status = 0;
        res.score = 0;
        res.score_f32 = 0.f;
        res.displacement = 0;
        res.end_row = res.end_col = res.start_row = res.start_col = 0;
        res.best_path = res.rev_best_path = 0;
        res.fen = res.rsn = res.rec_col = res.rev_end_row = 0;
        res.cells = 0;
        res.run_off = 0;
        res.n_runs = 0;
        res.n_runs_rev = 0;
        if (L > STRIDE || (uint64_t)n * STRIDE > ws.trace_cap) {
            res.status = RG_READ_TRACE_OVERFLOW;
            if (lane == 0) b.results[ridx] = res;
            continue;
        }
        // global_abpoa.rs:20: min_score = 2 * L * score(read[1], '-')
        const int min_score = GLOBAL ? 2 * L * g_rd : 0;
        // local_poa.rs:22-26
        const int max_multiple = (L % 8 != 0) ? (L / 8) * 8 : L - 8;

        int M[C], Y[C];          // previous row over my columns (full-matrix semantics)
        int GA[C], GD[C], GY[C]; // best predecessor m at c, at c-1, best predecessor y at c
        unsigned GS[C];          // slots: u | d << 8 | y << 16 ; later the trace code
        uint64_t cells = 0;
        uint32_t prev_bsp = 0;
        // end-cell selection
        int best_val = 0;
        bool best_set = false;
        uint32_t best_row = 0, best_col = 0;

        int4 ri_next = reinterpret_cast<const int4*>(g.rowinfo)[0];
        for (uint32_t i = 0; i + 1 < n; i++) {
            const int4 riv = ri_next;
            ri_next = reinterpret_cast<const int4*>(g.rowinfo)[i + 1];
            const uint32_t rbits = (uint32_t)riv.w;
            const uint8_t rf = (rbits >> 8) & 0xffu;
            const bool real_nwp = rf & RF_NWP;                         // decides the reference's code branch
            const bool gather = real_nwp && !(rf & RF_SINGLE_PREV);    // predecessors come from the ring
            const uint32_t pb = (uint32_t)riv.z, np = real_nwp ? (rbits >> 24) : 0u;
            const int li = rbits & 0xffu;
            const unsigned mps = (rbits >> 16) & 0xffu;
            const int best_p = riv.y;

            uint32_t left = 0, right = (uint32_t)L;
            if (GLOBAL) {
                uint32_t ms, me;
                if (i == 0) {
                    ms = me = 0;
                } else if (!real_nwp) {
                    ms = me = prev_bsp + 1;
                } else {
                    uint32_t pl = 0xffffffffu, pr = 0;
                    for (uint32_t k = 0; k < np; k++) {
                        uint32_t p = g.pred_idx[pb + k];
                        uint32_t bs = (p == i - 1) ? prev_bsp : rowmeta[p].bsp;
                        pl = min(pl, bs);
                        pr = max(pr, bs);
                    }
                    ms = pl + 1;
                    me = pr + 1;
                }
                band_for_row(ms, me, riv.x, L, bta, left, right);
                left_right_x64(left, right, (uint32_t)L);
            }
            uint32_t row_bsp = 0;

            if (i == 0) {
#pragma unroll 1
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    int v;
                    unsigned cd;
                    if (GLOBAL) {
                        // global_abpoa.rs:34-35,57-60
                        if (c == 0) {
                            v = 0;
                            cd = K_STOP;
                        } else if (c < (int)right) {
                            v = c * g_rd;
                            cd = K_L;
                        } else {
                            v = min_score;
                            cd = K_UNSET;
                        }
                    } else {
                        v = 0;
                        cd = K_STOP;
                    }
                    M[k] = v;
                    Y[k] = 0;
                    GS[k] = cd;
                }
                if (GLOBAL) cells += right - left;
            } else {
                // ---- predecessor maxima at column c (GA), c-1 (GD) and, affine, y at c (GY)
                if (!gather) {
                    int up = __shfl_up_sync(FULL, M[C - 1], 1);
                    if (lane == 0) up = GLOBAL ? min_score : 0;  // column -1 never read
#pragma unroll 1
                    for (int k = C - 1; k >= 0; k--) {
                        GA[k] = M[k];
                        GD[k] = (k == 0) ? up : M[k - 1];
                        GY[k] = Y[k];
                        GS[k] = 0;
                    }
                    if (AFFINE && real_nwp) {
                        // gap_local_poa.rs:131-187 with `first == false`: running maxima start at (0, row 0)
#pragma unroll 1
                        for (int k = 0; k < C; k++) {
                            unsigned s = 0;
                            if (!(GA[k] + o > 0)) s |= SLOT_ROW0;         // u_m stays 0 @ idx 0
                            if (!(GD[k] > 0)) s |= SLOT_ROW0 << 8;
                            if (!(GY[k] > 0)) s |= SLOT_ROW0 << 16;
                            GS[k] = s;
                        }
                    }
                } else {
#pragma unroll 1
                    for (int k = 0; k < C; k++) {
                        GA[k] = NEG_INF;
                        GD[k] = NEG_INF;
                        GY[k] = NEG_INF;
                        GS[k] = AFFINE ? (SLOT_ROW0 | (SLOT_ROW0 << 8) | (SLOT_ROW0 << 16)) : 0u;
                    }
                    for (uint32_t q = 0; q < np; q++) {
                        const uint32_t p = g.pred_idx[pb + q];
                        const int32_t* mp = ring_m + (size_t)(p & RM) * STRIDE + cbase;
                        const int32_t* yp = ring_y + (size_t)(p & RM) * STRIDE + cbase;
                        int carry = (lane == 0) ? (GLOBAL ? min_score : 0) : mp[-1];
#pragma unroll 1
                        for (int k = 0; k < C; k++) {
                            const int mv = mp[k];
                            unsigned s = GS[k];
                            if (AFFINE) {
                                // strict > against a running maximum that starts at 0 (in "m + o" space for u_m)
                                const int ga = GA[k] == NEG_INF ? -o : GA[k];  // running max of m such that m + o > 0
                                if (mv + o > max(ga + o, 0)) {
                                    GA[k] = mv;
                                    s = (s & ~0xffu) | q;
                                }
                                if (carry > max(GD[k], 0)) {
                                    GD[k] = carry;
                                    s = (s & ~0xff00u) | (q << 8);
                                }
                                const int yv = yp[k];
                                if (yv > max(GY[k], 0)) {
                                    GY[k] = yv;
                                    s = (s & ~0xff0000u) | (q << 16);
                                }
                            } else {
                                // first predecessor initialises, later ones need strictly more (global_abpoa.rs:122-138)
                                if (q == 0 || mv > GA[k]) {
                                    GA[k] = mv;
                                    s = (s & ~0xffu) | q;
                                }
                                if (q == 0 || carry > GD[k]) {
                                    GD[k] = carry;
                                    s = (s & ~0xff00u) | (q << 8);
                                }
                            }
                            GS[k] = s;
                            carry = mv;
                        }
                    }
                }
                const int32_t* srow = s_sc + li * 8;    // sc[lnz][read]
                const int32_t* srowt = s_sct + li * 8;  // sc[read][lnz]
                int start = 1, end = L, rgt = L;
                int col0 = 0;
                if (GLOBAL) {
                    start = left == 0 ? 1 : (int)left;
                    rgt = (int)right;
                    end = (rgt == L) ? ((rgt - start) / 8) * 8 + start : rgt;
                    // global_abpoa.rs:36-46: column 0 of every row, through the numerically smallest predecessor
                    int pc0;
                    if (!gather)
                        pc0 = __shfl_sync(FULL, M[0], 0);
                    else
                        pc0 = ring_m[(size_t)((uint32_t)best_p & RM) * STRIDE];
                    col0 = pc0 + g_gr;
                    cells += right - left;
                } else {
                    end = max_multiple + 1;  // SIMD blocks cover [1, max_multiple]
                }
                // ---- pass A/B: candidates and in-lane chain  v[c] = max(a[c], v[c-1] + g)
                // chain value entering column `start` (global: m[i][start-1]; local: column 0 == 0)
                const int seed = GLOBAL ? ((start - 1 == 0) ? col0 : min_score) : 0;
                int hl;  // affine: hh of my last column for the next lane's first generator
                {
                    int xl = NEG_INF;
                    int hprev_in = 0;
                    if (AFFINE) {
                        // hh of the previous lane's last column must be known first
                        const int k = C - 1, c = cbase + k;
                        const unsigned rc = (c >= 1 && c < L) ? read[c - 1] : 4u;
                        int d, u;
                        if (!real_nwp) {
                            d = GD[k] + srowt[rc];
                            u = max(GY[k] + e, GA[k] + o + e);
                        } else {
                            d = max(GD[k], 0) + srowt[rc];
                            const int um = max(GA[k] == NEG_INF ? 0 : GA[k] + o, 0), uy = max(GY[k], 0);
                            u = max(um, uy) + e;
                        }
                        hl = max(0, max(d, u));
                        if (c < 1 || c >= L) hl = 0;
                        hprev_in = __shfl_up_sync(FULL, hl, 1);
                        if (lane == 0) hprev_in = 0;
                    }
                    int hp = hprev_in;
#pragma unroll 1
                    for (int k = 0; k < C; k++) {
                        const int c = cbase + k;
                        const bool incol = c >= start && c < rgt;
                        const unsigned rc = (c >= 1 && c < L) ? read[c - 1] : 4u;
                        const unsigned s = GS[k];
                        if (!AFFINE) {
                            const bool tail = c >= end;
                            const int us = GA[k] + g_gr;
                            const int ds = GD[k] + ((tail && real_nwp) ? srowt[rc] : srow[rc]);
                            // SIMD blocks: D only if strictly better (global_abpoa.rs:107); scalar tail: D on ties (:175)
                            const bool pickd = tail ? (ds >= us) : (ds > us);
                            int du = pickd ? ds : us;
                            unsigned cd = pickd ? (K_D | (((s >> 8) & SMASK) << DSH)) : (K_U | ((s & SMASK) << USH));
                            int a = du;
                            if (!GLOBAL) {
                                // local: clamp folded into the chain (local_poa.rs:100-103,116-118); none in the
                                // multi-predecessor scalar tail (:126-163)
                                if (!(tail && real_nwp)) a = max(du, 0);
                            }
                            if (!incol) a = NEG_INF;
                            GA[k] = du;   // kept for the direction decision
                            GD[k] = a;
                            GS[k] = cd;
                            int gen = a;
                            if (c == start) gen = max(a, seed + g_rd);
                            xl = max(xl + g_rd, gen);
                            GY[k] = xl;  // in-lane chain value
                        } else {
                            int d, u;
                            unsigned cd = 0;
                            if (!real_nwp) {
                                d = GD[k] + srowt[rc];
                                const int uy = GY[k] + e, um = GA[k] + o + e;
                                if (!NOY && uy > um) cd |= 16u;
                                u = max(uy, um);
                            } else {
                                d = max(GD[k], 0) + srowt[rc];
                                const int um = max(GA[k] == NEG_INF ? 0 : GA[k] + o, 0), uy = max(GY[k], 0);
                                unsigned usl;
                                if (NOY || um > uy) {
                                    u = um + e;
                                    usl = s & 0xffu;
                                } else {
                                    u = uy + e;
                                    usl = (s >> 16) & 0xffu;
                                    cd |= 16u;
                                }
                                cd |= ((usl & SMASK) << USH) | ((((s >> 8) & 0xffu) & SMASK) << DSH);
                            }
                            GA[k] = d;
                            GD[k] = u;
                            GS[k] = cd;
                            // x[c] = max(x[c-1] + c1, hh[c-1] + c2), x[0] = 0, hh[0] = 0
                            const int c1 = e + max(o, 0), c2 = o + e;
                            int gen = (c >= 1 && c < L) ? hp + c2 : NEG_INF;
                            if (c == 1) gen = max(gen, 0 + c1);
                            xl = max(xl + c1, gen);
                            GY[k] = xl;
                            hp = (c >= 1 && c < L) ? max(0, max(d, u)) : 0;
                        }
                    }
                }
                // ---- cross-lane scan of the chain
                const int cstep = AFFINE ? (e + max(o, 0)) : g_rd;
                int z = GY[C - 1] - (cbase + C - 1) * cstep;
                if (GY[C - 1] <= NEG_INF / 2) z = NEG_INF;
                const int winc = warp_incl_max(z, lane);
                int wexc = __shfl_up_sync(FULL, winc, 1);
                if (lane == 0) wexc = NEG_INF;
                const int xin0 = (wexc > NEG_INF / 2) ? wexc + cbase * cstep : NEG_INF;
                // ---- pass C
                int lane_best = NEG_INF;
                int lane_bcol = -1;
                unsigned xn_bits = 0;
                int lv = xin0;  // non-affine: the "left" candidate m[c-1] + g of the current cell
#pragma unroll 1
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    int v = GY[k];
                    if (xin0 > NEG_INF / 2) v = max(v, xin0 + k * cstep);
                    unsigned cd = GS[k];
                    if (!AFFINE) {
                        const bool incol = c >= start && c < rgt;
                        const int du = GA[k];
                        const bool tail = c >= end;
                        int m;
                        if (incol) {
                            if (c == start) lv = seed + g_rd;
                            // SIMD blocks (global_abpoa.rs:156-161) and scalar tail (:174-181) agree: L iff l > max(d,u)
                            m = v;
                            if (lv > du) cd = K_L;
                            if (!GLOBAL && !(tail && real_nwp)) {
                                const int raw = max(du, lv);
                                // SIMD blocks stop at <= 0 (local_poa.rs:100-103), the scalar tail at < 0 (:116-118)
                                if (tail ? (raw < 0) : (raw <= 0)) {
                                    m = 0;
                                    cd = K_STOP;
                                }
                            }
                        } else if (GLOBAL) {
                            m = (c == 0) ? col0 : min_score;
                            cd = (c == 0) ? (K_U | ((mps & SMASK) << USH)) : (unsigned)K_UNSET;
                        } else {
                            m = 0;
                            cd = K_STOP;
                        }
                        lv = m + g_rd;
                        M[k] = m;
                        GS[k] = cd;
                        if (GLOBAL) {
                            // best_col: right-most maximum over [left, right) (global_abpoa.rs:79,162-164,220-222)
                            if (c >= (int)left && c < rgt && m >= lane_best) {
                                lane_best = m;
                                lane_bcol = c;
                            }
                        } else if (c >= 1 && c < L && m >= lane_best) {
                            lane_best = m;
                            lane_bcol = c;
                        }
                    } else {
                        const bool incol = c >= 1 && c < L;
                        const int d = GA[k], u = GD[k], l = v;
                        int m = 0;
                        if (incol) {
                            if (d < 0 && l < 0 && u < 0) {
                                m = 0;
                                cd |= K_STOP;  // path = 'O'; path_y keeps its flag and predecessor
                            } else if (d < u) {  // utils.rs:129-140
                                if (u < l) {
                                    m = l;
                                    cd |= K_L;
                                } else {
                                    m = u;
                                    cd |= K_U;
                                }
                            } else if (d < l) {
                                m = l;
                                cd |= K_L;
                            } else {
                                m = d;
                                cd |= K_D;
                            }
                            if (l > m + o) xn_bits |= 1u << k;  // X flag of column c+1: x[c] > m[c] + o
                            if (m > lane_best) {  // first strict maximum (gap_local_poa.rs:114-117)
                                lane_best = m;
                                lane_bcol = c;
                            }
                        } else {
                            cd = K_STOP;
                            if (c == 0 && 0 > o) xn_bits |= 1u;  // x[0] = 0, m[0] = 0
                        }
                        M[k] = m;
                        Y[k] = incol ? u : 0;
                        GS[k] = cd;
                    }
                }
                if (AFFINE) {
                    unsigned prev_last = __shfl_up_sync(FULL, xn_bits >> (C - 1), 1) & 1u;
                    if (lane == 0) prev_last = 0;
                    const unsigned xf = (xn_bits << 1) | prev_last;
#pragma unroll 1
                    for (int k = 0; k < C; k++) {
                        const int c = cbase + k;
                        if (c >= 1 && c < L && ((xf >> k) & 1u)) GS[k] |= 8u;
                    }
                }
                // ---- row reductions
                if (GLOBAL) {
                    const int tmax = __reduce_max_sync(FULL, lane_best);
                    const unsigned eq = __ballot_sync(FULL, lane_best == tmax && lane_bcol >= 0);
                    // a row whose band holds column 0 only processes no cell: best_col keeps its initial value `left`
                    // (global_abpoa.rs:80,160-162,222) — with a band amplitude of 0 that is every row
                    row_bsp = eq ? (uint32_t)__shfl_sync(FULL, lane_bcol, 31 - __clz(eq)) : left;
                } else {
                    const int tmax = __reduce_max_sync(FULL, lane_best);
                    const unsigned eq = __ballot_sync(FULL, lane_best == tmax && lane_bcol >= 0);
                    if (eq) {
                        // mode 1: last maximum wins (>=, local_poa.rs:104,164); mode 3: first strict maximum
                        const int src = AFFINE ? (__ffs(eq) - 1) : (31 - __clz(eq));
                        const int bc = __shfl_sync(FULL, lane_bcol, src);
                        const bool take = AFFINE ? (tmax > best_val) : (tmax >= best_val);
                        if (take) {
                            best_val = tmax;
                            best_row = i;
                            best_col = (uint32_t)bc;
                        }
                    }
                }
            }
            // ---- stores
            {
                unsigned code[C];
#pragma unroll
                for (int k = 0; k < C; k++) code[k] = GS[k];
                TC* dst = trace + (size_t)i * STRIDE + cbase;
#pragma unroll
                for (int k = 0; k < C; k++) dst[k] = (TC)code[k];
            }
            if (rf & RF_IS_PRED) {
                int32_t* dm = ring_m + (size_t)(i & RM) * STRIDE + cbase;
                int32_t* dy = ring_y + (size_t)(i & RM) * STRIDE + cbase;
#pragma unroll 1
                for (int k = 0; k < C; k++) {
                    dm[k] = M[k];
                    if (AFFINE) dy[k] = Y[k];
                }
            }
            if (GLOBAL) {
                if (lane == 0) {
                    RowMeta rm;
                    rm.base = 0;
                    rm.left = left;
                    rm.right = right;
                    rm.bsp = row_bsp;
                    rowmeta[i] = rm;
                }
                prev_bsp = row_bsp;
                if (rf & RF_F_PRED) {
                    // global_abpoa.rs:227-240: first strict maximum over F's predecessors, column L-1
                    int lv = NEG_INF;
#pragma unroll 1
                    for (int k = 0; k < C; k++)
                        if (cbase + k == L - 1) lv = M[k];
                    const int lastcell = __reduce_max_sync(FULL, lv);
                    if (!best_set || lastcell > best_val) {
                        best_set = true;
                        best_val = lastcell;
                        best_row = i;
                        best_col = (uint32_t)(L - 1);
                    }
                }
            }
            __syncwarp();
        }

        res.cells = GLOBAL ? cells : (uint64_t)(n - 2) * (uint64_t)(L - 1);
        res.score = best_val;
        res.score_f32 = (float)best_val;
        res.end_row = best_row;
        res.end_col = best_col;
        // ---- traceback (scalar walk, uniform over the warp; lane 0 writes the runs)
        uint32_t row = best_row, col = best_col;
        RunEmitter em;
        em.init(runs, ws.run_cap);
        bool not_enough = false;
        for (;;) {
            const uint32_t cd = trace[(size_t)row * STRIDE + col];
            const uint32_t kind = cd & 7u;
            if (kind == K_STOP) break;
            if (kind == K_UNSET) {  // gaf_output.rs:779-782: a -1 cell
                not_enough = true;
                break;
            }
            const bool rnwp = g.rowflags[row] & RF_NWP;
            if (kind == K_D) {
                const unsigned sl = (cd >> DSH) & SMASK;
                const uint32_t p = (AFFINE && rnwp && sl == SLOT_ROW0) ? 0u : (rnwp ? g.pred_idx[g.pred_off[row] + sl] : row - 1);
                uint32_t op;
                if (GLOBAL) {
                    // gaf_output.rs:790-798: D/d decided AFTER the move (lnz[new row] vs seq[new col])
                    const uint32_t nc = col - 1;
                    const int a = (p == 0) ? 6 : g.lnz[p];
                    const int bq = (nc == 0) ? 6 : read[nc - 1];
                    op = (a == bq) ? RG_OP_D : RG_OP_d;
                } else if (AFFINE) {
                    op = (g.lnz[row] == read[col - 1]) ? RG_OP_D : RG_OP_d;
                } else {
                    op = RG_OP_D;
                }
                em.step(op, row, lane);
                row = p;
                col -= 1;
            } else if (kind == K_L) {
                if (AFFINE && (cd & 8u)) {
                    uint32_t c2 = cd;
                    while (c2 & 8u) {  // gaf_output.rs:570-573
                        em.step(RG_OP_L, row, lane);
                        col -= 1;
                        c2 = trace[(size_t)row * STRIDE + col];
                    }
                } else {
                    em.step(RG_OP_L, row, lane);
                    col -= 1;
                }
            } else {  // K_U
                if (AFFINE && (cd & 16u)) {
                    uint32_t c2 = cd;
                    bool first = true;
                    while (c2 & 16u) {  // gaf_output.rs:581-587
                        const bool cn = g.rowflags[row] & RF_NWP;
                        const unsigned sl = (c2 >> USH) & SMASK;
                        const uint32_t p = (cn && sl == SLOT_ROW0) ? 0u : (cn ? g.pred_idx[g.pred_off[row] + sl] : row - 1);
                        em.step(first ? RG_OP_U : RG_OP_Y, row, lane);
                        first = false;
                        row = p;
                        c2 = trace[(size_t)row * STRIDE + col];
                    }
                } else {
                    const unsigned sl = (cd >> USH) & SMASK;
                    const uint32_t p = (AFFINE && rnwp && sl == SLOT_ROW0) ? 0u : (rnwp ? g.pred_idx[g.pred_off[row] + sl] : row - 1);
                    em.step(RG_OP_U, row, lane);
                    row = p;
                }
            }
        }
        em.flush(lane);
        if (not_enough) res.status |= RG_READ_BAND_NOT_ENOUGH;
        if (em.overflow) res.status |= RG_READ_TRACE_OVERFLOW;
        res.start_row = row;
        res.start_col = col;
        uint32_t nr = (em.overflow || not_enough) ? 0 : em.n;
        unsigned long long ro = 0;
        if (lane == 0) ro = atomicAdd(&b.counters[1], (unsigned long long)nr);
        ro = __shfl_sync(FULL, ro, 0);
        if (ro + nr > b.out_run_cap) {
            res.status |= RG_READ_TRACE_OVERFLOW;
            nr = 0;
        }
        __syncwarp();
        for (uint32_t k = lane; k < nr; k += 32) b.out_runs[ro + k] = runs[k];
        res.run_off = ro;
        res.n_runs = nr;
        if (lane == 0) b.results[ridx] = res;
        __syncwarp();
    }
}

template <int MODE, int C>
static int launch_mc(const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b, int trace_bytes,
                     int blocks, cudaStream_t st) {
    if (trace_bytes == 1)
        k_poa_lin<MODE, C, uint8_t, 1><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(g, s, ws, b);
    else
        k_poa_lin<MODE, C, uint16_t, 5><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(g, s, ws, b);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <int MODE>
static int launch_m(int C, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
                    int trace_bytes, int blocks, cudaStream_t st) {
    switch (C) {
        case 4: return launch_mc<MODE, 4>(g, s, ws, b, trace_bytes, blocks, st);
        case 8: return launch_mc<MODE, 8>(g, s, ws, b, trace_bytes, blocks, st);
        case 16: return launch_mc<MODE, 16>(g, s, ws, b, trace_bytes, blocks, st);
        case 32: return launch_mc<MODE, 32>(g, s, ws, b, trace_bytes, blocks, st);
        default: return -2;
    }
}
int launch_poa_lin(int mode, int C, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
                   int trace_bytes, int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
        case RG_MODE_GLOBAL: return launch_m<RG_MODE_GLOBAL>(C, g, s, ws, b, trace_bytes, blocks, st);
        case RG_MODE_LOCAL: return launch_m<RG_MODE_LOCAL>(C, g, s, ws, b, trace_bytes, blocks, st);
        case RG_MODE_GAP_LOCAL: return launch_m<RG_MODE_GAP_LOCAL>(C, g, s, ws, b, trace_bytes, blocks, st);
        case RG_MODE_LOCAL_SCALAR: return launch_m<RG_MODE_LOCAL_SCALAR>(C, g, s, ws, b, trace_bytes, blocks, st);
        default: return -2;
    }
}
// trace cell bytes for this family: kind 3 | x 1 | y 1 | 2 slots
int poa_lin_trace_bytes(uint32_t max_indeg) { return max_indeg <= 1 ? 1 : 2; }

}  // namespace rg
