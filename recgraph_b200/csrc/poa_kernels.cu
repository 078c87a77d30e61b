// POA family on sm_100a: one warp per read in flight ("slot"), rows of the linearised graph walked in topological
// order, the read axis spread over the 32 lanes in tiles of 32 columns. Replaces the per-read CPU loops of
// gap_global_abpoa.rs:11-250 (+ helpers 254-455) and the traceback gaf_output.rs:96-253.
//
// Design notes:
//  * integer DP, no tensor cores: the inner loop is IADD/VIMNMX, a 5-step shuffle max-plus scan for the
//    horizontal (x) dependency, one REDUX for the row arg-max that drives the adaptive band of the next row;
//  * the previous row lives in shared memory while the band is narrow (<= WS columns) and in an L2-resident
//    ring otherwise; rows that are predecessors of later segment starts are also written to the ring;
//  * traceback codes are packed (8 bit/cell when in-degree <= 4, else 16 bit) and written coalesced to HBM;
//    the same warp walks them back right after the forward pass and emits run-length step records.
#include <cuda_runtime.h>

#include <cstdio>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

// LIN = true: the scalar linear-gap routine global_abpoa::exec (global_abpoa.rs:260-427; the `-s` retry of mode 0 and
// the non-AVX2 path). It is the same banded recurrence with o = 0, e = gap score and no y matrix (x[c] then equals
// m[c-1] + gap), different fall-backs for unavailable sources (:331-366) and get_max_d_u_l's tie order (utils.rs:129-140).
template <typename TC, int SB, bool LIN>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    k_poa_gap_global(DevGraph g, DevScoring sc, PoaWorkspace ws, PoaBatch b, int WS) {
    // dynamic shared memory: per warp two (m, y) row buffers of WS columns each
    extern __shared__ int32_t s_dyn[];
    __shared__ int32_t s_sc[48];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    int32_t* const s_mw = s_dyn + (size_t)wib * 4 * WS;  // [2][WS] m
    int32_t* const s_yw = s_mw + 2 * WS;                 // [2][WS] y
    const uint32_t slot = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (threadIdx.x < 48) s_sc[threadIdx.x] = (&sc.sc[0][0])[threadIdx.x];
    __syncthreads();
    if (slot >= ws.slots) return;

    const uint32_t n = g.n;
    const uint32_t RM = g.ring - 1;
    RowMeta* rowmeta = ws.rowmeta + (size_t)slot * n;
    int32_t* ring_m = ws.ring_m + (size_t)slot * g.ring * ws.wstride;
    int32_t* ring_y = ws.ring_y + (size_t)slot * g.ring * ws.wstride;
    TC* trace = reinterpret_cast<TC*>(ws.trace) + (size_t)slot * ws.trace_cap;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const int o = LIN ? 0 : sc.o, e = LIN ? sc.sc[0][5] : sc.e;  // LIN: one gap score for every character (checked by the host)
    const int c1 = e + max(o, 0), c2 = o + e;
    constexpr uint32_t SMASK = (1u << SB) - 1;

    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&b.counters[0], 1ull);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= (unsigned long long)b.n_reads) break;
        const int ridx = b.order ? b.order[ticket] : (int)ticket;
        const uint8_t* read = b.reads + b.read_off[ridx];
        const int32_t L = (int32_t)(b.read_off[ridx + 1] - b.read_off[ridx]) + 1;  // includes '$'
        int32_t bta;
        if (sc.fixed_bta >= 0)
            bta = sc.fixed_bta;
        else {
            float v = __fadd_rn(sc.b, __fmul_rn(sc.f, (float)L));  // (b + f * L as f32) as usize, main.rs:175
            bta = !(v > 0.0f) ? 0 : (v >= 536870912.0f ? (1 << 29) : (int32_t)v);
        }
        bta = min(bta, 1 << 29);

        int status = 0;
        uint64_t cells = 0;
        uint64_t off = 0;  // trace cells used so far
        uint32_t prev_left = 0, prev_right = 0, prev_bsp = 0;
        int best_end_val = NEG_INF;  // best last-cell among F predecessors (first max in ascending row order)
        uint32_t best_end_row = 0;
        int last_val = 0;  // last cell of row n-2
        bool abort_read = false;

        for (uint32_t i = 0; i + 1 < n; i++) {
            const uint8_t rf = g.rowflags[i];
            const bool nwp = rf & RF_NWP;
            const uint32_t pb = g.pred_off[i], pe = nwp ? g.pred_off[i + 1] : pb;
            uint32_t ms, me;
            if (i == 0) {
                ms = 0;
                me = 0;
            } else if (!nwp) {
                ms = me = prev_bsp + 1;
            } else {
                uint32_t pl = 0xffffffffu, pr = 0;
                for (uint32_t k = pb; k < pe; k++) {
                    uint32_t p = g.pred_idx[k];
                    uint32_t bs = (p == i - 1) ? prev_bsp : rowmeta[p].bsp;
                    pl = min(pl, bs);
                    pr = max(pr, bs);
                }
                ms = pl + 1;
                me = pr + 1;
            }
            uint32_t left, right;
            band_for_row(ms, me, g.r_values[i], L, bta, left, right);
            const uint32_t W = right - left;
            // `right - left` overflows in the reference when right < left (gap_global_abpoa.rs:59). An EMPTY row (b + f * L < 1
            // only) is legal in mode 2: zero cells, best_scoring_pos = left, none of its cells available to later rows; only the
            // end-cell selection indexes it unconditionally (:205-215). The scalar mode-0 routine reads m[best_p][0] of its
            // smallest predecessor without a band check (global_abpoa.rs:318-321): an empty row stays a panic there.
            if (right < left || (W == 0 && (LIN || (rf & RF_F_PRED) || i == n - 2))) {
                status |= RG_READ_REF_PANIC;
                abort_read = true;
                break;
            }
            if (W == 0) {
                if (lane == 0) {
                    RowMeta rm;
                    rm.base = (int32_t)(off >> 4);
                    rm.left = left;
                    rm.right = right;
                    rm.bsp = left;
                    rowmeta[i] = rm;
                }
                __syncwarp();
                prev_left = left;
                prev_right = right;
                prev_bsp = left;
                continue;
            }
            if (off + W > ws.trace_cap) {
                status |= RG_READ_TRACE_OVERFLOW;
                abort_read = true;
                break;
            }
            cells += W;
            const int li = g.lnz[i];
            const uint32_t best_p = g.min_pred[i];
            const uint32_t mps = g.min_pred_slot[i];
            const int cur = i & 1;
            const bool to_smem = W <= (uint32_t)WS;
            const bool to_ring = (rf & RF_IS_PRED) || !to_smem;
            int32_t* cur_m = s_mw + cur * WS;
            int32_t* cur_y = s_yw + cur * WS;
            int32_t* rg_m = ring_m + (size_t)(i & RM) * ws.wstride;
            int32_t* rg_y = ring_y + (size_t)(i & RM) * ws.wstride;
            const bool prev_in_smem = (prev_right - prev_left) <= (uint32_t)WS;

            int lin_seed0 = 0;  // LIN, left == 0: m[best_p][0] + gap (global_abpoa.rs:318-321; index 0 of the stored row)
            if (LIN && i > 0 && left == 0) {
                const int32_t* mp0;
                if (best_p == i - 1 && prev_in_smem)
                    mp0 = s_mw + (cur ^ 1) * WS;
                else
                    mp0 = ring_m + (size_t)(best_p & RM) * ws.wstride;
                lin_seed0 = mp0[0] + e;
            }
            int carry_u = NEG_INF;   // running prefix max of the x scan (in "minus c1*col" space)
            int carry_h = NEG_INF;   // h of the last column of the previous tile
            unsigned carry_xn = 0;   // "next cell comes from x" flag of the previous tile's last column
            int row_best = NEG_INF;
            uint32_t row_best_col = left;
            int lastcell = 0;

            for (uint32_t t0 = 0; t0 < W; t0 += 32) {
                const uint32_t c = left + t0 + lane;
                const bool act = c < right;
                int mval = NEG_INF, yval = NEG_INF, xval = NEG_INF, hval = NEG_INF;
                uint32_t code = 0;
                int A = NEG_INF;
                bool dav = false;
                int dd = NEG_INF;
                uint32_t dslot = 0, uslot = 0;
                bool yflag = false;
                if (i == 0) {
                    if (act) {
                        yval = (c == 0) ? 0 : o + e * (int)c;
                        mval = yval;
                        code = (c == 0) ? DIR_O : DIR_L;
                    }
                } else {
                    const bool fc = act && c == 0;
                    if (act && !fc) {
                        // ---- vertical (y) and diagonal (d) candidates over the predecessors
                        int u_m = 0, u_y = 0;
                        uint32_t u_m_slot = 0, u_y_slot = 0;
                        bool ufirst = true;
                        const int sco = s_sc[li * 8 + read[c - 1]];
                        const uint32_t np = nwp ? (pe - pb) : 1u;
                        for (uint32_t q = 0; q < np; q++) {
                            const uint32_t p = nwp ? g.pred_idx[pb + q] : i - 1;
                            uint32_t lp, rp;
                            const int32_t *mp, *yp;
                            if (p == i - 1) {
                                lp = prev_left;
                                rp = prev_right;
                                if (prev_in_smem) {
                                    mp = s_mw + (cur ^ 1) * WS;
                                    yp = s_yw + (cur ^ 1) * WS;
                                } else {
                                    mp = ring_m + (size_t)(p & RM) * ws.wstride;
                                    yp = ring_y + (size_t)(p & RM) * ws.wstride;
                                }
                            } else {
                                RowMeta pm = rowmeta[p];
                                lp = pm.left;
                                rp = pm.right;
                                mp = ring_m + (size_t)(p & RM) * ws.wstride;
                                yp = ring_y + (size_t)(p & RM) * ws.wstride;
                            }
                            if (c >= lp && c < rp) {
                                int cum = mp[c - lp] + o, cuy = LIN ? NEG_INF : yp[c - lp];
                                if (ufirst) {
                                    ufirst = false;
                                    u_m = cum;
                                    u_y = cuy;
                                    u_m_slot = q;
                                    u_y_slot = q;
                                }
                                if (cum > u_m) {
                                    u_m = cum;
                                    u_m_slot = q;
                                }
                                if (cuy > u_y) {
                                    u_y = cuy;
                                    u_y_slot = q;
                                }
                            }
                            if (c > lp && c <= rp) {
                                int cd = mp[c - 1 - lp];
                                if (!dav || cd > dd) {  // first predecessor wins ties (gap_global_abpoa.rs:275-283)
                                    dslot = q;
                                    dd = cd;
                                    dav = true;
                                }
                            }
                        }
                        if (ufirst) {  // gap_global_abpoa.rs:132-141 / global_abpoa.rs:343-346
                            yval = LIN ? e * (int)(i + c) : 2 * o + e * (int)(best_p + 1) + e * (int)c;
                            uslot = mps;
                        } else if (u_y > u_m) {
                            yval = u_y + e;
                            uslot = u_y_slot;
                            yflag = true;
                        } else {
                            yval = u_m + e;
                            uslot = u_m_slot;
                        }
                        if (dav) dd += sco;
                        if (LIN && !dav) {  // global_abpoa.rs:355-358: d = gap * (i + left), predecessor = best_p
                            dd = e * (int)(i + left);
                            dslot = mps;
                            dav = true;
                        }
                        hval = dav ? max(dd, yval) : yval;
                    }
                    // ---- horizontal (x): max-plus prefix scan  x[c] = max(x[c-1] + c1, h[c-1] + c2)
                    int hprev = __shfl_up_sync(FULL, hval, 1);
                    if (lane == 0) hprev = carry_h;
                    if (act) {
                        if (c == left) {
                            int seed = (left == 0) ? o + e * (int)(best_p + 1)                           // :88
                                                   : 2 * o + e * (int)(best_p + 1) + e * (int)c;          // :117
                            if (LIN) seed = (left == 0) ? lin_seed0 : e * (int)(i + c);  // global_abpoa.rs:318-321,334-337
                            A = seed - c1 * (int)c;
                        } else {
                            A = (hprev == NEG_INF ? NEG_INF : hprev + c2) - c1 * (int)c;
                        }
                    }
                    int u = warp_incl_max(A, lane);
                    u = max(u, carry_u);
                    carry_u = __shfl_sync(FULL, u, 31);
                    carry_h = __shfl_sync(FULL, hval, 31);
                    if (act) {
                        xval = u + c1 * (int)c;
                        uint32_t dir;
                        if (fc) {
                            mval = xval;
                            yval = 0;  // y[i][0] keeps its initial 0
                            hval = NEG_INF;
                            dir = DIR_U;
                            uslot = mps;
                            yflag = false;
                        } else if (LIN) {  // get_max_d_u_l(d, u, l), utils.rs:129-140
                            if (dd < yval) {
                                if (yval < xval) {
                                    dir = DIR_L;
                                    mval = xval;
                                } else {
                                    dir = DIR_U;
                                    mval = yval;
                                }
                            } else if (dd < xval) {
                                dir = DIR_L;
                                mval = xval;
                            } else {
                                dir = DIR_D;
                                mval = dd;
                            }
                        } else if (dav) {
                            if (dd < xval) {
                                if (xval < yval) {
                                    dir = DIR_U;
                                    mval = yval;
                                    // gap_global_abpoa.rs:153-154: set_path_cell(u_pred, 'u') panics when u_pred == 0
                                    uint32_t up = nwp ? g.pred_idx[pb + uslot] : i - 1;
                                    if (up == 0) status |= RG_READ_REF_PANIC;
                                } else {
                                    dir = DIR_L;
                                    mval = xval;
                                }
                            } else if (dd < yval) {
                                dir = DIR_U;
                                mval = yval;
                            } else {
                                dir = DIR_D;
                                mval = dd;
                            }
                        } else {
                            if (xval < yval) {
                                dir = DIR_U;
                                mval = yval;
                            } else {
                                dir = DIR_L;
                                mval = xval;
                            }
                        }
                        code = dir | (yflag ? 8u : 0u) | ((dslot & SMASK) << 4) | ((uslot & SMASK) << (4 + SB));
                    }
                    // path_x flag of column c: x[c-1] > m[c-1] + o  (gap_global_abpoa.rs:358-364), c > left only
                    unsigned xn = __ballot_sync(FULL, act && xval > mval + o);
                    bool xflag = lane == 0 ? (carry_xn != 0) : ((xn >> (lane - 1)) & 1u);
                    carry_xn = (xn >> 31) & 1u;
                    if (act && c > left && xflag) code |= 4u;
                }
                // ---- row arg-max, right-most (>=)  (gap_global_abpoa.rs:198-203)
                int tmax = __reduce_max_sync(FULL, act ? mval : NEG_INF);
                unsigned eq = __ballot_sync(FULL, act && mval == tmax);
                if (tmax >= row_best) {
                    row_best = tmax;
                    row_best_col = left + t0 + (31 - __clz(eq));
                }
                // ---- stores
                if (act) {
                    trace[off + (c - left)] = (TC)code;
                    if (to_smem) {
                        cur_m[c - left] = mval;
                        cur_y[c - left] = yval;
                    }
                    if (to_ring) {
                        rg_m[c - left] = mval;
                        rg_y[c - left] = yval;
                    }
                }
                if (t0 + 32 >= W) lastcell = __shfl_sync(FULL, mval, (int)(W - 1 - t0));
            }
            if (lane == 0) {
                RowMeta rm;
                rm.base = (int32_t)(off >> 4);   // rows start at multiples of 16 cells: 2^31 x 16 cells of trace per read
                rm.left = left;
                rm.right = right;
                rm.bsp = row_best_col;
                rowmeta[i] = rm;
            }
            __syncwarp();
            off += (W + 15) & ~15ull;
            prev_left = left;
            prev_right = right;
            prev_bsp = row_best_col;
            if ((rf & RF_F_PRED) && lastcell > best_end_val) {
                best_end_val = lastcell;
                best_end_row = i;
            }
            if (i == n - 2) last_val = lastcell;
        }

        status = __reduce_or_sync(FULL, (unsigned)status);
        rg_read_result res;
        res.status = status;
        res.score = 0;
        res.score_f32 = 0.f;
        res.displacement = 0;
        res.end_row = res.end_col = res.start_row = res.start_col = 0;
        res.best_path = res.rev_best_path = 0;
        res.fen = res.rsn = res.rec_col = res.rev_end_row = 0;
        res.cells = cells;
        res.run_off = 0;
        res.n_runs = 0;
        res.n_runs_rev = 0;

        if (!abort_read && !(status & RG_READ_REF_PANIC)) {
            // end cell (gap_global_abpoa.rs:206-214): row n-2 unless an F predecessor is strictly better
            uint32_t last_row = n - 2;
            int best_value = last_val;
            if (best_end_val > last_val) {
                last_row = best_end_row;
                best_value = best_end_val;
            }
            __threadfence_block();
            RowMeta meta = rowmeta[last_row];
            uint32_t row = last_row, col = meta.right - 1;
            res.score = best_value;
            res.end_row = row;
            res.end_col = col;
            // ---- traceback (gaf_output.rs:96-253) fused with band_ampl_enough (gap_global_abpoa.rs:371-455).
            // All lanes walk redundantly (uniform control flow, broadcast loads); lane 0 writes.
            RunEmitter em;
            em.init(runs, ws.run_cap);
            int bandchk = -1;  // -1 undecided, 0 false, 1 true
            bool panic = false;
            for (;;) {
                uint32_t code = trace[((int64_t)meta.base << 4) + ((int64_t)col - (int64_t)meta.left)];
                uint32_t dir = code & 3u;
                if (dir == DIR_O) break;
                if (bandchk < 0) {
                    if (row == 0 || col == 0)
                        bandchk = 1;
                    else if ((col == meta.left && meta.left != 0) || (col == meta.right - 1 && meta.right != (uint32_t)L))
                        bandchk = 0;
                }
                const bool rnwp = g.rowflags[row] & RF_NWP;
                if (dir == DIR_D) {
                    uint32_t p = rnwp ? g.pred_idx[g.pred_off[row] + ((code >> 4) & SMASK)] : row - 1;
                    em.step(g.lnz[row] == read[col - 1] ? RG_OP_D : RG_OP_d, row, lane);
                    row = p;
                    col -= 1;
                    meta = rowmeta[row];
                } else if (dir == DIR_L) {
                    if (code & 4u) {
                        while (code & 4u) {
                            em.step(RG_OP_L, row, lane);
                            if (col <= meta.left) {
                                panic = true;
                                break;
                            }
                            col -= 1;
                            code = trace[((int64_t)meta.base << 4) + ((int64_t)col - (int64_t)meta.left)];
                        }
                    } else {
                        em.step(RG_OP_L, row, lane);
                        if (col <= meta.left)
                            panic = true;
                        else
                            col -= 1;
                    }
                } else {  // DIR_U
                    if (code & 8u) {
                        bool first = true;
                        while (code & 8u) {
                            const bool cn = g.rowflags[row] & RF_NWP;
                            uint32_t p = cn ? g.pred_idx[g.pred_off[row] + ((code >> (4 + SB)) & SMASK)] : row - 1;
                            em.step(first ? RG_OP_U : RG_OP_Y, row, lane);
                            first = false;
                            row = p;
                            meta = rowmeta[row];
                            if (col < meta.left || col >= meta.right) {
                                panic = true;
                                break;
                            }
                            code = trace[((int64_t)meta.base << 4) + ((int64_t)col - (int64_t)meta.left)];
                        }
                    } else {
                        uint32_t p = rnwp ? g.pred_idx[g.pred_off[row] + ((code >> (4 + SB)) & SMASK)] : row - 1;
                        em.step(RG_OP_U, row, lane);
                        row = p;
                        meta = rowmeta[row];
                    }
                }
                if (panic || col < meta.left || col >= meta.right) {
                    panic = true;
                    break;
                }
            }
            em.flush(lane);
            if (panic) res.status |= RG_READ_REF_PANIC;
            if (bandchk == 0) res.status |= RG_READ_BAND_WARNING;
            if (em.overflow) res.status |= RG_READ_TRACE_OVERFLOW;
            res.start_row = row;
            res.start_col = col;
            // ---- publish runs
            uint32_t nr = em.overflow ? 0 : em.n;
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
        }
        if (lane == 0) b.results[ridx] = res;
        __syncwarp();
    }
}

static const void* poa_kernel(int mode, int trace_bytes) {
    if (mode == RG_MODE_GAP_GLOBAL)
        return trace_bytes == 1 ? (const void*)k_poa_gap_global<uint8_t, 2, false> : (const void*)k_poa_gap_global<uint16_t, 6, false>;
    if (mode == RG_MODE_GLOBAL_SCALAR)
        return trace_bytes == 1 ? (const void*)k_poa_gap_global<uint8_t, 2, true> : (const void*)k_poa_gap_global<uint16_t, 6, true>;
    return nullptr;
}

// Shared-memory row width and resident blocks per SM for a batch whose longest read has Lmax columns.
int poa_launch_config(int mode, int trace_bytes, uint32_t Lmax, int* ws_cols, int* blocks_per_sm) {
    const void* k = poa_kernel(mode, trace_bytes);
    if (!k) return -2;
    int ws = 64;
    while (ws < (int)Lmax && ws < 1536) ws += 64;  // 8 warps * 4 rows * 1536 cols * 4 B = 192 KB
    size_t smem = (size_t)WARPS_PER_BLOCK * 4 * ws * sizeof(int32_t);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, WARPS_PER_BLOCK * 32, smem) != cudaSuccess) return -1;
    *ws_cols = ws;
    *blocks_per_sm = nb < 1 ? 1 : nb;
    return 0;
}

int launch_poa(int mode, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
               int trace_bytes, int blocks, int ws_cols, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)WARPS_PER_BLOCK * 4 * ws_cols * sizeof(int32_t);
    if (mode == RG_MODE_GAP_GLOBAL) {
        if (trace_bytes == 1)
            k_poa_gap_global<uint8_t, 2, false><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(g, s, ws, b, ws_cols);
        else
            k_poa_gap_global<uint16_t, 6, false><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(g, s, ws, b, ws_cols);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if (mode == RG_MODE_GLOBAL_SCALAR) {
        if (trace_bytes == 1)
            k_poa_gap_global<uint8_t, 2, true><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(g, s, ws, b, ws_cols);
        else
            k_poa_gap_global<uint16_t, 6, true><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(g, s, ws, b, ws_cols);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    return -2;
}

// ------------------------------------------------------------------------------------------------------------
// INT32 issue-rate microbenchmark: the roofline denominator for the integer DP.
template <int KIND>
__global__ void k_int_peak(int* out, int iters, int seed) {
    int a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    int b0 = seed | 1, b1 = seed + 7;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (KIND == 0) {  // IADD3
                a0 = a0 + b0 + a1;
                a1 = a1 + b1 + a2;
                a2 = a2 + b0 + a3;
                a3 = a3 + b1 + a0;
            } else if (KIND == 1) {  // VIMNMX fed by a LOP3 (a pure max / min chain is folded away by the compiler): 2 ALU-pipe
                                     // instructions per step, both counted below
                a0 = max(a0, a1 ^ b0);
                a1 = max(a1, a2 ^ b1);
                a2 = max(a2, a3 ^ b0);
                a3 = max(a3, a0 ^ b1);
            } else {  // VIADDMNMX (add + max fused)
                a0 = __viaddmax_s32(a0, b0, a1);
                a1 = __viaddmax_s32(a1, b1, a2);
                a2 = __viaddmax_s32(a2, b0, a3);
                a3 = __viaddmax_s32(a3, b1, a0);
            }
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3) == 0x12345678) out[0] = a0;
}

int launch_int_peak(double* iadd, double* imnmx, double* viaddmnmx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int* d_out = nullptr;
    if (cudaMalloc(&d_out, 4) != cudaSuccess) return -1;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double res[3];
    for (int kind = 0; kind < 3; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0, st);
            if (kind == 0)
                k_int_peak<0><<<blocks, threads, 0, st>>>(d_out, iters, rep);
            else if (kind == 1)
                k_int_peak<1><<<blocks, threads, 0, st>>>(d_out, iters, rep);
            else
                k_int_peak<2><<<blocks, threads, 0, st>>>(d_out, iters, rep);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        // instructions per thread: iters * 16 * 4 (kind 1: a LOP3 + a VIMNMX per step, both on the ALU pipe, both counted)
        double instr = (double)blocks * threads * (double)iters * (kind == 1 ? 128.0 : 64.0);
        res[kind] = instr / (best * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *iadd = res[0];
    *imnmx = res[1];
    *viaddmnmx = res[2];
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace rg
