// Warp-cooperative traceback for the affine-gap trace (any layout: TRACE::code(row, column) returns the cell's code
// dir 2 | x 1 | y 1 | diagonal slot SB | vertical slot SB).
//
// Semantics are those of gaf_of_gap_abpoa (gaf_output.rs:96-253) fused with band_ampl_enough
// (gap_global_abpoa.rs:371-455); the output is the run-length step list of include/recgraph_b200.h.
// A traceback is a chain of dependent loads from HBM (the codes were written long ago), ~1 us per step when
// walked cell by cell. Most steps repeat the previous kind of move, so the 32 lanes load the next 32 cells
// ALONG THE CURRENT DIRECTION (vertical, diagonal or horizontal) in one round trip, each lane decides whether
// its cell continues the run, and the longest valid prefix is consumed at once. When the first cell does not
// continue the run it is handled by the exact scalar rules, using the code that was already loaded.
#pragma once
#include "poa_common.cuh"

namespace rg {

enum { WM_U = 0, WM_Y = 1, WM_D = 2, WM_L = 3, WM_X = 4 };

struct WalkOut {
    uint32_t row, col;
    int bandchk;  // -1 undecided, 0 false, 1 true
    bool panic;
};

template <int SB, typename TRACE>
__device__ __forceinline__ WalkOut walk_affine(const TRACE& trace, const RowMeta* __restrict__ rowmeta,
                                               const DevGraph& g, const uint8_t* __restrict__ read, int32_t L,
                                               uint32_t row, uint32_t col, RunEmitter& em, int lane) {
    constexpr unsigned SMASK = (1u << SB) - 1;
    int bandchk = -1;
    bool panic = false;
    int mode = WM_U;
    bool in_chain = false;  // inside a `while path_y == 'Y'` / `while path_x == 'X'` loop of the reference
    for (;;) {
        // ---- every lane looks at the cell `lane` steps further along the current direction
        const int dr = (mode == WM_L || mode == WM_X) ? 0 : 1;
        const int dc = (mode == WM_U || mode == WM_Y) ? 0 : 1;
        const int64_t rr = (int64_t)row - (int64_t)lane * dr;
        const int64_t cc = (int64_t)col - (int64_t)lane * dc;
        const bool valid = rr >= 0 && cc >= 0;
        uint32_t cd = 0;
        RowMeta mt;
        mt.base = 0;
        mt.left = 1;
        mt.right = 0;
        mt.bsp = 0;
        unsigned pslot = PREV_NONE_SLOT;
        bool is_match = false;
        if (valid) {
            cd = trace.code(rr, cc);
            mt = rowmeta[rr];
            pslot = g.prev_slot[rr];
            if (mode == WM_D && cc >= 1) is_match = g.lnz[rr] == read[cc - 1];
        }
        const bool inband = valid && (uint32_t)cc >= mt.left && (uint32_t)cc < mt.right;
        const uint32_t dir = cd & 3u;
        const bool xf = cd & 4u, yf = cd & 8u;
        const unsigned dslot = (cd >> 4) & SMASK, uslot = (cd >> (4 + SB)) & SMASK;
        const bool up_adj = pslot == PREV_ALWAYS_SLOT || uslot == pslot;   // U/Y move leads to row-1
        const bool dg_adj = pslot == PREV_ALWAYS_SLOT || dslot == pslot;   // D move leads to row-1
        // the current cell must lie inside its row's band (the reference would index out of bounds)
        if (!__shfl_sync(FULL, (int)inband, 0)) {
            panic = true;
            break;
        }
        bool cont;
        switch (mode) {
            case WM_U: cont = inband && dir == DIR_U && !yf && up_adj && rr >= 1; break;
            case WM_Y: cont = inband && yf && up_adj && rr >= 1; break;
            case WM_D: cont = inband && dir == DIR_D && dg_adj && rr >= 1 && cc >= 1; break;
            case WM_L: cont = inband && dir == DIR_L && !xf && (uint32_t)cc > mt.left; break;
            default: cont = inband && xf && (uint32_t)cc > mt.left; break;
        }
        // band_ampl_enough is evaluated at the top of every OUTER iteration (not inside the X / Y chains)
        int verdict = -1;
        if (!in_chain) {
            if (rr == 0 || cc == 0)
                verdict = 1;
            else if (((uint32_t)cc == mt.left && mt.left != 0) || ((uint32_t)cc == mt.right - 1 && mt.right != (uint32_t)L))
                verdict = 0;
        }
        // in a chain state the first cell is a chain cell only while its flag is set
        const unsigned cmask = __ballot_sync(FULL, cont);
        int k = __ffs(~cmask) - 1;  // leading lanes that continue the run
        if (k < 0) k = 32;
        if (in_chain) {
            // chain cell without its flag: leave the chain, the cell is re-read as an outer cell (no step taken)
            const bool flag0 = __shfl_sync(FULL, (int)(mode == WM_Y ? yf : xf), 0);
            if (!flag0) {
                in_chain = false;
                // fall through to the scalar outer step below with k = 0
                k = 0;
            }
        }
        if (k > 0) {
            // verdicts of the consumed outer cells, in order
            if (bandchk < 0 && !in_chain) {
                const unsigned dec = __ballot_sync(FULL, verdict >= 0 && lane < k);
                if (dec) {
                    const int src = __ffs(dec) - 1;
                    bandchk = __shfl_sync(FULL, verdict, src);
                }
            }
            if (mode == WM_D) {
                unsigned mm = __ballot_sync(FULL, is_match);
                int done = 0;
                while (done < k) {
                    const unsigned bit = (mm >> done) & 1u;
                    unsigned rest = (bit ? ~mm : mm) >> done;  // first position where the op changes
                    int run = rest ? __ffs(rest) - 1 : 32 - done;
                    run = min(run, k - done);
                    em.bulk(bit ? RG_OP_D : RG_OP_d, row - done, (uint32_t)run, lane);
                    done += run;
                }
                row -= k;
                col -= k;
            } else if (mode == WM_U || mode == WM_Y) {
                em.bulk(mode == WM_U ? RG_OP_U : RG_OP_Y, row, (uint32_t)k, lane);
                row -= k;
            } else {
                em.bulk(RG_OP_L, row, (uint32_t)k, lane);
                col -= k;
            }
            continue;
        }
        // ---- scalar step on the current cell (lane 0's loads), exact reference rules
        const uint32_t cd0 = __shfl_sync(FULL, cd, 0);
        const uint32_t left0 = __shfl_sync(FULL, mt.left, 0), right0 = __shfl_sync(FULL, mt.right, 0);
        const uint32_t dir0 = cd0 & 3u;
        const bool rnwp = g.rowflags[row] & RF_NWP;
        if (in_chain) {
            // flag is set (else we left the chain above) but the run could not continue: take one chain step
            if (mode == WM_Y) {
                uint32_t p = rnwp ? g.pred_idx[g.pred_off[row] + ((cd0 >> (4 + SB)) & SMASK)] : row - 1;
                em.step(RG_OP_Y, row, lane);
                row = p;
            } else {
                em.step(RG_OP_L, row, lane);
                if (col <= left0) {
                    panic = true;
                    break;
                }
                col -= 1;
            }
            continue;
        }
        if (dir0 == DIR_O) break;
        if (bandchk < 0) {
            if (row == 0 || col == 0)
                bandchk = 1;
            else if ((col == left0 && left0 != 0) || (col == right0 - 1 && right0 != (uint32_t)L))
                bandchk = 0;
        }
        if (dir0 == DIR_D) {
            uint32_t p = rnwp ? g.pred_idx[g.pred_off[row] + ((cd0 >> 4) & SMASK)] : row - 1;
            if (col == 0) {
                panic = true;
                break;
            }
            em.step(g.lnz[row] == read[col - 1] ? RG_OP_D : RG_OP_d, row, lane);
            row = p;
            col -= 1;
            mode = WM_D;
        } else if (dir0 == DIR_L) {
            em.step(RG_OP_L, row, lane);
            if (col <= left0) {
                panic = true;
                break;
            }
            col -= 1;
            if (cd0 & 4u) {
                mode = WM_X;
                in_chain = true;
            } else {
                mode = WM_L;
            }
        } else {
            uint32_t p = rnwp ? g.pred_idx[g.pred_off[row] + ((cd0 >> (4 + SB)) & SMASK)] : row - 1;
            em.step(RG_OP_U, row, lane);
            row = p;
            if (cd0 & 8u) {
                mode = WM_Y;
                in_chain = true;
            } else {
                mode = WM_U;
            }
        }
    }
    WalkOut w;
    w.row = row;
    w.col = col;
    w.bandchk = bandchk;
    w.panic = panic;
    return w;
}

}  // namespace rg
