// Mode 2 (gap_global_abpoa.rs:11-250), register-blocked variant for reads of up to 32*C columns.
//
// One warp per read in flight. Lane t OWNS the C contiguous DP columns [t*C, (t+1)*C) for the whole read: the
// previous row's m / y values of those columns stay in registers from row to row (no shared or global traffic on
// rows inside a segment), every per-cell step is fully unrolled straight-line integer code, and the horizontal
// affine dependency x[c] = max(x[c-1] + c1, h[c-1] + c2) is resolved by an in-lane chain plus ONE cross-lane
// max-plus scan per row. Rows that are predecessors of later segment starts are written to an L2-resident ring
// (absolute column index, NEG_INF outside the band) and gathered from there on segment-start rows.
// Traceback codes are packed and written with one or two 128-bit stores per lane and row, at the fixed address
// row * 32C + column, so the traceback needs no per-row offset lookup.
#include <cuda_runtime.h>

#include <type_traits>

#include "device.h"
#include "poa_common.cuh"
#include "poa_walk.cuh"

namespace rg {

#define NEGH (NEG_INF / 2)  // anything below is "no value"

// ---- trace layout of this kernel: four bit planes per row (instead of one code byte per cell)
//   plane 0 (T): the cell's m is max(d, x), i.e. the move is not vertical     plane 2 (X): path_x flag (x extends)
//   plane 1 (D): d >= x, i.e. diagonal rather than horizontal when T is set   plane 3 (Y): path_y flag (y extends)
// Lane t owns the 4*C bits of its C columns: bit p*C + PlaneFmt<C>::perm(k) = plane p of column t*C + k, stored in
// NW = max(1, C/8) words at word offset row * 32*NW + t*NW. The bit order inside a plane is the one the packed row
// produces for free: the flag words of two consecutive cell pairs go through ONE byte permute with sign replication
// (a byte of 0xff / 0x00 per cell) and one AND-OR, so bit b of byte j collects pair 2b + (j & 1), half j >> 1. Rows that gather several predecessors additionally keep 2*SB slot planes
// (diagonal source, vertical source) in a side array indexed by DevGraph::nwp_ord.
template <int C>
struct PlaneFmt {
    static constexpr int NW = (C >= 8) ? C / 8 : 1;
    static constexpr int ROWW = 32 * NW;
    static constexpr int H = C / 2, G = C / 4;   // cells per half, bits per group (4 groups: lo-even, lo-odd, hi-even, hi-odd)
    // plane bit of column k of a lane's block
    __host__ __device__ static constexpr unsigned perm(unsigned k) {
        const unsigned half = k >= (unsigned)H ? 1u : 0u, kk = k - half * H;
        return (2u * half + (kk & 1u)) * G + (kk >> 1);
    }
};

template <int C>
__device__ __forceinline__ void store_planes(uint32_t* dst, const unsigned (&P)[4]) {
    if constexpr (C == 32) {
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(P[0], P[1], P[2], P[3]);
    } else if constexpr (C == 16) {
        reinterpret_cast<uint2*>(dst)[0] = make_uint2(P[0] | (P[1] << 16), P[2] | (P[3] << 16));
    } else {
        dst[0] = P[0] | (P[1] << C) | (P[2] << (2 * C)) | (P[3] << (3 * C));
    }
}

// code byte layout of the 32-bit producers (dir 2 | x 1 | y 1 | dslot SB | uslot SB) -> planes
template <int C, int SB>
__device__ __forceinline__ void planes_from_codes(const unsigned (&code)[C], unsigned (&P)[4], unsigned (&SD)[SB], unsigned (&SU)[SB]) {
    P[0] = P[1] = P[2] = P[3] = 0u;
#pragma unroll
    for (int q = 0; q < SB; q++) SD[q] = SU[q] = 0u;
#pragma unroll
    for (int k = 0; k < C; k++) {
        const unsigned cd = code[k], dir = cd & 3u;
        const unsigned pk = PlaneFmt<C>::perm((unsigned)k);
        P[0] |= (dir != (unsigned)DIR_U ? 1u : 0u) << pk;
        P[1] |= (dir == (unsigned)DIR_D ? 1u : 0u) << pk;
        P[2] |= ((cd >> 2) & 1u) << pk;
        P[3] |= ((cd >> 3) & 1u) << pk;
#pragma unroll
        for (int q = 0; q < SB; q++) {
            SD[q] |= ((cd >> (4 + q)) & 1u) << k;
            SU[q] |= ((cd >> (4 + SB + q)) & 1u) << k;
        }
    }
}

// Reader side (traceback): rebuilds the code byte of one cell.
template <int C, int SB>
struct PlaneTrace {
    const uint32_t* planes;
    const uint32_t* side;
    const uint8_t* rowflags;
    const uint32_t* nwp_ord;
    __device__ __forceinline__ uint32_t code(int64_t rr, int64_t cc) const {
        constexpr int NW = PlaneFmt<C>::NW;
        const uint32_t ln = (uint32_t)cc / C, kc = (uint32_t)cc % C, k = PlaneFmt<C>::perm(kc);
        const uint32_t* w = planes + (size_t)rr * PlaneFmt<C>::ROWW + ln * NW;
        unsigned t, d, x, y;
        if constexpr (C == 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(w);
            t = (v.x >> k) & 1u, d = (v.y >> k) & 1u, x = (v.z >> k) & 1u, y = (v.w >> k) & 1u;
        } else if constexpr (C == 16) {
            const uint2 v = *reinterpret_cast<const uint2*>(w);
            t = (v.x >> k) & 1u, d = (v.x >> (16 + k)) & 1u, x = (v.y >> k) & 1u, y = (v.y >> (16 + k)) & 1u;
        } else {
            const unsigned v = w[0];
            t = (v >> k) & 1u, d = (v >> (C + k)) & 1u, x = (v >> (2 * C + k)) & 1u, y = (v >> (3 * C + k)) & 1u;
        }
        uint32_t cd = (t ? (d ? (unsigned)DIR_D : (unsigned)DIR_L) : (unsigned)DIR_U) | (x << 2) | (y << 3);
        if (rr == 0 && cc == 0) cd = DIR_O;
        const uint8_t rf = rowflags[rr];
        if ((rf & RF_NWP) && !(rf & RF_SINGLE_PREV)) {
            const uint32_t* sp = side + (size_t)nwp_ord[rr] * (2 * SB * 32) + ln;
#pragma unroll
            for (int q = 0; q < SB; q++) {
                cd |= ((sp[q * 32] >> kc) & 1u) << (4 + q);               // the slot planes keep the column order
                cd |= ((sp[(SB + q) * 32] >> kc) & 1u) << (4 + SB + q);
            }
        }
        return cd;
    }
};

// max.s16x2 whose two result predicates (a >= b per half; INV: a < b) set bit `bl` / `bh` of two accumulators.
// Plain selects and ORs on purpose: a predicated read-modify-write chain on the accumulator makes ptxas keep dozens of
// predicates alive (it hoists the maxima) and spill them through general registers.
template <bool INV>
__device__ __forceinline__ unsigned vmax_flag(unsigned a, unsigned b, unsigned& acc_lo, unsigned& acc_hi, unsigned bl, unsigned bh) {
    bool ph, pl;
    const unsigned val = __vibmax_s16x2(a, b, &ph, &pl);
    acc_lo |= (pl != INV) ? bl : 0u;
    acc_hi |= (ph != INV) ? bh : 0u;
    return val;
}

template <int C>
__device__ __forceinline__ void store_row(int32_t* dst, const int (&v)[C]) {
#pragma unroll
    for (int j = 0; j < C; j += 4) reinterpret_cast<int4*>(dst)[j / 4] = make_int4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

// ---- packed 16-bit helpers (sm_100a: VIADD.16x2, VIMNMX.S16x2 with one predicate per half, VIADDMNMX.S16x2)
// Packed rows keep BIASED fields: field = (score - base16) + 32768 as an unsigned 16-bit number, two cells per register.
// With every field of every operand inside [lowest garbage, highest real] and that span below 32768, plain 32-bit
// additions and subtractions of packed words are exact per field (no carry or borrow crosses the halves), so they can
// issue on either integer pipe, and "a >= b" of both halves is bit 15 / bit 31 of a + 0x80008000 - b.
#define FLOOR16 (-24000)  // padding columns / unused cells start here (relative to base16)
#define PADSUB16 (-2000)  // diagonal "score" of padding columns: keeps them strictly below their left neighbour
#define BIAS2 0x80008000u
#define PK_ENTER_LO (-12000)  // a row may enter the packed form when its real cells lie in base16 + [LO, HI]
#define PK_ENTER_HI (1000)
#define PK_GATHER_LO (-20000)  // predecessor rows gathered while packed (base16 may be up to ~4000 stale)
#define PK_GATHER_HI (3000)   // + 32 rows of upward drift (60 per row: 30 of score, 30 of the moving base) = 4920: still 32768 below the lowest padding value
#define PK_SPREAD (14000)     // guard: leave the packed form when a real cell falls this far below the row maximum
#define PK_REBASE (2000)      // guard: re-base when the row maximum drifted this far from base16
#define PK_GUARD_ROWS 32      // rows between two guards; fields move by at most 90 per row (60 of score, see en16, + |e| of the moving base)
__device__ __forceinline__ unsigned pk16(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
__device__ __forceinline__ int lo16(unsigned v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned v) { return (int)v >> 16; }
__device__ __forceinline__ unsigned pk16b(int lo, int hi) { return pk16(lo, hi) ^ BIAS2; }   // signed pair -> biased fields
__device__ __forceinline__ unsigned add2(int c) { return (unsigned)c * 65537u; }                // addend: c on both fields

#ifndef RG_BLK_CTAS
#define RG_BLK_CTAS 1
#endif
// One CTA of 12 warps per SM (170 registers per thread): the packed row body is spill-free at that size and three warps
// per scheduler hide its fixed-latency dependencies better than two (8 warps x 255 registers: 449 vs 484 ms per 3 552
// reads; 16 warps x 128 registers spill 2.7 KB per thread and are slower than either).
#ifndef RG_BLK_WARPS
#define RG_BLK_WARPS 12
#endif
constexpr int BLK_WARPS = RG_BLK_WARPS;   // warps (= reads in flight) per CTA of this kernel
template <int C, int SB, bool SIMPLE>
__global__ void __launch_bounds__(BLK_WARPS * 32, RG_BLK_CTAS)
    k_gap_global_blk(DevGraph g, DevScoring sc, PoaWorkspace ws, PoaBatch b) {
    static_assert(C % 4 == 0, "C must be a multiple of 4");
    constexpr int STRIDE = 32 * C;
    constexpr unsigned SMASK = (1u << SB) - 1;
    __shared__ int32_t s_sc[48];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    // slots go round the CTAs first: when trace memory allows fewer reads in flight than 12 per SM (large graphs), they
    // spread over all SMs instead of filling a few
    const uint32_t slot = (uint32_t)wib * gridDim.x + blockIdx.x;
    if (threadIdx.x < 48) s_sc[threadIdx.x] = (&sc.sc[0][0])[threadIdx.x];
    __syncthreads();
    if (slot >= ws.slots) return;
    // 16-bit packed fast path: two cells per register, pair r of a lane = columns (r, r + C/2) of its block.
    // Substitution scores of the pairs for each graph base: [5 bases][C/2 pairs][32 lanes] words per warp.
    constexpr int H = C / 2;
    extern __shared__ unsigned s_subtab[];
    unsigned* const subtab = s_subtab + (size_t)wib * 5 * H * 32;

    const uint32_t n = g.n;
    const uint32_t RM = g.ring - 1;
    RowMeta* rowmeta = ws.rowmeta + (size_t)slot * n;
    int32_t* ring_m = ws.ring_m + (size_t)slot * g.ring * STRIDE;
    int32_t* ring_y = ws.ring_y + (size_t)slot * g.ring * STRIDE;
    uint32_t* planes = reinterpret_cast<uint32_t*>(ws.trace + (size_t)slot * ws.trace_cap);
    uint32_t* side = reinterpret_cast<uint32_t*>(ws.trace + (size_t)slot * ws.trace_cap + ws.side_off);
    constexpr int NW = PlaneFmt<C>::NW;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const int o = sc.o, e = sc.e;
    const int c1 = e + max(o, 0), c2 = o + e;
    const int cbase = lane * C;

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
        if (L > STRIDE || (uint64_t)n * PlaneFmt<C>::ROWW * 4 > ws.side_off ||
            ws.side_off + (uint64_t)g.n_gather * 2 * SB * 128 > ws.trace_cap) {  // host sizes the launch so that this never happens
            res.status = RG_READ_TRACE_OVERFLOW;
            if (lane == 0) b.results[ridx] = res;
            continue;
        }

        // read codes of my columns (column c aligns read[c-1]); N (4) for padding
        unsigned rcw[C / 4];
#pragma unroll
        for (int j = 0; j < C / 4; j++) {
            unsigned w = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int c = cbase + j * 4 + q;
                unsigned code = (c >= 1 && c < L) ? read[c - 1] : 4u;
                w |= code << (8 * q);
            }
            rcw[j] = w;
        }

        // match/mismatch scoring (SIMPLE): per base one bit mask over my columns, so that the substitution score of
        // a cell is a bit test instead of a shared-memory table lookup
        unsigned eqm0 = 0, eqm1 = 0, eqm2 = 0, eqm3 = 0;
        if (SIMPLE) {
#pragma unroll
            for (int k = 0; k < C; k++) {
                const unsigned rc = (rcw[k / 4] >> (8 * (k % 4))) & 0xffu;
                eqm0 |= (rc == 0u ? 1u : 0u) << k;
                eqm1 |= (rc == 1u ? 1u : 0u) << k;
                eqm2 |= (rc == 2u ? 1u : 0u) << k;
                eqm3 |= (rc == 3u ? 1u : 0u) << k;
            }
        }
        const int s_match = sc.sc[0][0], s_mis = sc.sc[0][1];
        // exact while no 16-bit lane can wrap: small scores, strictly negative gap steps
        bool en16 = SIMPLE && (e + max(o, 0) < 0) && (o + e < 0) && abs(o) <= 30 && abs(e) <= 30 && abs(s_match) <= 30 &&
                    abs(s_mis) <= 30 && ws.use16;
        if (SIMPLE && en16) {
#pragma unroll
            for (int r = 0; r < H; r++) {
                const unsigned rl = (rcw[r / 4] >> (8 * (r % 4))) & 0xffu, rh = (rcw[(r + H) / 4] >> (8 * ((r + H) % 4))) & 0xffu;
#pragma unroll
                for (unsigned bse = 0; bse < 5; bse++) {
                    // padding columns (c >= L) get a diagonal score so low that their m never reaches their left
                    // neighbour's: x and y of a padding cell are already strictly below it (c1, c2 < 0), so a padding
                    // column can never be the row maximum and needs no band mask in the packed path
                    const int sl = (cbase + r >= L) ? PADSUB16 : ((bse < 4 && rl == bse) ? s_match : s_mis);
                    const int sh = (cbase + r + H >= L) ? PADSUB16 : ((bse < 4 && rh == bse) ? s_match : s_mis);
                    // addend: sl on the lo field, sh on the hi field; - e because every packed row moves base16 by e (row16)
                    subtab[(bse * H + r) * 32 + lane] = (unsigned)((sh - e) * 65536 + (sl - e));
                }
            }
            __syncwarp();
        }
        bool rep16 = false;   // the previous row is held packed (A[r], B[r] for r < H hold m and y relative to base16)
        int base16 = 0, rows16 = 0, prev_tmax = 0;
        const unsigned C1F = pk16(e + max(o, 0), e + max(o, 0));  // per-field operand of VIADDMNMX.U16x2
        const unsigned C2A = add2(o + e);   // 32-bit addend
        const unsigned OF = pk16(o, o);     // per-field operand of VIADDMNMX.U16x2
        const unsigned KX = BIAS2 - add2(o + 1);
        // The integer ALU pipe issues one warp instruction every two clocks and so does the multiply-add pipe. The row
        // body is ALU-bound (maxima, shifts, logic), so the differences behind the flags are written as multiply-adds
        // with multipliers the compiler cannot fold (ws.use16 is 1 whenever this code runs): they issue as IMAD.
        const unsigned ONE = ws.use16, NEG1 = 0u - ONE;
        const long long t_start = clock64();
        int A[C], B[C];  // previous row: m and y of my columns (NEG_INF outside its band)
        int status = 0;
        uint64_t cells = 0;
        uint32_t prev_bsp = 0, prev_left = 0, prev_right = 0;
        int best_end_val = NEG_INF;
        uint32_t best_end_row = 0;
        int last_val = 0;
        bool abort_read = false;
        int last_nonfull = -1;  // last row whose band is not the whole read [0, L)
#ifdef RG_ROWSTATS
        unsigned st_cnt[5] = {0, 0, 0, 0, 0};
        long long st_cyc[5] = {0, 0, 0, 0, 0};
#endif

        // per-row values shared by the row bodies below
        int li = 0;                  // graph base of the row
        int best_p = 0;              // smallest predecessor (seed of the first column, gap_global_abpoa.rs:88)
        unsigned P[4];               // trace planes of my columns (T, D, X, Y)
        bool have_planes = false;
        int bestv = NEG_INF;
        int bcol = 0;
        unsigned bp16 = 0;           // packed row: maxima of my lo / hi cells (biased fields)
        // Packed row (two cells per register) for a row whose band and whose predecessor row(s) span the whole read.
        // A / B hold the (gathered) predecessor row on entry and this row on exit; leaves P, bestv, bcol.
        auto row16 = [&]() {
            const unsigned* tab = subtab + (size_t)li * H * 32 + lane;
            // The row is computed relative to base16 + e: y = max(m' + o, y') + e of the previous row's fields m', y' is then
            // max(m' + o, y') as it stands, and the e is folded into the substitution addends — no per-cell "+ e" at all.
            base16 += e;
            unsigned D16[H];   // B[r] becomes this row's y in pass A and stays
            // Flag accumulators in the plane bit order (PlaneFmt::perm). A flag is bit 15 / bit 31 of a difference word;
            // the words of pairs r, r + 1 (r even) become one byte of 0xff / 0x00 per cell with one PRMT (sign replication:
            // bytes lo r, lo r+1, hi r, hi r+1) and enter bit r / 2 of the four byte groups with one AND-OR. (The same step as
            // a multiply-add — a word of 0xff / 0x00 bytes times 0xfefefeff << q, the inverse of 255 modulo 2^32 — measured
            // no faster: the row is bound by issue slots, not by one pipe.)
            unsigned fy = 0, fd = 0, ft = 0, fx = 0;
            auto flag2 = [](unsigned& acc, unsigned f_even, unsigned f_odd, int r_even) {
                unsigned w;   // PTX prmt: bit 3 of a selector nibble replicates the byte's sign (__byte_perm drops that bit)
                asm("prmt.b32 %0, %1, %2, 0xfbd9;" : "=r"(w) : "r"(f_even), "r"(f_odd));
                acc |= w & (0x01010101u << (r_even / 2));
            };
            // ---- pass A: y and d of both halves of every pair (descending: A[r-1] is still the previous row)
            const unsigned up = __shfl_up_sync(FULL, (unsigned)A[H - 1], 1);
            const unsigned dg0 = __byte_perm(up, (unsigned)A[H - 1], 0x5432);  // lo <- cell cbase-1, hi <- cell H-1
            unsigned fyo = 0;
#pragma unroll
            for (int r = H - 1; r >= 0; r--) {
                const unsigned fyr = (unsigned)B[r] + KX - (unsigned)A[r];                  // Y: y > m + o
                const unsigned yv = __viaddmax_u16x2((unsigned)A[r], OF, (unsigned)B[r]);   // max(m + o, y) (+ e: the base moved)
                if (r & 1) fyo = fyr; else flag2(fy, fyr, fyo, r);
                const unsigned dd = ((r == 0) ? dg0 : (unsigned)A[r - 1]) + tab[r * 32];
                D16[r] = dd;
                B[r] = (int)yv;
                A[r] = (int)__vmaxu2(dd, yv);  // h
            }
            const unsigned FLB = (unsigned)(FLOOR16 + 32768);
            if (lane == 0) {  // first-column cell (gap_global_abpoa.rs:78-92): m = x only
                D16[0] = (D16[0] & 0xffff0000u) | FLB;
                B[0] = (int)(((unsigned)B[0] & 0xffff0000u) | FLB);
                A[0] = (int)(((unsigned)A[0] & 0xffff0000u) | FLB);
            }
            // ---- pass B: two in-lane chains (lo cells 0..H-1, hi cells H..C-1); generator of cell c is h[c-1] + c2.
            // Only the value leaving my last column is needed here (for the cross-lane scan); pass C runs the chain again
            // from the true incoming value instead of keeping 16 intermediate registers. The chain runs on x - c2, whose
            // generator is h[c-1] itself (one instruction per pair).
            const unsigned hup = __shfl_up_sync(FULL, (unsigned)A[H - 1], 1);
            const unsigned g0s = __byte_perm(hup, (unsigned)A[H - 1], 0x5432);
            unsigned g0 = g0s + C2A;
            if (lane == 0) g0 = (g0 & 0xffff0000u) | ((unsigned)(o + e * (best_p + 1) - base16 + 32768) & 0xffffu);  // seed, :88
            unsigned xe;
            {
                unsigned xl = (FLB | (FLB << 16)) - C2A;
#pragma unroll
                for (int r = 0; r < H; r++) xl = __viaddmax_u16x2(xl, C1F, (r == 0) ? g0 - C2A : (unsigned)A[r - 1]);
                xe = xl + C2A;
            }
            // cross-lane max-plus scan on the in-lane value of my last column (biased integers)
            const int c1s = e + max(o, 0);
            const int xlo_end = (int)(xe & 0xffffu);
            const int agg = max((int)(xe >> 16), xlo_end + H * c1s);
            const int z = agg - (cbase + C - 1) * c1s;
            const int winc = warp_incl_max(z, lane);
            int wexc = __shfl_up_sync(FULL, winc, 1);
            if (lane == 0) wexc = NEG_INF;
            const int xin = max(wexc + cbase * c1s, (int)FLB);            // x entering my first column
            const int xf = max(xlo_end, xin + (H - 1) * c1s);             // final x of my column H-1
            // chain state "one column before my first": the first step adds c1 back
            unsigned xl = (unsigned)(xin - c1s) | ((unsigned)(max(xf + c1s, (int)FLB) - c1s) << 16);
            // ---- pass C
            unsigned bestp = FLB | (FLB << 16);
            unsigned hprev = 0;
            unsigned fde = 0, fte = 0, fxe = 0;
#pragma unroll
            for (int r = 0; r < H; r++) {
                const unsigned gen = (r == 0) ? g0 : hprev + C2A;
                hprev = (unsigned)A[r];
                xl = __viaddmax_u16x2(xl, C1F, gen);
                const unsigned x = xl;
                const unsigned t = __vmaxu2(D16[r], x);
                const unsigned fdr = D16[r] * ONE + (x * NEG1 + BIAS2);      // D: dd >= x
                const unsigned m = __vmaxu2(t, (unsigned)B[r]);
                const unsigned ftr = t * ONE + ((unsigned)B[r] * NEG1 + BIAS2);   // T: max(dd, x) >= y
                const unsigned fxr = m * NEG1 + (x * ONE + KX);              // x > m + o
                if (r & 1) {
                    flag2(fd, fde, fdr, r - 1);
                    flag2(ft, fte, ftr, r - 1);
                    flag2(fx, fxe, fxr, r - 1);
                } else {
                    fde = fdr, fte = ftr, fxe = fxr;
                }
                bestp = __vmaxu2(m, bestp);   // (pairs of these fuse into one three-input maximum)
                A[r] = (int)m;
            }
            if constexpr (H < 16) {  // the byte groups hold G = H / 2 bits each: close the gaps
                constexpr int G = H / 2;
                auto squeeze = [](unsigned v) {
                    constexpr unsigned gm = (1u << G) - 1u;
                    return (v & gm) | (((v >> 8) & gm) << G) | (((v >> 16) & gm) << (2 * G)) | (((v >> 24) & gm) << (3 * G));
                };
                fy = squeeze(fy), fd = squeeze(fd), ft = squeeze(ft), fx = squeeze(fx);
            }
            // path_x of column c is the flag of column c - 1 (gap_global_abpoa.rs:358-364): move every bit to the plane
            // position of the next column (even cell -> odd cell of its pair group: + G; odd cell -> next even: - G + 1;
            // last cell of the lo half -> first of the hi half: + 1; last cell of the block -> next lane)
            unsigned carry = __shfl_up_sync(FULL, fx >> (C - 1), 1) & 1u;   // fx bit of column k: x[c] > m[c] + o
            if (lane == 0) carry = 0;
            {
                constexpr int G = C / 4;
                constexpr unsigned gm = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
                constexpr unsigned EVEN = gm | (gm << (2 * G));
                constexpr unsigned ODDIN = ((gm >> 1) << G) | ((gm >> 1) << (3 * G));   // odd cells except the last of each half
                constexpr unsigned LOLAST = 1u << (2 * G - 1);
                unsigned nx = ((fx & EVEN) << G) | ((fx & LOLAST) << 1) | carry;
                if constexpr (G > 1) nx |= (fx & ODDIN) >> (G - 1);
                P[2] = nx;
            }
            P[0] = ft;
            P[1] = fd;
            P[3] = fy;
            if (lane == 0) {                     // first-column cell: vertical move to the smallest predecessor, no flags
                P[0] &= ~1u;
                P[1] &= ~1u;
                P[2] &= ~1u;
                P[3] &= ~1u;
            }
            have_planes = true;
            bp16 = bestp;
            bestv = base16 + (int)max(bestp & 0xffffu, bestp >> 16) - 32768;
        };
        // right-most column of my block that holds my maximum of the packed row just computed (A = its m values)
        auto bcol16 = [&]() {
            const unsigned blo = bp16 & 0xffffu, bhi = bp16 >> 16;
            const bool hi = bhi >= blo;
            const unsigned tgt = hi ? bhi : blo;
            int pos = 0;
#pragma unroll
            for (int r = 0; r < H; r++) {
                const unsigned f = hi ? ((unsigned)A[r] >> 16) : ((unsigned)A[r] & 0xffffu);
                if (f == tgt) pos = r;
            }
            bcol = cbase + (hi ? H : 0) + pos;
        };
        // fy in column order (the slot planes of a gathered row are kept in column order)
        auto unperm = [](unsigned v) {
            unsigned out = 0;
#pragma unroll
            for (int k = 0; k < C; k++) out |= ((v >> PlaneFmt<C>::perm((unsigned)k)) & 1u) << k;
            return out;
        };

        // Packed gather of the predecessor rows of a segment start (all of them span the whole read, none is row 0):
        // ring rows -> biased pairs in A / B, winner slots as bit planes. Returns false (A / B garbage) when a value
        // does not fit the packed range.
        unsigned mm_lo[SB], mm_hi[SB], ym_lo[SB], ym_hi[SB];
        auto gather16 = [&](uint32_t pb, uint32_t np) -> bool {
        const int lo_ok = base16 + PK_GATHER_LO;
        bool ok = true;
#pragma unroll
        for (int bq = 0; bq < SB; bq++) mm_lo[bq] = mm_hi[bq] = ym_lo[bq] = ym_hi[bq] = 0u;
        for (uint32_t q = 0; q < np; q++) {
            const uint32_t p = g.pred_idx[pb + q];
            const int32_t* mp = ring_m + (size_t)(p & RM) * STRIDE + cbase;
            const int32_t* yp = ring_y + (size_t)(p & RM) * STRIDE + cbase;
            unsigned wm_lo = 0, wm_hi = 0, wy_lo = 0, wy_hi = 0;  // cells where predecessor q wins
            auto take = [&](int r, int vml, int vmh, int vyl, int vyh) {
                const int cl = cbase + r, ch = cbase + r + H;
                const bool rl = cl < L, rh = ch < L;
                ok = ok && (!rl || ((unsigned)(vml - lo_ok) <= (unsigned)(PK_GATHER_HI - PK_GATHER_LO) && (cl == 0 || (unsigned)(vyl - lo_ok) <= (unsigned)(PK_GATHER_HI - PK_GATHER_LO))));
                ok = ok && (!rh || ((unsigned)(vmh - lo_ok) <= (unsigned)(PK_GATHER_HI - PK_GATHER_LO) && (unsigned)(vyh - lo_ok) <= (unsigned)(PK_GATHER_HI - PK_GATHER_LO)));
                const unsigned nm = pk16(rl ? vml - base16 : FLOOR16, rh ? vmh - base16 : FLOOR16);
                const unsigned ny = pk16((rl && cl != 0) ? vyl - base16 : FLOOR16, rh ? vyh - base16 : FLOOR16);
                if (q == 0) {
                    A[r] = (int)nm;
                    B[r] = (int)ny;
                } else {
                    bool ph, pl;
                    A[r] = (int)__vibmax_s16x2((unsigned)A[r], nm, &ph, &pl);  // p = (current >= new): keep
                    if (!pl) wm_lo |= 1u << r;
                    if (!ph) wm_hi |= 1u << r;
                    B[r] = (int)__vibmax_s16x2((unsigned)B[r], ny, &ph, &pl);
                    if (!pl) wy_lo |= 1u << r;
                    if (!ph) wy_hi |= 1u << r;
                }
            };
            if constexpr (H >= 4) {
#pragma unroll
                for (int j = 0; j < H; j += 4) {
                    const int4 mlv = reinterpret_cast<const int4*>(mp)[j / 4], mhv = reinterpret_cast<const int4*>(mp + H)[j / 4];
                    const int4 ylv = reinterpret_cast<const int4*>(yp)[j / 4], yhv = reinterpret_cast<const int4*>(yp + H)[j / 4];
                    take(j, mlv.x, mhv.x, ylv.x, yhv.x);
                    take(j + 1, mlv.y, mhv.y, ylv.y, yhv.y);
                    take(j + 2, mlv.z, mhv.z, ylv.z, yhv.z);
                    take(j + 3, mlv.w, mhv.w, ylv.w, yhv.w);
                }
            } else {  // C == 4: one 128-bit load holds both halves
                const int4 mv = reinterpret_cast<const int4*>(mp)[0], yv = reinterpret_cast<const int4*>(yp)[0];
                take(0, mv.x, mv.z, yv.x, yv.z);
                take(1, mv.y, mv.w, yv.y, yv.w);
            }
#pragma unroll
            for (int bq = 0; bq < SB; bq++) {
                const bool on = (q >> bq) & 1u;
                mm_lo[bq] = on ? (mm_lo[bq] | wm_lo) : (mm_lo[bq] & ~wm_lo);
                mm_hi[bq] = on ? (mm_hi[bq] | wm_hi) : (mm_hi[bq] & ~wm_hi);
                ym_lo[bq] = on ? (ym_lo[bq] | wy_lo) : (ym_lo[bq] & ~wy_lo);
                ym_hi[bq] = on ? (ym_hi[bq] | wy_hi) : (ym_hi[bq] & ~wy_hi);
            }
        }
            if (!__all_sync(FULL, ok)) return false;
#pragma unroll
            for (int r = 0; r < H; r++) {   // the gather compares signed pairs; rows are kept biased
                A[r] = (int)((unsigned)A[r] ^ BIAS2);
                B[r] = (int)((unsigned)B[r] ^ BIAS2);
            }
            return true;
        };
        // slot planes of a gathered row: the diagonal source of column c is the m-winner of column c-1, the vertical source
        // is the y-winner where y extends and the m-winner where it opens; the first-column cell goes to the smallest
        // predecessor (slot mps)
        auto slots16 = [&](unsigned mps, unsigned (&SD)[SB], unsigned (&SU)[SB]) {
            const unsigned py = unperm(P[3]);
#pragma unroll
            for (int bq = 0; bq < SB; bq++) {
                const unsigned mm = mm_lo[bq] | (mm_hi[bq] << H), ym = ym_lo[bq] | (ym_hi[bq] << H);
                unsigned cin = __shfl_up_sync(FULL, mm >> (C - 1), 1) & 1u;
                if (lane == 0) cin = 0;
                SD[bq] = (mm << 1) | cin;
                if constexpr (C < 32) SD[bq] &= (1u << C) - 1u;
                SU[bq] = (py & ym) | (~py & mm);
                if (lane == 0) SU[bq] = (SU[bq] & ~1u) | ((mps >> bq) & 1u);
            }
        };
        auto side_store = [&](uint32_t row, const unsigned (&SD)[SB], const unsigned (&SU)[SB]) {
            uint32_t* sp = side + (size_t)g.nwp_ord[row] * (2 * SB * 32) + lane;
#pragma unroll
            for (int bq = 0; bq < SB; bq++) {
                sp[bq * 32] = SD[bq];
                sp[(SB + bq) * 32] = SU[bq];
            }
        };
        // packed -> 32-bit row (padding columns are outside the band: NEG_INF; y[i][0] is 0 by definition)
        auto unpack16 = [&](int (&VM)[C], int (&VY)[C]) {
#pragma unroll
            for (int r = H - 1; r >= 0; r--) {
                const unsigned pa = (unsigned)A[r] ^ BIAS2, py = (unsigned)B[r] ^ BIAS2;
                const int cl = cbase + r, ch = cbase + r + H;
                VM[r + H] = (ch < L) ? base16 + hi16(pa) : NEG_INF;
                VY[r + H] = (ch < L) ? base16 + hi16(py) : NEG_INF;
                VM[r] = (cl < L) ? base16 + lo16(pa) : NEG_INF;
                VY[r] = (cl < L) ? ((cl == 0) ? 0 : base16 + lo16(py)) : NEG_INF;
            }
        };
        // the ring keeps 32-bit rows: unpack on the way out
        auto ring_store16 = [&](uint32_t row) {
            int VM[C], VY[C];
            unpack16(VM, VY);
            store_row<C>(ring_m + (size_t)(row & RM) * STRIDE + cbase, VM);
            store_row<C>(ring_y + (size_t)(row & RM) * STRIDE + cbase, VY);
        };
        // Range guard + re-basing, every PK_GUARD_ROWS packed rows (biased fields compare like the scores). Padding
        // columns (c >= L) are ignored: they sit at the floor right after a packed gather. Returns true when the read no
        // longer fits the packed range.
        auto guard16 = [&](int tmax) -> bool {
            unsigned mn = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < H; r++) {
                unsigned yy = (unsigned)B[r];
                if (r == 0 && lane == 0) yy = (yy & 0xffff0000u) | ((unsigned)A[0] & 0xffffu);  // the first-column cell's y is unused
                const unsigned v = __vminu2((unsigned)A[r], yy);
                const unsigned real = ((cbase + r < L) ? 0u : 0xffffu) | ((cbase + r + H < L) ? 0u : 0xffff0000u);
                mn = __vminu2(mn, __vmaxu2(v, real));
            }
            const int lm = __reduce_min_sync(FULL, (int)min(mn & 0xffffu, mn >> 16)) - 32768;
            const int rel = tmax - base16;
            if (lm - rel < -PK_SPREAD) return true;
            if (rel > PK_REBASE || rel < -PK_REBASE) {
                // |rel| <= PK_REBASE + PK_GUARD_ROWS * 90: no field can leave its half; padding is pulled back to the floor
                const unsigned dl = add2(-rel), fl = (unsigned)(FLOOR16 + 32768) * 65537u;
#pragma unroll
                for (int r = 0; r < H; r++) {
                    A[r] = (int)__vmaxu2((unsigned)A[r] + dl, fl);
                    B[r] = (int)__vmaxu2((unsigned)B[r] + dl, fl);
                }
                base16 += rel;
            }
            return false;
        };

        int4 ri_next = reinterpret_cast<const int4*>(g.rowinfo)[0];
        for (uint32_t i = 0; i + 1 < n; i++) {
            // ---- steady state: consecutive rows in the packed representation whose band (and whose predecessors' bands)
            // span the whole read. One compact loop: the row body is straight-line code, everything that only some rows
            // need (gathering predecessors, slot planes, ring copy, range guard) is a side block. Any other kind of row
            // (32-bit rows, end-cell candidates, shifted bands, row 0 as predecessor) leaves the loop and goes through
            // the dispatcher below.
            if (rep16) {
                bool full_next = false;   // the previous steady row proved that an in-segment successor has the band [0, L)
                for (; i + 2 < n; i++) {
#ifdef RG_ROWSTATS
                    const long long st_t1 = clock64();
#endif
                    const int4 rv = ri_next;  // rowinfo[i]
                    const uint32_t rb = (uint32_t)rv.w;
                    const uint32_t rfl = (rb >> 8) & 0xffu;
                    if (rfl & RF_F_PRED) break;
                    const bool gat = (rfl & (RF_NWP | RF_SINGLE_PREV)) == RF_NWP;
                    uint32_t ms = prev_bsp + 1, me = prev_bsp + 1;
                    if (gat) {
                        if (rv.y < 1 || rv.y <= last_nonfull) break;
                        const uint32_t pb = (uint32_t)rv.z, np = rb >> 24;
                        uint32_t pl = 0xffffffffu, pr = 0;
                        for (uint32_t k = 0; k < np; k++) {
                            const uint32_t p = g.pred_idx[pb + k];
                            const uint32_t bs = (p == i - 1) ? prev_bsp : rowmeta[p].bsp;
                            pl = min(pl, bs);
                            pr = max(pr, bs);
                        }
                        ms = pl + 1;
                        me = pr + 1;
                    }
                    if (gat || !full_next) {
                        uint32_t lf, rt;
                        band_for_row(ms, me, rv.x, L, bta, lf, rt);
                        if (lf != 0 || rt != (uint32_t)L) break;
                    }
                    {   // the first-column seed o + e * (best_p + 1) (:88) follows the ROW INDEX of the smallest predecessor,
                        // not the scores: after a long skip edge it can sit thousands above / below the row. Outside the
                        // packed window the row is done by the exact 32-bit path (the dispatcher re-checks and unpacks).
                        const int sd = o + e * (rv.y + 1) - (base16 + e);   // row16 moves the base by e first
                        if (sd < PK_GATHER_LO || sd > PK_GATHER_HI) break;
                    }
                    if (gat && !gather16((uint32_t)rv.z, rb >> 24)) {
                        rep16 = false;  // A / B are garbage now; the dispatcher's general path gathers from the ring itself
                        break;
                    }
                    ri_next = reinterpret_cast<const int4*>(g.rowinfo)[i + 1];
                    li = rb & 0xffu;
                    best_p = rv.y;
                    row16();
                    const int tmax = __reduce_max_sync(FULL, bestv);
                    const unsigned eq = __ballot_sync(FULL, bestv == tmax);
                    // Column of the right-most row maximum (gap_global_abpoa.rs:198-203). It only feeds the bands of the rows
                    // that have this row as predecessor (utils.rs:17-66): their left edge is min(bsp + 1, L - r - bta) and
                    // their right edge min(L, max(bsp + 1, ..) + bta). Every such row j has r_j >= r_i - 1 (set_r_values,
                    // utils.rs:106-126), so with r_i - 1 >= L - bta its left edge is 0 whatever bsp is, and any lower bound
                    // lb of bsp with lb + 1 + bta >= L gives the right edge L exactly like bsp itself. The start of the
                    // half block that holds the maximum is such a bound for most rows; the exact column is searched
                    // otherwise.
                    uint32_t row_bsp = (uint32_t)__shfl_sync(FULL, cbase + (((bp16 >> 16) >= (bp16 & 0xffffu)) ? H : 0), 31 - __clz(eq));
                    full_next = rv.x >= 1 && rv.x - 1 >= L - bta && (int)row_bsp + 1 + bta >= L;   // (the successor's band too)
                    if (!full_next) {
                        bcol16();
                        row_bsp = (uint32_t)__shfl_sync(FULL, bcol, 31 - __clz(eq));
                    }
                    store_planes<C>(planes + (size_t)i * PlaneFmt<C>::ROWW + lane * NW, P);
                    if (lane == 0) {
                        RowMeta rm;
                        rm.base = 0;
                        rm.left = 0;
                        rm.right = (uint32_t)L;
                        rm.bsp = row_bsp;
                        rowmeta[i] = rm;
                    }
                    if (gat) {
                        unsigned SD[SB], SU[SB];
                        slots16((rb >> 16) & 0xffu, SD, SU);
                        side_store(i, SD, SU);
                    }
                    prev_tmax = tmax;
                    prev_bsp = row_bsp;
                    cells += (uint32_t)L;
                    rows16++;
#ifdef RG_ROWSTATS
                    st_cnt[gat ? 1 : 0]++;
                    st_cyc[gat ? 1 : 0] += clock64() - st_t1;
#endif
                    if ((rows16 & (PK_GUARD_ROWS - 1)) == 0 && guard16(tmax)) {
                        en16 = false;  // this read does not fit 16 bits: stay on the 32-bit paths
                        if (rfl & RF_IS_PRED) ring_store16(i);
                        unpack16(A, B);
                        rep16 = false;
                        i++;
                        break;
                    }
                    if (rfl & RF_IS_PRED) ring_store16(i);
                    __syncwarp();
                }
                if (i + 1 >= n) break;
            }
            // packed row info, fetched one row ahead (hides the L1/L2 latency behind the previous row's work)
#ifdef RG_ROWSTATS
            const long long st_t0 = clock64();
            int st_kind = 4;
#endif
            const int4 riv = ri_next;
            ri_next = reinterpret_cast<const int4*>(g.rowinfo)[i + 1];
            const uint32_t rbits = (uint32_t)riv.w;
            const uint8_t rf = (rbits >> 8) & 0xffu;
            // a segment start whose only predecessor is row i-1 behaves exactly like a row inside a segment
            const bool nwp = (rf & RF_NWP) && !(rf & RF_SINGLE_PREV);
            const uint32_t pb = (uint32_t)riv.z, pe = pb + (nwp ? (rbits >> 24) : 0u);
            best_p = riv.y;
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
            band_for_row(ms, me, riv.x, L, bta, left, right);
            if (right < left) {  // reference: `right - left` overflows (gap_global_abpoa.rs:59)
                status |= RG_READ_REF_PANIC;
                abort_read = true;
                break;
            }
            if (right == left) {
                // An EMPTY row (only possible with b + f * L < 1) is legal in the reference: zero cells, best_scoring_pos =
                // left (:203), and the rows after it find none of its cells available (every access is guarded by the
                // band bounds, :254-346). Only the end-cell selection indexes it unconditionally (:205-215).
                if ((rf & RF_F_PRED) || i == n - 2) {
                    status |= RG_READ_REF_PANIC;
                    abort_read = true;
                    break;
                }
                rep16 = false;
#pragma unroll
                for (int k = 0; k < C; k++) A[k] = B[k] = NEG_INF;
                P[0] = P[1] = P[2] = P[3] = 0u;
                store_planes<C>(planes + (size_t)i * PlaneFmt<C>::ROWW + lane * NW, P);
                if (nwp) {
                    unsigned SD0[SB], SU0[SB];
#pragma unroll
                    for (int q = 0; q < SB; q++) SD0[q] = SU0[q] = 0u;
                    side_store(i, SD0, SU0);
                }
                last_nonfull = (int)i;
                if (rf & RF_IS_PRED) {
                    store_row<C>(ring_m + (size_t)(i & RM) * STRIDE + cbase, A);
                    store_row<C>(ring_y + (size_t)(i & RM) * STRIDE + cbase, B);
                }
                if (lane == 0) {
                    RowMeta rm;
                    rm.base = 0;
                    rm.left = left;
                    rm.right = right;
                    rm.bsp = left;
                    rowmeta[i] = rm;
                }
                __syncwarp();
                prev_bsp = left;
                prev_left = left;
                prev_right = right;
                continue;
            }
            cells += right - left;
            li = rbits & 0xffu;
            const unsigned mps = (rbits >> 16) & 0xffu;
            unsigned code[C];            // 32-bit producers: one code per cell, converted to planes below
            unsigned SD[SB], SU[SB];     // predecessor-slot planes (rows that gather predecessors)
            have_planes = false;
            bestv = NEG_INF;
            bcol = 0;

            // Fast path (a row inside a segment whose band and whose previous row's band start at column 0): every
            // active cell below the previous row's right edge has its vertical and diagonal source; cells of a band
            // that grew to the right take the y fallback of gap_global_abpoa.rs:110-141 and have no diagonal beyond
            // the first new column.
            const bool fast = (i > 0) && !nwp;
            const bool plain = left == 0 && prev_left == 0 && right <= prev_right;

            // first-column seed of a packed row (see the steady loop): must fit the packed window around the base the row will use
            const int sd16 = o + e * (best_p + 1) - ((rep16 ? base16 : prev_tmax) + e);
            const bool seed_ok = sd16 >= PK_GATHER_LO && sd16 <= PK_GATHER_HI;
            bool f16 = en16 && seed_ok && fast && plain && right == (uint32_t)L && prev_right == (uint32_t)L;
            // Packed segment-start row: every predecessor row (none of them row 0) and this row span the whole read, so
            // that no availability / fallback rule of gap_global_abpoa.rs:110-141 can trigger. The predecessor rows are
            // gathered from the ring straight into the packed representation; per cell the slot of the winning
            // predecessor (first in list order wins ties, :266-345) is kept as bit planes over my columns.
            bool g16 = en16 && seed_ok && nwp && i > 0 && left == 0 && right == (uint32_t)L && best_p >= 1 && best_p > last_nonfull;
            if (g16) {
                if (!rep16) base16 = prev_tmax;
                if (gather16(pb, pe - pb)) {
                    if (!rep16) rows16 = 0;
                    rep16 = true;
                } else {
                    g16 = false;   // the 32-bit general path gathers from the ring itself: the packed state is dropped
                    rep16 = false;
                }
            }
            if (f16 && !rep16) {
                // 32-bit -> packed: every real cell must fit comfortably; padding columns (c >= L) and the unused
                // y of the first-column cell start at the floor and become "healthy" within a row or two
                bool ok = true;
#pragma unroll
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    if (c < L) {
                        ok = ok && (A[k] - prev_tmax >= PK_ENTER_LO) && (A[k] - prev_tmax <= PK_ENTER_HI);
                        if (c != 0) ok = ok && (B[k] - prev_tmax >= PK_ENTER_LO) && (B[k] - prev_tmax <= PK_ENTER_HI);
                    }
                }
                if (__all_sync(FULL, ok)) {
                    base16 = prev_tmax;
#pragma unroll
                    for (int r = 0; r < H; r++) {
                        const int cl = cbase + r, ch = cbase + r + H;
                        const int al = (cl < L) ? A[r] - base16 : FLOOR16, ah = (ch < L) ? A[r + H] - base16 : FLOOR16;
                        const int yl = (cl < L && cl != 0) ? B[r] - base16 : FLOOR16, yh = (ch < L) ? B[r + H] - base16 : FLOOR16;
                        A[r] = (int)pk16b(al, ah);
                        B[r] = (int)pk16b(yl, yh);
                    }
                    rep16 = true;
                    rows16 = 0;
                } else {
                    f16 = false;
                }
            }
            if (!f16 && !g16 && rep16 && nwp) rep16 = false;  // segment starts read their predecessors from the ring
            if (!f16 && !g16 && rep16) {
                unpack16(A, B);
                rep16 = false;
            }
            // 32-bit row inside a segment (predecessor = row i-1, held in registers). PLAIN: both bands start at column 0
            // and this one lies inside the previous one, so every active cell has its vertical and diagonal source.
            // Otherwise cells without a vertical source take the y fallback of gap_global_abpoa.rs:110-141, cells
            // without a diagonal source have none, and the x chain starts at column `left` (:94-109).
            auto row32 = [&](auto plain_tag) {
                constexpr bool PLAIN = decltype(plain_tag)::value;
                const int32_t* srow = s_sc + li * 8;
                const unsigned em = li == 0 ? eqm0 : (li == 1 ? eqm1 : (li == 2 ? eqm2 : (li == 3 ? eqm3 : 0u)));
                int up = __shfl_up_sync(FULL, A[C - 1], 1);
                if (lane == 0) up = NEG_INF;
                const int fbq = 2 * o + e * (best_p + 1);  // gap_global_abpoa.rs:117,139
                const int seed = (left == 0) ? o + e * (best_p + 1) : fbq + e * (int)left;  // :88 / :99
                const bool fc0 = (lane == 0) && left == 0;
                unsigned ybits = 0;
                int D[C];
                // pass A (descending so that A[k-1] is still the previous row when cell k reads its diagonal)
#pragma unroll
                for (int k = C - 1; k >= 0; k--) {
                    const int um = A[k] + o;
                    const int uy = B[k];
                    int yv = max(um, uy) + e;
                    bool yb = uy > um;
                    if (!PLAIN && !(A[k] > NEGH)) {
                        yv = fbq + e * (cbase + k);
                        yb = false;
                    }
                    if (yb) ybits |= 1u << k;
                    int sub;
                    if (SIMPLE) {
                        sub = ((em >> k) & 1u) ? s_match : s_mis;
                    } else {
                        const unsigned rc = (rcw[k / 4] >> (8 * (k % 4))) & 0xffu;
                        sub = srow[rc];
                    }
                    const int dsrc = (k == 0) ? up : A[k - 1];
                    int dd = dsrc + sub;
                    if (!PLAIN && !(dsrc > NEGH)) dd = NEG_INF;
                    int h = max(dd, yv);
                    if (k == 0 && fc0) {  // first-column cell: m = x only
                        dd = NEG_INF;
                        h = NEG_INF;
                    }
                    B[k] = yv;
                    D[k] = dd;
                    A[k] = h;
                }
                // pass B: in-lane x chain; generator of column c is h[c-1] + c2, the seed sits at column `left`
                int hprev = __shfl_up_sync(FULL, A[C - 1], 1);
                int X[C];
                {
                    int xl = NEG_INF;
#pragma unroll
                    for (int k = 0; k < C; k++) {
                        int gen = ((k == 0) ? hprev : A[k - 1]) + c2;
                        if (PLAIN) {
                            if (k == 0 && lane == 0) gen = seed;
                        } else {
                            if (cbase + k < (int)left) gen = NEG_INF;
                            if (cbase + k == (int)left) gen = seed;
                        }
                        xl = max(xl + c1, gen);
                        X[k] = xl;
                    }
                }
                const int z = X[C - 1] - (cbase + C - 1) * c1;
                const int winc = warp_incl_max(z, lane);
                int wexc = __shfl_up_sync(FULL, winc, 1);
                if (lane == 0) wexc = NEG_INF;
                int xin = wexc + cbase * c1;
                // pass C
                unsigned xn_prev = 0;  // "x[c-1] > m[c-1] + o" of the previous cell
                unsigned xn_last = 0;
#pragma unroll
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    const int x = max(X[k], xin);
                    xin += c1;
                    const int dd = D[k];
                    int yv = B[k];
                    int ye = yv;
                    if (k == 0 && fc0) {
                        ye = NEG_INF;
                        yv = 0;
                    }
                    const int t = max(dd, x);
                    const int m = max(t, ye);
                    unsigned cd = (t < ye) ? (unsigned)DIR_U : ((dd < x) ? (unsigned)DIR_L : (unsigned)DIR_D);
                    if ((ybits >> k) & 1u) cd |= 8u;
                    if (k == 0 && fc0) cd = DIR_U | ((mps & SMASK) << (4 + SB));
                    if (k > 0 && xn_prev && (PLAIN || c > (int)left)) cd |= 4u;
                    xn_prev = (x > m + o) ? 1u : 0u;
                    if (k == C - 1) xn_last = xn_prev;
                    code[k] = cd;
                    const bool act = PLAIN ? (c < (int)right) : (c >= (int)left && c < (int)right);
                    A[k] = act ? m : NEG_INF;
                    B[k] = act ? yv : NEG_INF;
                    bestv = max(bestv, A[k]);
                }
                unsigned pl = __shfl_up_sync(FULL, xn_last, 1);
                if (lane != 0 && pl && (PLAIN || cbase > (int)left)) code[0] |= 4u;
                // right-most maximum inside the lane (only lanes holding the row maximum matter)
                {
                    const int lanemax = bestv;
                    unsigned eqb = 0;
#pragma unroll
                    for (int k = 0; k < C; k++)
                        if (A[k] == lanemax) eqb |= 1u << k;
                    bcol = cbase + (31 - __clz(eqb | 1u));
                }
            };
            if (g16) {
                row16();
                bcol16();
                slots16(mps, SD, SU);
#ifdef RG_ROWSTATS
                st_kind = 1;
#endif
            } else if (f16) {
                row16();
                bcol16();
#ifdef RG_ROWSTATS
                st_kind = 0;
#endif
            } else if (fast) {
                if (plain)
                    row32(std::true_type{});
                else
                    row32(std::false_type{});
#ifdef RG_ROWSTATS
                st_kind = plain ? 2 : 3;
#endif
            } else if (i == 0) {
                // gap_global_abpoa.rs:68-77
#pragma unroll
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    const bool act = c < (int)right;
                    int v = (c == 0) ? 0 : o + e * c;
                    A[k] = act ? v : NEG_INF;
                    B[k] = A[k];
                    code[k] = (c == 0) ? DIR_O : DIR_L;
                    if (act && v >= bestv) {
                        bestv = v;
                        bcol = c;
                    }
                }
            } else {
                // General path (segment starts, growing / shifted bands). Deliberately NOT unrolled: the per-lane
                // arrays live in (L1-resident) local memory and the loops are rolled, so that this rarely taken
                // path stays small in the instruction cache next to the unrolled fast path.
                int GA[C], GB[C], GD[C], GX[C];  // best pred m at c, best pred y at c, best pred m at c-1, in-lane x
                unsigned GS[C];                  // slots: um | uy << 8 | d << 16, later the cell's trace code
                unsigned pred0_slot = 0xffffffffu;
                if (!nwp) {
                    int up = __shfl_up_sync(FULL, A[C - 1], 1);
                    if (lane == 0) up = NEG_INF;
#pragma unroll
                    for (int k = 0; k < C; k++) {
                        GA[k] = A[k];
                        GB[k] = B[k];
                        GD[k] = (k == 0) ? up : A[k - 1];
                        GS[k] = 0;
                    }
                } else {
                    // gather over the predecessor rows (ring, absolute columns, NEG_INF outside their bands);
                    // first predecessor in list order wins ties (strict >), gap_global_abpoa.rs:266-345
#pragma unroll 1
                    for (int k = 0; k < C; k++) {
                        GA[k] = NEG_INF;
                        GB[k] = NEG_INF;
                        GD[k] = NEG_INF;
                        GS[k] = 0;
                    }
                    for (uint32_t q = 0; q < pe - pb; q++) {
                        const uint32_t p = g.pred_idx[pb + q];
                        if (p == 0) pred0_slot = q;
                        const int32_t* mp = ring_m + (size_t)(p & RM) * STRIDE + cbase;
                        const int32_t* yp = ring_y + (size_t)(p & RM) * STRIDE + cbase;
                        int carry = (lane == 0) ? NEG_INF : mp[-1];
#pragma unroll 1
                        for (int j = 0; j < C; j += 4) {
                            int4 mv = reinterpret_cast<const int4*>(mp)[j / 4];
                            int4 yv = reinterpret_cast<const int4*>(yp)[j / 4];
                            const int mm[4] = {mv.x, mv.y, mv.z, mv.w};
                            const int yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const int k = j + t;
                                unsigned sl = GS[k];
                                if (mm[t] > GA[k]) {
                                    GA[k] = mm[t];
                                    sl = (sl & ~0xffu) | q;
                                }
                                if (yy[t] > GB[k]) {
                                    GB[k] = yy[t];
                                    sl = (sl & ~0xff00u) | (q << 8);
                                }
                                if (carry > GD[k]) {
                                    GD[k] = carry;
                                    sl = (sl & ~0xff0000u) | (q << 16);
                                }
                                GS[k] = sl;
                                carry = mm[t];
                            }
                        }
                    }
                }
                const int fb0 = 2 * o + e * (best_p + 1);  // gap_global_abpoa.rs:117,139
                const int32_t* srow = s_sc + li * 8;
                const bool has_fc = (lane == 0) && left == 0;  // first column (gap_global_abpoa.rs:78-92)
                const int seed = (left == 0) ? o + e * (best_p + 1) : fb0 + e * (int)left;
                // ---- pass A (y, d) and pass B (in-lane x chain; generator of column c is h[c-1] + c2)
                int hprev;
                {
                    int hl = NEG_INF, xl = NEG_INF, hp = NEG_INF;
                    // the previous lane's last h is needed first: compute it for every lane up front
                    {
                        const int k = C - 1, c = cbase + k;
                        const int mu = GA[k], uy = GB[k];
                        const bool uav = mu > NEGH;
                        const int yv = uav ? max(mu + o, uy) + e : fb0 + e * c;
                        const unsigned rc = (c >= 1 && c < L) ? read[c - 1] : 4u;
                        const int dd = (GD[k] > NEGH) ? GD[k] + srow[rc] : NEG_INF;
                        hl = max(dd, yv);
                        if (c < (int)left || c >= (int)right || (k == 0 && has_fc)) hl = NEG_INF;
                    }
                    hprev = __shfl_up_sync(FULL, hl, 1);
                    if (lane == 0) hprev = NEG_INF;
                    hp = hprev;
#pragma unroll 1
                    for (int k = 0; k < C; k++) {
                        const int c = cbase + k;
                        const int mu = GA[k], uy = GB[k];
                        const bool uav = mu > NEGH;
                        const int um = mu + o;
                        const bool yf = uav && (uy > um);
                        const int yv = uav ? max(um, uy) + e : fb0 + e * c;
                        const unsigned sl = GS[k];
                        const unsigned us = uav ? (yf ? ((sl >> 8) & 0xffu) : (sl & 0xffu)) : mps;
                        const unsigned rc = (c >= 1 && c < L) ? read[c - 1] : 4u;
                        const int dv = GD[k];
                        const int dd = (dv > NEGH) ? dv + srow[rc] : NEG_INF;
                        GB[k] = yv;
                        GD[k] = dd;
                        GS[k] = (yf ? 8u : 0u) | (((sl >> 16) & SMASK) << 4) | ((us & SMASK) << (4 + SB));
                        int gen = (c > (int)left && c < (int)right && hp > NEGH) ? hp + c2 : NEG_INF;
                        if (c == (int)left) gen = seed;
                        xl = max(xl + c1, gen);
                        GX[k] = xl;
                        hp = (k == 0 && has_fc) ? NEG_INF : max(dd, yv);  // the first-column cell has m = x only
                    }
                }
                // ---- cross-lane max-plus scan: x entering lane t = max over earlier lanes
                int z = GX[C - 1] - (cbase + C - 1) * c1;
                if (GX[C - 1] <= NEGH) z = NEG_INF;
                int winc = warp_incl_max(z, lane);
                int wexc = __shfl_up_sync(FULL, winc, 1);
                if (lane == 0) wexc = NEG_INF;
                const int xin0 = (wexc > NEGH) ? wexc + cbase * c1 : NEG_INF;
                // ---- pass C: x, m, direction, flags
                unsigned xn_bits = 0;
#pragma unroll 1
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    const bool act = c >= (int)left && c < (int)right;
                    int x = GX[k];
                    if (xin0 > NEGH) x = max(x, xin0 + k * c1);
                    const bool fc = (k == 0) && has_fc;
                    const int dd = GD[k];
                    int yv = GB[k];
                    const int ye = fc ? NEG_INF : yv;
                    const int m = max(max(dd, x), ye);
                    unsigned cd = GS[k];
                    unsigned dir;
                    if (dd < x)
                        dir = (x < ye) ? DIR_U : DIR_L;
                    else
                        dir = (dd < ye) ? DIR_U : DIR_D;
                    if (fc) {
                        dir = DIR_U;
                        yv = 0;
                        cd = (mps & SMASK) << (4 + SB);
                    }
                    // gap_global_abpoa.rs:153-154: set_path_cell(u_pred, 'u') panics when u_pred == 0
                    if (nwp && act && !fc && dd > NEGH && dd < x && x < ye && ((cd >> (4 + SB)) & SMASK) == pred0_slot) status |= RG_READ_REF_PANIC;
                    GS[k] = cd | dir;
                    if (x > m + o) xn_bits |= 1u << k;
                    GA[k] = act ? m : NEG_INF;
                    GB[k] = act ? yv : NEG_INF;
                    if (act && m >= bestv) {
                        bestv = m;
                        bcol = c;
                    }
                }
                // path_x flag of column c: x[c-1] > m[c-1] + o (gap_global_abpoa.rs:358-364), only for c > left
                unsigned prev_last = __shfl_up_sync(FULL, xn_bits >> (C - 1), 1) & 1u;
                if (lane == 0) prev_last = 0;
                const unsigned xf = (xn_bits << 1) | prev_last;
#pragma unroll
                for (int k = 0; k < C; k++) {
                    const int c = cbase + k;
                    A[k] = GA[k];
                    B[k] = GB[k];
                    code[k] = GS[k] | ((c > (int)left && c < (int)right && ((xf >> k) & 1u)) ? 4u : 0u);
                }
            }
            // ---- row arg-max, right-most (>=)  (gap_global_abpoa.rs:198-203)
            const int tmax = __reduce_max_sync(FULL, bestv);
            prev_tmax = tmax;
            const unsigned eq = __ballot_sync(FULL, bestv == tmax && bestv > NEGH);
            const uint32_t row_bsp = (uint32_t)__shfl_sync(FULL, bcol, 31 - __clz(eq));
            // ---- stores: packed trace codes, ring copy for predecessor rows, row meta
            if (!have_planes) planes_from_codes<C, SB>(code, P, SD, SU);
            store_planes<C>(planes + (size_t)i * PlaneFmt<C>::ROWW + lane * NW, P);
            if (nwp) side_store(i, SD, SU);
            if (rep16) {
                rows16++;
                bool leave = (rf & RF_F_PRED) || i == n - 2;   // the end-cell candidates are read from the 32-bit row
                if ((rows16 & (PK_GUARD_ROWS - 1)) == 0 && guard16(tmax)) {
                    leave = true;
                    en16 = false;  // this read does not fit 16 bits: stay on the 32-bit paths
                }
                if (leave) {
                    unpack16(A, B);
                    rep16 = false;
                }
            }
            if (left != 0 || right != (uint32_t)L) last_nonfull = (int)i;
            if (rf & RF_IS_PRED) {
                if (rep16) {
                    ring_store16(i);
                } else {
                    store_row<C>(ring_m + (size_t)(i & RM) * STRIDE + cbase, A);
                    store_row<C>(ring_y + (size_t)(i & RM) * STRIDE + cbase, B);
                }
            }
            if (lane == 0) {
                RowMeta rm;
                rm.base = 0;
                rm.left = left;
                rm.right = right;
                rm.bsp = row_bsp;
                rowmeta[i] = rm;
            }
            __syncwarp();
#ifdef RG_ROWSTATS
            {
                const long long dt = clock64() - st_t0;
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (st_kind == q) {
                        st_cnt[q]++;
                        st_cyc[q] += dt;
                    }
            }
#endif
            prev_bsp = row_bsp;
            prev_left = left;
            prev_right = right;
            if ((rf & RF_F_PRED) || i == n - 2) {
                int lv = NEG_INF;
#pragma unroll
                for (int k = 0; k < C; k++)
                    if (cbase + k == (int)right - 1) lv = A[k];
                const int lastcell = __reduce_max_sync(FULL, lv);
                if ((rf & RF_F_PRED) && lastcell > best_end_val) {
                    best_end_val = lastcell;
                    best_end_row = i;
                }
                if (i == n - 2) last_val = lastcell;
            }
        }

        const long long t_dp = clock64();
        status = __reduce_or_sync(FULL, (unsigned)status);
        res.status = status;
        res.cells = cells;
        if (!abort_read && !(status & RG_READ_REF_PANIC)) {
            // end cell (gap_global_abpoa.rs:206-214): row n-2 unless an F predecessor is strictly better
            uint32_t last_row = n - 2;
            int best_value = last_val;
            if (best_end_val > last_val) {
                last_row = best_end_row;
                best_value = best_end_val;
            }
            RowMeta meta = rowmeta[last_row];
            uint32_t row = last_row, col = meta.right - 1;
            res.score = best_value;
            res.end_row = row;
            res.end_col = col;
            // ---- traceback (gaf_output.rs:96-253) fused with band_ampl_enough (gap_global_abpoa.rs:371-455)
            RunEmitter em;
            em.init(runs, ws.run_cap);
            PlaneTrace<C, SB> tr;
            tr.planes = planes;
            tr.side = side;
            tr.rowflags = g.rowflags;
            tr.nwp_ord = g.nwp_ord;
            const WalkOut wo = walk_affine<SB>(tr, rowmeta, g, read, L, row, col, em, lane);
            row = wo.row;
            col = wo.col;
            const int bandchk = wo.bandchk;
            const bool panic = wo.panic;
            em.flush(lane);
            if (panic) res.status |= RG_READ_REF_PANIC;
            if (bandchk == 0) res.status |= RG_READ_BAND_WARNING;
            if (em.overflow) res.status |= RG_READ_TRACE_OVERFLOW;
            res.start_row = row;
            res.start_col = col;
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
        {   // diagnostic: SM cycles spent in the forward pass / in traceback+publish (kilo-cycles)
            const long long t_end = clock64();
            res.fen = (uint32_t)((t_dp - t_start) >> 10);
            res.rsn = (uint32_t)((t_end - t_dp) >> 10);
#ifdef RG_ROWSTATS
            res.best_path = st_cnt[0];
            res.rev_best_path = st_cnt[1];
            res.rec_col = st_cnt[2];
            res.rev_end_row = st_cnt[3];
            res.displacement = (int32_t)st_cnt[4];
            res.score_f32 = (float)(st_cyc[0] >> 10);
            res.n_runs_rev = (uint32_t)(st_cyc[1] >> 10);
            res.end_col = (uint32_t)(st_cyc[2] >> 10);   // diagnostics build only: overwrites result fields
            res.start_row = (uint32_t)(st_cyc[3] >> 10);
            res.start_col = (uint32_t)(st_cyc[4] >> 10);
#endif
        }
        if (lane == 0) b.results[ridx] = res;
        __syncwarp();
    }
}

// Columns per lane for a batch whose longest read has Lmax columns (incl. '$'); 0 = use the striped kernel.
int gap_blk_cols(uint32_t Lmax) {
    if (Lmax <= 128) return 4;
    if (Lmax <= 256) return 8;
    if (Lmax <= 512) return 16;
    if (Lmax <= 1024) return 32;
    return 0;
}

// match/mismatch table of score_matrix.rs:35-66: M on the A,C,G,T diagonal, X elsewhere among A,C,G,T,N
static bool simple_scoring(const DevScoring& s) {
    for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++)
            if (s.sc[a][b] != ((a == b && a < 4) ? s.sc[0][0] : s.sc[0][1])) return false;
    return true;
}

template <int C>
static int launch_c(const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b, int trace_bytes,
                    int blocks, cudaStream_t st) {
    const bool simple = simple_scoring(s);
    const size_t smem = (size_t)BLK_WARPS * 5 * (C / 2) * 32 * sizeof(unsigned);
    const void* k = trace_bytes == 1 ? (simple ? (const void*)k_gap_global_blk<C, 2, true> : (const void*)k_gap_global_blk<C, 2, false>)
                                     : (simple ? (const void*)k_gap_global_blk<C, 6, true> : (const void*)k_gap_global_blk<C, 6, false>);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (trace_bytes == 1) {
        if (simple)
            k_gap_global_blk<C, 2, true><<<blocks, BLK_WARPS * 32, smem, st>>>(g, s, ws, b);
        else
            k_gap_global_blk<C, 2, false><<<blocks, BLK_WARPS * 32, smem, st>>>(g, s, ws, b);
    } else {
        if (simple)
            k_gap_global_blk<C, 6, true><<<blocks, BLK_WARPS * 32, smem, st>>>(g, s, ws, b);
        else
            k_gap_global_blk<C, 6, false><<<blocks, BLK_WARPS * 32, smem, st>>>(g, s, ws, b);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_gap_global_blk(int C, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
                          int trace_bytes, int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 4: return launch_c<4>(g, s, ws, b, trace_bytes, blocks, st);
        case 8: return launch_c<8>(g, s, ws, b, trace_bytes, blocks, st);
        case 16: return launch_c<16>(g, s, ws, b, trace_bytes, blocks, st);
        case 32: return launch_c<32>(g, s, ws, b, trace_bytes, blocks, st);
        default: return -2;
    }
}

template <int C>
static int occ_c(int trace_bytes, int* nb) {
    const void* k = trace_bytes == 1 ? (const void*)k_gap_global_blk<C, 2, true> : (const void*)k_gap_global_blk<C, 6, true>;
    const size_t smem = (size_t)BLK_WARPS * 5 * (C / 2) * 32 * sizeof(unsigned);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, BLK_WARPS * 32, smem) == cudaSuccess ? 0 : -1;
}
int gap_blk_warps_per_block() { return BLK_WARPS; }
int gap_blk_blocks_per_sm(int C, int trace_bytes, int* nb) {
    switch (C) {
        case 4: return occ_c<4>(trace_bytes, nb);
        case 8: return occ_c<8>(trace_bytes, nb);
        case 16: return occ_c<16>(trace_bytes, nb);
        case 32: return occ_c<32>(trace_bytes, nb);
        default: return -2;
    }
}

}  // namespace rg
