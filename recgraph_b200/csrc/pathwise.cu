// Pathwise alignment, modes 4 / 5 (pathwise_alignment.rs:5-340, pathwise_alignment_semiglobal.rs:6-277) and their
// traceback (pathwise_alignment_output.rs:7-184), in the exact ABSOLUTE-score form of SURVEY §3.4:
// per row and per incoming edge ("group") the LEADER path does a linear-gap DP over the read, every other member
// path copies the leader's move (D / U / L) applied to its own scores. This is what the reference's delta-encoded
// tensor computes (verified against the literal restatement in oracle/pathwise.cpp).
//
// One CTA per read. Scores live in an L2-resident ring of rows, layout [row][column][path] so that a warp whose
// lanes are paths reads and writes 128-byte lines. Per row:
//   phase 1 (per group, all threads over column blocks): leader candidates, CTA-wide max-plus scan for the
//           horizontal dependency, one move byte per column in shared memory;
//   phase 2 (warps over column chunks, lanes = paths): members apply the move; each path also records its OWN
//           arg-max (2 bits per path-cell, stored as two ballot bit-planes) — that is what build_alignment
//           re-derives from the stored scores when it walks the best path back.
#include <cuda_runtime.h>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

constexpr int PT = 256;  // threads per CTA
enum { MV_D = 1, MV_U = 2, MV_L = 3 };

__device__ __forceinline__ int block_excl_max(int z, int tid, int* s_w) {
    // exclusive prefix max over the CTA's threads (identity NEG_INF)
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

__global__ void __launch_bounds__(PT) k_pathwise(DevPathGraph g, DevScoring sc, PwWorkspace ws, PoaBatch b, int mode) {
    extern __shared__ unsigned char s_dyn[];
    __shared__ int32_t s_sc[48];
    __shared__ int s_w[PT / 32];
    __shared__ unsigned long long s_ticket;
    __shared__ int s_best_val, s_best_set;
    __shared__ uint32_t s_best_row, s_best_path;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slot = blockIdx.x;
    if (tid < 48) s_sc[tid] = (&sc.sc[0][0])[tid];
    const uint32_t n = g.n, P = g.P, PW = g.PW, Lp = ws.Lp, Pp = ws.Pp, RM = g.ring - 1;
    // dynamic smem: moves [max_groups][Lp] bytes | du [Lp] ints | results [Pp] ints | ending [Pp] uints
    unsigned char* s_mv = s_dyn;
    int32_t* s_du = reinterpret_cast<int32_t*>(s_dyn + (((size_t)g.max_groups * Lp + 15) & ~(size_t)15));
    int32_t* s_res = s_du + Lp;
    uint32_t* s_end = reinterpret_cast<uint32_t*>(s_res + Pp);
    int32_t* S = ws.S + (size_t)slot * g.ring * Lp * Pp;
    int32_t* lead = ws.lead + (size_t)slot * g.ring * Lp;
    uint32_t* trace = ws.trace + (size_t)slot * n * Lp * PW * 2;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    const bool global_mode = (mode == RG_MODE_PATHWISE_GLOBAL);
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
        res.cells = (uint64_t)(n - 2) * (uint64_t)(L - 1);
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
            s_res[q] = 0;
            s_end[q] = 0;
        }
        // ---- row 0: every path carries the accumulated read gaps (pathwise_alignment_semiglobal.rs:26-32)
        for (uint32_t idx = tid; idx < (uint32_t)L * Pp; idx += PT) {
            const uint32_t j = idx / Pp, q = idx % Pp;
            S[(size_t)j * Pp + q] = (q < P) ? (int)j * g_rd : 0;
        }
        for (int j = tid; j < L; j += PT) lead[j] = j * g_rd;
        __syncthreads();

        const int Cc = (L - 1 + PT - 1) / PT;        // phase 1: columns per thread
        const int chunk = (L + PT / 32 - 1) / (PT / 32);  // phase 2: columns per warp

        for (uint32_t i = 1; i + 1 < n; i++) {
            const uint32_t g0 = g.grp_off[i], g1 = g.grp_off[i + 1];
            const int li = g.lnz[i];
            const int32_t* srow = s_sc + li * 8;
            const uint32_t alpha_i = g.alphas[i];
            int32_t* Si = S + (size_t)(i & RM) * Lp * Pp;
            int32_t* lead_i = lead + (size_t)(i & RM) * Lp;
            // ================= phase 1: the leader's DP of every group =================
            for (uint32_t gi = g0; gi < g1; gi++) {
                const PwGroup gr = g.grp[gi];
                unsigned char* mv = s_mv + (size_t)(gi - g0) * Lp;
                const int32_t* lp = gr.lead_is_alpha_of_pred ? lead + (size_t)(gr.pred & RM) * Lp : nullptr;
                const int32_t* Sp = S + (size_t)(gr.pred & RM) * Lp * Pp + gr.leader;
                const int m0 = global_mode ? (lp ? lp[0] : Sp[0]) + g_gr : 0;  // column 0 of this row for the leader
                const int jb = 1 + tid * Cc, je = min(L, jb + Cc);
                int v = NEG_INF;
                int pl = (jb < L) ? (lp ? lp[jb - 1] : Sp[(size_t)(jb - 1) * Pp]) : 0;
                for (int j = jb; j < je; j++) {
                    const int pc = lp ? lp[j] : Sp[(size_t)j * Pp];
                    const int d = pl + srow[read[j - 1]];
                    const int u = pc + g_gr;
                    const int du = max(d, u);
                    s_du[j] = du;
                    mv[j] = (d >= u) ? MV_D : MV_U;  // equality tests in the order d, u (pathwise_alignment_semiglobal.rs:46-57)
                    int gen = du;
                    if (j == 1) gen = max(du, m0 + g_rd);
                    v = max(v + g_rd, gen);
                    pl = pc;
                }
                int z = (jb < je && v > NEG_INF / 2) ? v - (je - 1) * g_rd : NEG_INF;
                const int wexc = block_excl_max(z, tid, s_w);
                int lcand = (wexc > NEG_INF / 2) ? wexc + jb * g_rd : NEG_INF;  // m[jb-1] + g_rd
                const bool own_alpha = gr.leader == alpha_i;
                for (int j = jb; j < je; j++) {
                    if (j == 1) lcand = m0 + g_rd;
                    const int du = s_du[j];
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
            // ================= phase 2: members apply their leader's move =================
            {
                const int jb = warp * chunk, je = min(L, jb + chunk);
                int row_best = NEG_INF;
                uint32_t row_path = 0;
                for (uint32_t pass = 0; pass < PW; pass++) {
                    const uint32_t q = pass * 32 + lane;
                    // my group
                    int gq = -1;
                    uint32_t pq = 0;
                    if (q < P)
                        for (uint32_t gi = g0; gi < g1; gi++)
                            if ((g.grp_mask[(size_t)gi * PW + pass] >> lane) & 1u) {
                                gq = (int)(gi - g0);
                                pq = g.grp[gi].pred;
                            }
                    const bool member = gq >= 0;
                    const int32_t* Sp = S + (size_t)(pq & RM) * Lp * Pp + q;
                    const unsigned char* mv = s_mv + (size_t)(member ? gq : 0) * Lp;
                    const int col0 = (member && global_mode) ? Sp[0] + g_gr : 0;
                    int prev_new = 0, sp_prev = 0;
                    if (member && jb > 0 && jb < je) {
                        // value of column jb-1 of this row: walk the L-run back to its anchor (another warp owns it)
                        int j0 = jb - 1;
                        while (j0 >= 1 && mv[j0] == MV_L) j0--;
                        int base;
                        if (j0 == 0)
                            base = col0;
                        else if (mv[j0] == MV_D)
                            base = Sp[(size_t)(j0 - 1) * Pp] + srow[read[j0 - 1]];
                        else
                            base = Sp[(size_t)j0 * Pp] + g_gr;
                        prev_new = base + (jb - 1 - j0) * g_rd;
                        sp_prev = Sp[(size_t)(jb - 1) * Pp];
                    }
                    for (int j = jb; j < je; j++) {
                        int nv = 0;
                        unsigned code = 0;
                        if (member) {
                            const int sp = Sp[(size_t)j * Pp];
                            if (j == 0) {
                                nv = col0;
                            } else {
                                const int sj = srow[read[j - 1]];
                                const int dq = sp_prev + sj, uq = sp + g_gr, lq = prev_new + g_rd;
                                const unsigned m = mv[j];
                                nv = (m == MV_D) ? dq : ((m == MV_U) ? uq : lq);
                                // own arg-max in build_alignment's order: d, then u, else l
                                const int bq = max(dq, max(uq, lq));
                                code = (bq == dq) ? MV_D : ((bq == uq) ? MV_U : MV_L);
                            }
                            sp_prev = sp;
                            prev_new = nv;
                        }
                        Si[(size_t)j * Pp + q] = nv;
                        const unsigned p0 = __ballot_sync(FULL, code & 1u), p1 = __ballot_sync(FULL, code & 2u);
                        if (lane == 0)
                            reinterpret_cast<uint2*>(trace)[((size_t)i * Lp + j) * PW + pass] = make_uint2(p0, p1);
                        if (j == L - 1) {
                            // candidates of the last column (best_ending_node / results of mode 4)
                            const int cand = member ? nv : NEG_INF;
                            const int mx = __reduce_max_sync(FULL, cand);
                            const unsigned eq = __ballot_sync(FULL, member && cand == mx);
                            if (eq && mx > row_best) {  // first strict maximum in path order
                                row_best = mx;
                                row_path = pass * 32 + (__ffs(eq) - 1);
                            }
                            if (global_mode && member) {
                                // pathwise_alignment.rs:305-319: paths whose last node this row is
                                for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                                    if (g.grp[fg].pred == i && ((g.grp_mask[(size_t)fg * PW + pass] >> lane) & 1u)) {
                                        s_res[q] = nv;
                                        s_end[q] = i;
                                    }
                            }
                        }
                    }
                }
                if (!global_mode && je == L && jb < je && lane == 0 && row_best > NEG_INF / 2) {
                    // pathwise_alignment_semiglobal.rs:269-273: a row replaces the incumbent only if strictly better
                    if (!s_best_set || row_best > s_best_val) {
                        s_best_set = 1;
                        s_best_val = row_best;
                        s_best_row = i;
                        s_best_path = row_path;
                    }
                }
            }
            __syncthreads();
        }

        // ================= end cell, traceback (thread 0), publish =================
        if (tid == 0) {
            uint32_t best_path, ending;
            int score;
            if (global_mode) {
                // max of (score, path): highest path id wins ties (pathwise_alignment.rs:320-325)
                best_path = 0;
                for (uint32_t q = 1; q < P; q++)
                    if (s_res[q] >= s_res[best_path]) best_path = q;
                ending = s_end[best_path];
                score = s_res[best_path];
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
            const uint32_t bw = best_path / 32, bb = best_path % 32;
            while (ii > 0 && j > 0) {
                const uint2 pl = reinterpret_cast<const uint2*>(trace)[((size_t)ii * Lp + j) * PW + bw];
                const unsigned code = ((pl.x >> bb) & 1u) | (((pl.y >> bb) & 1u) << 1);
                uint32_t pred = ii - 1;
                if (code != MV_L)
                    for (uint32_t gi = g.grp_off[ii]; gi < g.grp_off[ii + 1]; gi++)
                        if ((g.grp_mask[(size_t)gi * PW + bw] >> bb) & 1u) pred = g.grp[gi].pred;
                if (code == MV_D) {
                    em.step(g.lnz[ii] != read[j - 1] ? RG_OP_d : RG_OP_D, ii, 0);
                    ii = pred;
                    j--;
                } else if (code == MV_U) {
                    em.step(RG_OP_U, ii, 0);
                    ii = pred;
                } else {
                    em.step(RG_OP_L, ii, 0);
                    j--;
                }
            }
            while (j > 0) {  // pathwise_alignment_output.rs:111-114
                em.step(RG_OP_L, ii, 0);
                j--;
            }
            if (global_mode) {
                while (ii > 0) {  // :116-138
                    uint32_t pred = 0;
                    if (!g.nwp[ii])
                        pred = ii - 1;
                    else
                        for (uint32_t gi = g.grp_off[ii]; gi < g.grp_off[ii + 1]; gi++)
                            if ((g.grp_mask[(size_t)gi * PW + bw] >> bb) & 1u) pred = g.grp[gi].pred;
                    em.step(RG_OP_U, ii, 0);
                    ii = pred;
                }
            }
            em.flush(0);
            res.start_row = ii;
            res.start_col = 0;
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
    }
}

size_t pathwise_smem_bytes(const DevPathGraph& g, const PwWorkspace& ws) {
    return (((size_t)g.max_groups * ws.Lp + 15) & ~(size_t)15) + (size_t)ws.Lp * 4 + (size_t)ws.Pp * 8;
}

int launch_pathwise(int mode, const DevPathGraph& g, const DevScoring& s, const PwWorkspace& ws, const PoaBatch& b,
                    int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = pathwise_smem_bytes(g, ws);
    if (smem > 200 * 1024) return -3;
    if (cudaFuncSetAttribute(k_pathwise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    k_pathwise<<<blocks, PT, smem, st>>>(g, s, ws, b, mode);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int pathwise_blocks_per_sm(const DevPathGraph& g, const PwWorkspace& ws, int* nb) {
    const size_t smem = pathwise_smem_bytes(g, ws);
    if (smem > 200 * 1024) return -3;
    if (cudaFuncSetAttribute(k_pathwise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k_pathwise, PT, smem) == cudaSuccess ? 0 : -1;
}

}  // namespace rg
