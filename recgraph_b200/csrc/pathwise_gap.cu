// Affine-gap pathwise alignment, the reference's experimental modes 6 (global, pathwise_alignment_gap.rs:4-574) and 7
// (semiglobal, pathwise_alignment_gap_semi.rs:5-473) with their CIGAR builders (pathwise_alignment_output.rs:186-451).
//
// These modes print a CIGAR line, not a GAF record, and their builders walk back by comparing RAW entries of the three
// delta-encoded tensors dpm / x / y (a non-alpha path's entry is its score minus the node's alpha path's) — including tests
// such as `dpm[i][j][best] < y[i][j][best]` on delta entries and `max == d` on scores without the substitution score. To be
// bit-exact the device therefore keeps the same three tensors in the same delta encoding, n x L x P each, per read in
// flight (the reference's own footprint), and the walk is the reference's walk on them.
//
// One CTA per read, one THREAD PER PATH: a thread owns its path's three entries of the current cell. Cells of a row are
// visited in order (the horizontal affine dependency runs through the leader's choice at every column); the values other
// threads need — the group leader's and the alpha path's entries of the previous column — go through shared memory, the
// predecessor row comes from the tensors (one coalesced line per cell). A row's paths are partitioned into groups, one
// per incoming edge, with the leader rule of pathwise_alignment_gap.rs:311-316,415-419 (same groups as modes 4/5).
// Not a throughput kernel: these modes are experimental in the reference and none of BASELINE's configurations uses them.
#include <cuda_runtime.h>

#include <climits>

#include "device.h"
#include "poa_common.cuh"

namespace rg {

namespace {

struct GapTensors {
    int32_t* D;   // dpm
    int32_t* X;
    int32_t* Y;
    uint32_t L, Pp;
    __device__ __forceinline__ size_t at(uint32_t i, uint32_t j) const { return ((size_t)i * L + j) * Pp; }
};

__device__ __forceinline__ bool has_path(const DevPathGraph& g, uint32_t row, uint32_t q) {
    return q < g.P && ((g.node_bits[(size_t)row * g.PW + q / 32] >> (q % 32)) & 1u);
}
// raw tensor entry as the reference sees it: slots of paths that do not go through the node are never written (zero)
__device__ __forceinline__ int raw(const DevPathGraph& g, const int32_t* T, const GapTensors& t, uint32_t i, uint32_t j, uint32_t q) {
    return has_path(g, i, q) ? T[t.at(i, j) + q] : 0;
}
__device__ __forceinline__ int abs_at(const DevPathGraph& g, const GapTensors& t, uint32_t i, uint32_t j, uint32_t bp) {
    const uint32_t a = g.alphas[i];
    return a == bp ? raw(g, t.D, t, i, j, bp) : raw(g, t.D, t, i, j, bp) + raw(g, t.D, t, i, j, a);
}
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(a, max(b, c)); }

}  // namespace

__global__ void __launch_bounds__(128, 4) k_pathwise_gap(DevPathGraph g, DevScoring sc, PwGapWorkspace ws, PoaBatch b, int mode) {
    __shared__ int32_t s_sc[48];
    __shared__ int s_cd[2][128], s_cx[2][128];      // current row, previous column: dpm / x of every path (final deltas)
    __shared__ int s_td[128], s_tx[128], s_ty[128];  // this cell before the alpha fix-up
    __shared__ unsigned long long s_ticket;
    const uint32_t q = threadIdx.x;
    const uint32_t slot = blockIdx.x;
    if (q < 48) s_sc[q] = (&sc.sc[0][0])[q];
    const uint32_t n = g.n, P = g.P, PW = g.PW, Pp = ws.Pp;
    const bool semi = mode == RG_MODE_PATHWISE_GAP_SEMIGLOBAL;
    const int o = sc.o, e = sc.e;
    rg_run* runs = ws.runs + (size_t)slot * ws.run_cap;
    __syncthreads();

    for (;;) {
        if (q == 0) s_ticket = atomicAdd(&b.counters[0], 1ull);
        __syncthreads();
        const unsigned long long ticket = s_ticket;
        __syncthreads();
        if (ticket >= (unsigned long long)b.n_reads) break;
        const int ridx = b.order ? b.order[ticket] : (int)ticket;
        const uint8_t* read = b.reads + b.read_off[ridx];
        const uint32_t L = (uint32_t)(b.read_off[ridx + 1] - b.read_off[ridx]) + 1;
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
        if (L > ws.Lp) {
            res.status = RG_READ_TRACE_OVERFLOW;
            if (q == 0) b.results[ridx] = res;
            continue;
        }
        GapTensors t;
        t.L = L;
        t.Pp = Pp;
        t.D = ws.T + (size_t)slot * 3 * n * ws.Lp * Pp;
        t.X = t.D + (size_t)n * ws.Lp * Pp;
        t.Y = t.X + (size_t)n * ws.Lp * Pp;

        // ---- row 0 (…_gap.rs:23-34): alphas[0] = 0 carries o + e * j in y and dpm, every other path stays 0
        for (uint32_t j = 0; j < L; j++) {
            const int v = (q == g.alphas[0] && j > 0) ? o + e * (int)j : 0;
            t.D[t.at(0, j) + q] = v;
            t.Y[t.at(0, j) + q] = v;
            t.X[t.at(0, j) + q] = 0;
        }
        __syncthreads();

        for (uint32_t i = 1; i + 1 < n; i++) {
            const bool memb = has_path(g, i, q);
            const uint32_t ai = g.alphas[i];
            const int32_t* srow = s_sc + g.lnz[i] * 8;
            // my group: predecessor, leader, and whether the leader is the predecessor's alpha path (case A) or a path
            // chosen because that alpha does not continue into this node (case B, "set new alpha")
            uint32_t p = 0, a = 0, ap = 0;
            bool caseA = true;
            if (memb)
                for (uint32_t gi = g.grp_off[i]; gi < g.grp_off[i + 1]; gi++)
                    if ((g.grp_mask[(size_t)gi * PW + q / 32] >> (q % 32)) & 1u) {
                        const PwGroup gr = g.grp[gi];
                        p = gr.pred;
                        a = gr.leader;
                        ap = g.alphas[p];
                        caseA = a == ap;
                    }
            const bool fix = memb && a != ai;   // "remove multiple alpha" (…_gap.rs:133-147,520-538)
            // ---- column 0
            {
                int xv = 0;
                if (memb && !semi) {
                    if (caseA)
                        xv = (q == a) ? ((p == 0) ? o + e : t.X[t.at(p, 0) + a] + e) : t.X[t.at(p, 0) + q];
                    else
                        xv = (q == a) ? ((p == 0) ? o + e : t.X[t.at(p, 0) + a] + t.X[t.at(p, 0) + ap] + e)
                                      : t.X[t.at(p, 0) + q] - t.X[t.at(p, 0) + a];
                }
                s_tx[q] = xv;
                __syncthreads();
                if (fix) {
                    const int va = s_tx[a] - s_tx[ai];
                    xv = (q == a) ? va : xv + va;
                }
                if (memb) {
                    t.X[t.at(i, 0) + q] = xv;
                    t.D[t.at(i, 0) + q] = xv;
                    t.Y[t.at(i, 0) + q] = 0;
                }
                s_cd[0][q] = memb ? xv : 0;
                s_cx[0][q] = memb ? xv : 0;
                __syncthreads();
            }
            // ---- columns 1 .. L-1
            int dq_prev = memb ? t.D[t.at(p, 0) + q] : 0;         // dpm[p][j-1][q]
            int da_prev = memb ? t.D[t.at(p, 0) + ap] : 0;        // dpm[p][j-1][alphas[p]]
            int dt_prev = (memb && !caseA) ? t.D[t.at(p, 0) + a] : 0;
            for (uint32_t j = 1; j < L; j++) {
                const int cb = (j - 1) & 1, nb = j & 1;
                int dn = 0, xn = 0, yn = 0;
                int dq_cur = 0, da_cur = 0, dt_cur = 0;
                if (memb) {
                    const size_t pj = t.at(p, j);
                    dq_cur = t.D[pj + q];
                    da_cur = t.D[pj + ap];
                    const int yq = t.Y[pj + q], ya = t.Y[pj + ap];
                    const int sub = srow[read[j - 1]];
                    const bool same_a = a == ai;
                    const int cxa = s_cx[cb][a], cda = s_cd[cb][a], cxi = s_cx[cb][ai], cdi = s_cd[cb][ai];
                    const int cxq = s_cx[cb][q], cdq = s_cd[cb][q];
                    int u_y, u_dpm, l_x, l_dpm, d;
                    if (caseA) {
                        u_y = ya + e;
                        u_dpm = da_cur + o + e;
                        l_x = same_a ? cxa + e : cxa + cxi + e;
                        l_dpm = same_a ? cda + o + e : cdi + cda + o + e;
                        d = da_prev + sub;
                    } else {
                        dt_cur = t.D[pj + a];
                        const int yt = t.Y[pj + a];
                        u_y = ya + yt + e;
                        u_dpm = da_cur + dt_cur + o + e;
                        l_x = same_a ? cxi + e : cxi + cxa + e;
                        l_dpm = same_a ? cdi + o + e : cdi + cda + o + e;
                        d = da_prev + dt_prev + sub;
                    }
                    const bool ysrc_dpm = u_dpm >= u_y, xsrc_dpm = l_dpm >= l_x;
                    const int u = ysrc_dpm ? u_dpm : u_y, l = xsrc_dpm ? l_dpm : l_x;
                    const int best = max3i(d, u, l);
                    if (q == a) {
                        yn = u;
                        xn = l;
                        dn = best;
                    } else {
                        if (caseA) {
                            yn = ysrc_dpm ? dq_cur : yq;
                        } else {
                            const int yt = t.Y[pj + a];
                            yn = ysrc_dpm ? dq_cur - dt_cur : yq - yt;
                        }
                        xn = xsrc_dpm ? (same_a ? cdq : cdq - cda) : (same_a ? cxq : cxq - cxa);
                        if (best == d)
                            dn = caseA ? dq_prev : dq_prev - dt_prev;
                        else if (best == u)
                            dn = yn;
                        else
                            dn = xn;
                    }
                }
                s_td[q] = dn;
                s_tx[q] = xn;
                s_ty[q] = yn;
                __syncthreads();
                if (fix) {
                    const int vd = s_td[a] - s_td[ai], vx = s_tx[a] - s_tx[ai], vy = s_ty[a] - s_ty[ai];
                    dn = (q == a) ? vd : dn + vd;
                    xn = (q == a) ? vx : xn + vx;
                    yn = (q == a) ? vy : yn + vy;
                }
                if (memb) {
                    const size_t ij = t.at(i, j) + q;
                    t.D[ij] = dn;
                    t.X[ij] = xn;
                    t.Y[ij] = yn;
                }
                s_cd[nb][q] = memb ? dn : 0;
                s_cx[nb][q] = memb ? xn : 0;
                dq_prev = dq_cur;
                da_prev = da_cur;
                dt_prev = dt_cur;
                __syncthreads();
            }
        }

        // ---- results, walk and publish (thread 0; the tensors are complete)
        if (q == 0) {
            uint32_t best_path = 0, ending = 0;
            int best_score = 0;
            if (!semi) {
                // …_gap.rs:543-562: per path the absolute score at its own last node, then max of (score, path)
                int bs = 0;
                bool first = true;
                for (uint32_t path = 0; path < P; path++) {
                    int r = 0;
                    for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                        if ((g.grp_mask[(size_t)fg * PW + path / 32] >> (path % 32)) & 1u) {
                            const uint32_t pred = g.grp[fg].pred;
                            r = (path == g.alphas[pred]) ? t.D[t.at(pred, L - 1) + path]
                                                         : t.D[t.at(pred, L - 1) + path] + t.D[t.at(pred, L - 1) + g.alphas[pred]];
                        }
                    if (first || r >= bs) {
                        first = false;
                        bs = r;
                        best_path = path;
                    }
                }
                best_score = bs;
                for (uint32_t fg = g.grp_off[n - 1]; fg < g.grp_off[n]; fg++)
                    if ((g.grp_mask[(size_t)fg * PW + best_path / 32] >> (best_path % 32)) & 1u) ending = g.grp[fg].pred;
            } else {
                // best_ending_node (…_gap_semi.rs:447-473): every one of the P slots competes (slots of paths that do not go
                // through the node hold 0), highest path id on ties; a row replaces the incumbent only if strictly better
                bool have = false;
                int mx = 0;
                for (uint32_t i = 0; i + 1 < n; i++) {
                    const uint32_t al = g.alphas[i];
                    const int av = raw(g, t.D, t, i, L - 1, al);
                    int bv = 0;
                    uint32_t bp = 0;
                    for (uint32_t path = 0; path < P; path++) {
                        int v = raw(g, t.D, t, i, L - 1, path);
                        if (has_path(g, i, path) && path != al) v += av;
                        if (path == 0 || v >= bv) {
                            bv = v;
                            bp = path;
                        }
                    }
                    if (!have || bv > mx) {
                        have = true;
                        mx = bv;
                        ending = i;
                        best_path = bp;
                    }
                }
                best_score = mx;
            }
            res.score = best_score;
            res.best_path = best_path;
            res.end_row = ending;
            res.end_col = L - 1;
            // ---- build_alignment_gap / build_alignment_semiglobal_gap (pathwise_alignment_output.rs:186-451)
            RunEmitter em;
            em.init(runs, ws.run_cap);
            uint32_t i = ending, j = L - 1;
            bool panic = false;
            uint64_t guard = 0;
            const uint64_t guard_max = 8ull * ((uint64_t)n + L) + 64;
            auto pred_of = [&](uint32_t row, bool& found) -> uint32_t {
                uint32_t pr = 0;
                found = false;
                for (uint32_t gi = g.grp_off[row]; gi < g.grp_off[row + 1]; gi++)
                    if ((g.grp_mask[(size_t)gi * PW + best_path / 32] >> (best_path % 32)) & 1u) {
                        pr = g.grp[gi].pred;
                        found = true;
                    }
                return pr;
            };
            while (i != 0 && j != 0 && !panic) {
                if (++guard > guard_max) {
                    panic = true;
                    break;
                }
                const int curr = abs_at(g, t, i, j, best_path);
                bool has_pred = false;
                uint32_t predecessor = 0;
                int d, u, l;
                if (!g.nwp[i]) {
                    d = abs_at(g, t, i - 1, j - 1, best_path);
                    u = abs_at(g, t, i - 1, j, best_path);
                    l = abs_at(g, t, i, j - 1, best_path);
                } else {
                    d = u = l = 0;
                    predecessor = pred_of(i, has_pred);
                    if (has_pred) {
                        d = abs_at(g, t, predecessor, j - 1, best_path);
                        u = abs_at(g, t, predecessor, j, best_path);
                        l = abs_at(g, t, i, j - 1, best_path);
                    }
                }
                const int mx = max3i(d, u, l);
                if (mx == d) {
                    em.step(curr < d ? RG_OP_d : RG_OP_D, i, 0);
                    i = has_pred ? predecessor : i - 1;
                    j -= 1;
                } else if (mx == u) {
                    em.step(RG_OP_U, i, 0);
                    i = has_pred ? predecessor : i - 1;
                    while (raw(g, t.D, t, i, j, best_path) < raw(g, t.Y, t, i, j, best_path)) {   // raw delta entries, as in the reference
                        if (++guard > guard_max) {
                            panic = true;
                            break;
                        }
                        em.step(RG_OP_U, i, 0);
                        if (g.nwp[i]) {
                            bool f2;
                            const uint32_t pr = pred_of(i, f2);
                            if (f2) {
                                has_pred = true;
                                predecessor = pr;
                            }
                        } else {
                            if (i == 0) {
                                panic = true;   // attempt to subtract with overflow
                                break;
                            }
                            has_pred = true;
                            predecessor = i - 1;
                        }
                        if (!has_pred) {
                            panic = true;   // predecessor.unwrap() on None
                            break;
                        }
                        i = predecessor;
                    }
                } else {
                    em.step(RG_OP_L, i, 0);
                    j -= 1;
                    while (raw(g, t.D, t, i, j, best_path) < raw(g, t.X, t, i, j, best_path)) {
                        if (++guard > guard_max || j == 0) {
                            panic = true;
                            break;
                        }
                        em.step(RG_OP_L, i, 0);
                        j -= 1;
                    }
                }
            }
            while (j > 0 && !panic) {
                em.step(RG_OP_L, i, 0);
                j -= 1;
            }
            auto to_source = [&](uint32_t row) -> uint32_t {   // pathwise_alignment_output.rs:421-447: steps back to row 0
                uint32_t steps = 0;
                while (row > 0) {
                    if (g.nwp[row]) {
                        bool f2;
                        const uint32_t pr = pred_of(row, f2);
                        if (f2) row = pr;
                    } else {
                        row -= 1;
                    }
                    if (++steps > n + 1) {
                        panic = true;
                        break;
                    }
                }
                return steps;
            };
            if (!semi) {
                // `while i > 0 { cigar.push('U'); i -= 1 }` (…_output.rs:300-303), then the reversed list loses its last element
                if (i > 0) em.bulk(RG_OP_UPAD, i, i, 0);
                res.start_row = 0;
            } else {
                res.start_row = to_source(i);     // starting_node
                res.rev_end_row = to_source(ending);   // final_node
            }
            em.flush(0);
            if (panic) res.status |= RG_READ_REF_PANIC;
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

int launch_pathwise_gap(int mode, const DevPathGraph& g, const DevScoring& s, const PwGapWorkspace& ws, const PoaBatch& b, int blocks,
                        void* stream) {
    k_pathwise_gap<<<blocks, ws.Pp, 0, (cudaStream_t)stream>>>(g, s, ws, b, mode);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace rg
