// C-ABI implementation (include/recgraph_b200.h): context, graph upload, batch alignment, result fetch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "device.h"
#include "host.h"

using namespace rg;

namespace {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    bool ensure(size_t n) {
        if (n <= cap && p) return true;
        release();
        if (n == 0) n = 1;
        if (cudaMalloc((void**)&p, n * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return false;
        }
        cap = n;
        return true;
    }
    bool upload(const std::vector<T>& v, cudaStream_t st) {
        if (!ensure(v.size())) return false;
        if (v.empty()) return true;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
};
template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { release(); }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    bool ensure(size_t n) {
        if (n <= cap && p) return true;
        release();
        if (n == 0) n = 1;
        if (cudaMallocHost((void**)&p, n * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return false;
        }
        cap = n;
        return true;
    }
};

}  // namespace

struct rg_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string last_error;
    // graph
    bool has_graph = false;
    FlatGraph fg;
    DevBuf<uint8_t> d_lnz, d_rowflags, d_min_pred_slot, d_prev_slot;
    DevBuf<uint32_t> d_pred_off, d_pred_idx, d_min_pred, d_nwp_ord;
    uint32_t n_gather = 0;
    DevBuf<int32_t> d_r_values;
    DevBuf<RowInfo> d_rowinfo;
    DevGraph dg{};
    // pathwise graph (forward)
    DevBuf<uint32_t> d_alphas, d_node_bits, d_grp_off, d_grp_mask;
    DevBuf<PwGroup> d_grp;
    DevBuf<uint8_t> d_pw_nwp;
    DevPathGraph dpg{}, dpg_rev{};
    DevBuf<uint32_t> d_rgrp_off, d_rgrp_mask;
    DevBuf<PwGroup> d_rgrp;
    DevBuf<uint8_t> d_rv_nwp;
    DevBuf<uint64_t> d_seg;
    DevBuf<int32_t> d_dfs, d_dfe;
    DevBuf<int32_t> d_rS, d_rLead, d_lastcol;
    DevBuf<uint32_t> d_rTrace;
    DevBuf<int2> d_fm, d_rwbest;
    bool has_path_graph = false;
    std::string path_graph_error;
    // score-transport kernel (pathwise_tr.cu): per-row records of both directions and its work-space
    DevBuf<PwtRow> d_pwt_rows, d_pwt_rrows;
    DevBuf<int32_t> d_nonmem_hi;
    DevBuf<int32_t> d_tr_tables, d_tr_ring_lead, d_tr_lastcol, d_tr_colmax, d_tr_stage;
    DevBuf<uint16_t> d_tr_ring_org;
    DevBuf<uint4> d_tr_ring_meta;
    DevBuf<uint8_t> d_tr_mv_f, d_tr_mv_r, d_tr_own;
    DevBuf<uint32_t> d_tr_own_pred;
    DevBuf<int2> d_tr_cb_f, d_tr_cb_r;
    DevBuf<int32_t> d_gapT;
    uint64_t tr_sig = 0;     // shape of the transport kernel's current work-space (mode, column class, batch size, rows)
    uint32_t tr_slots = 0;
    bool pw_v1 = getenv("RG_PW_V1") != nullptr;   // testing: the per-path kernel of round 1 (pathwise.cu) for A/B runs
    DevBuf<int32_t> d_pwS, d_pwLead;
    DevBuf<uint32_t> d_pwTrace;
    // scoring
    bool has_scoring = false;
    rg_scoring scoring{};
    DevScoring ds{};
    // staged reads
    int32_t n_reads = 0;
    uint64_t total_len = 0;
    uint32_t max_len = 0;
    std::vector<uint64_t> h_off;
    DevBuf<uint8_t> d_reads;
    DevBuf<uint64_t> d_read_off;
    DevBuf<int32_t> d_order;
    // workspace
    DevBuf<RowMeta> d_rowmeta;
    DevBuf<int32_t> d_ring_m, d_ring_y;
    DevBuf<uint8_t> d_trace;
    DevBuf<rg_run> d_slot_runs, d_out_runs;
    DevBuf<rg_read_result> d_results;
    DevBuf<unsigned long long> d_counters;
    // results on the host
    PinnedBuf<rg_read_result> h_results;
    PinnedBuf<rg_run> h_runs;
    uint64_t n_runs_total = 0;
    int last_mode = -1;
    bool results_ready = false;
    double kernel_ms = 0;
    uint64_t launches = 0, cells = 0;
    uint32_t slots_used = 0;
    bool no_s16 = getenv("RG_NO_S16") != nullptr;                // testing: 32-bit paths only
    bool force_striped = getenv("RG_FORCE_STRIPED") != nullptr;  // testing: keep the generic striped kernel
    PinnedBuf<unsigned long long> h_counters;

    int fail(int code, const std::string& msg) {
        last_error = msg;
        return code;
    }
    int cuda_fail(const char* what) {
        cudaError_t e = cudaGetLastError();
        last_error = std::string(what) + ": " + cudaGetErrorString(e);
        return RG_ERR_CUDA;
    }
};

// Groups of the pathwise DP for one direction, as CSR over rows.
static void build_groups(const FlatGraph& f, const std::vector<uint8_t>& nwp, const std::vector<uint32_t>& poff,
                         const std::vector<uint32_t>& pidx, const std::vector<uint32_t>& ebits, bool reverse,
                         std::vector<uint32_t>& grp_off, std::vector<PwGroup>& grp, std::vector<uint32_t>& grp_mask,
                         uint32_t& max_groups) {
    const uint32_t n = f.n, PW = f.PW;
    grp_off.assign(n + 1, 0);
    grp.clear();
    grp_mask.clear();
    max_groups = 1;
    auto bit = [&](const uint32_t* m, uint32_t p) { return (m[p / 32] >> (p % 32)) & 1u; };
    const uint32_t base_row = reverse ? n - 1 : 0;   // the row the DP starts from has no groups
    for (uint32_t i = 0; i < n; i++) {
        grp_off[i] = (uint32_t)grp.size();
        if (i == base_row) continue;
        const uint32_t* nb = &f.node_bits[(size_t)i * PW];
        auto add = [&](uint32_t pred, const uint32_t* emask, bool is_end_row) {
            std::vector<uint32_t> m(PW);
            bool any = false;
            for (uint32_t w = 0; w < PW; w++) {
                m[w] = emask ? (nb[w] & emask[w]) : (nb[w] & f.node_bits[(size_t)pred * PW + w]);
                any |= m[w] != 0;
            }
            if (!any) return;
            PwGroup gpr{};
            gpr.pred = pred;
            uint32_t ap = f.alphas[pred], ai = f.alphas[i];
            uint32_t leader;
            if (ap < f.P && bit(m.data(), ap))
                leader = ap;
            else if (ai < f.P && bit(m.data(), ai))
                leader = ai;
            else {
                leader = 0;
                while (!bit(m.data(), leader)) leader++;
            }
            gpr.leader = leader;
            gpr.lead_is_alpha_of_pred = (!is_end_row && leader == ap) ? 1u : 0u;
            grp.push_back(gpr);
            grp_mask.insert(grp_mask.end(), m.begin(), m.end());
        };
        const bool end_row = reverse ? (i == 0) : (i == n - 1);
        if (!nwp[i] && !end_row) {
            add(reverse ? i + 1 : i - 1, nullptr, false);
        } else {
            for (uint32_t k = poff[i]; k < poff[i + 1]; k++) add(pidx[k], &ebits[(size_t)k * PW], end_row);
        }
        if (!end_row) max_groups = std::max<uint32_t>(max_groups, (uint32_t)grp.size() - grp_off[i]);
    }
    grp_off[n] = (uint32_t)grp.size();
}

// Row records of the score-transport kernel for one direction (see device.h: PwtRow). Decides which rows transport and
// which materialise, numbers the tables, and sizes the table ring so that no table is overwritten while a row that
// refers to it can still be read (as a predecessor, or by the deferred last-column step of the next row).
static void build_pwt_rows(const FlatGraph& f, const std::vector<uint32_t>& grp_off, const std::vector<PwGroup>& grp,
                           const std::vector<uint32_t>& grp_mask, bool reverse, std::vector<PwtRow>& rows, uint32_t& TR) {
    const uint32_t n = f.n, PW = f.PW;
    rows.assign(n, PwtRow{});
    const uint32_t base_row = reverse ? n - 1 : 0;
    for (uint32_t i = 0; i < n; i++) {
        rows[i].g0 = grp_off[i];
        rows[i].lnz = f.lnz[i];
        rows[i].nmh = 255;
        for (uint32_t q = 0; q < f.P; q++)
            if (!((f.node_bits[(size_t)i * PW + q / 32] >> (q % 32)) & 1u)) rows[i].nmh = (uint8_t)q;
    }
    uint32_t next_id = 1, need = 2;
    rows[base_row].tid = 0;
    uint32_t prev = base_row;
    // (table id, path set) the per-origin maxima of modes 8/9 currently describe
    uint32_t mx_tid = 0;
    std::vector<uint32_t> mx_set(&f.node_bits[(size_t)base_row * PW], &f.node_bits[(size_t)base_row * PW] + PW);
    auto same_set = [&](const uint32_t* a, const uint32_t* b) {
        for (uint32_t w = 0; w < PW; w++)
            if (a[w] != b[w]) return false;
        return true;
    };
    for (uint32_t t = 1; t + 1 < n; t++) {
        const uint32_t i = reverse ? n - 1 - t : t;
        const uint32_t g0 = grp_off[i], g1 = grp_off[i + 1];
        const uint32_t* nb = &f.node_bits[(size_t)i * PW];
        PwtRow& r = rows[i];
        const bool transport = (g1 - g0 == 1) && same_set(&grp_mask[(size_t)g0 * PW], nb);
        if (transport) {
            r.kind |= PWT_T;
            r.pred = grp[g0].pred;
            r.leader = (uint8_t)grp[g0].leader;
            r.tid = rows[r.pred].tid;
            if (r.pred != prev) rows[r.pred].kind |= PWT_RING;
            if (r.tid != mx_tid || !same_set(nb, mx_set.data())) {
                r.kind |= PWT_MXREBUILD;
                mx_tid = r.tid;
                mx_set.assign(nb, nb + PW);
            }
        } else {
            r.tid = next_id++;
            for (uint32_t gi = g0; gi < g1; gi++) rows[grp[gi].pred].kind |= PWT_RING;
            mx_tid = r.tid;
            mx_set.assign(nb, nb + PW);
        }
        const uint32_t newest = next_id - 1;
        for (uint32_t gi = g0; gi < g1; gi++) need = std::max(need, newest - rows[grp[gi].pred].tid + 1);
        need = std::max(need, newest - rows[prev].tid + 1);
        prev = i;
    }
    const uint32_t end_row = reverse ? 0 : n - 1;
    if (!reverse)
        for (uint32_t gi = grp_off[end_row]; gi < grp_off[end_row + 1]; gi++) rows[grp[gi].pred].kind |= PWT_FPRED;
    TR = 2;
    while (TR < need) TR <<= 1;
}

static int upload_path_graph(rg_ctx* c) {
    FlatGraph& f = c->fg;
    c->has_path_graph = false;
    c->path_graph_error.clear();
    if (!f.has_paths) {
        c->path_graph_error = "the graph has no paths (P lines): pathwise modes need them";
        return RG_OK;
    }
    for (uint32_t i = 1; i + 1 < f.n; i++)
        if (f.alphas[i] >= f.P) {
            // pathwise_graph.rs:182 + pathwise_alignment_semiglobal.rs:62-70: index alphas[i] = P+1 panics
            c->path_graph_error = "a segment is covered by no path: the reference panics (alphas = P+1)";
            return RG_OK;
        }
    std::vector<uint32_t> grp_off, grp_mask;
    std::vector<PwGroup> grp;
    uint32_t max_groups = 1;
    build_groups(f, f.pw_nwp, f.pw_pred_off, f.pw_pred_idx, f.pw_edge_bits, false, grp_off, grp, grp_mask, max_groups);
    std::vector<uint32_t> rgrp_off, rgrp_mask;
    std::vector<PwGroup> rgrp;
    uint32_t rmax_groups = 1;
    build_groups(f, f.rv_nwp, f.rv_pred_off, f.rv_pred_idx, f.rv_edge_bits, true, rgrp_off, rgrp, rgrp_mask, rmax_groups);
    std::vector<PwtRow> trows, trrows;
    uint32_t TRf = 2, TRr = 2;
    build_pwt_rows(f, grp_off, grp, grp_mask, false, trows, TRf);
    build_pwt_rows(f, rgrp_off, rgrp, rgrp_mask, true, trrows, TRr);
    std::vector<int32_t> nonmem_hi(f.n, -1);
    for (uint32_t i = 0; i < f.n; i++)
        for (uint32_t q = 0; q < f.P; q++)
            if (!((f.node_bits[(size_t)i * f.PW + q / 32] >> (q % 32)) & 1u)) nonmem_hi[i] = (int32_t)q;
    cudaStream_t st = c->stream;
    if (!(c->d_pwt_rows.upload(trows, st) && c->d_pwt_rrows.upload(trrows, st) && c->d_nonmem_hi.upload(nonmem_hi, st)))
        return c->cuda_fail("path graph upload");
    bool ok = c->d_alphas.upload(f.alphas, st) && c->d_node_bits.upload(f.node_bits, st) && c->d_grp_off.upload(grp_off, st) &&
              c->d_grp_mask.upload(grp_mask, st) && c->d_grp.upload(grp, st) && c->d_pw_nwp.upload(f.pw_nwp, st) &&
              c->d_rgrp_off.upload(rgrp_off, st) && c->d_rgrp_mask.upload(rgrp_mask, st) && c->d_rgrp.upload(rgrp, st) &&
              c->d_rv_nwp.upload(f.rv_nwp, st) && c->d_seg.upload(f.row_seg_id, st) && c->d_dfs.upload(f.dfs, st) &&
              c->d_dfe.upload(f.dfe, st);
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) return c->cuda_fail("path graph upload");
    uint32_t ring = 2;
    while (ring <= f.pw_max_lookback) ring <<= 1;
    c->dpg.n = f.n;
    c->dpg.P = f.P;
    c->dpg.PW = f.PW;
    c->dpg.lnz = c->d_lnz.p;
    c->dpg.alphas = c->d_alphas.p;
    c->dpg.node_bits = c->d_node_bits.p;
    c->dpg.grp_off = c->d_grp_off.p;
    c->dpg.grp = c->d_grp.p;
    c->dpg.grp_mask = c->d_grp_mask.p;
    c->dpg.nwp = c->d_pw_nwp.p;
    c->dpg.seg = c->d_seg.p;
    c->dpg.dfs = c->d_dfs.p;
    c->dpg.dfe = c->d_dfe.p;
    c->dpg.ring = ring;
    c->dpg.max_groups = max_groups;
    c->dpg.rows = c->d_pwt_rows.p;
    c->dpg.nonmem_hi = c->d_nonmem_hi.p;
    c->dpg.TR = TRf;
    c->dpg.n_groups = (uint32_t)grp.size();
    c->dpg_rev = c->dpg;
    c->dpg_rev.rows = c->d_pwt_rrows.p;
    c->dpg_rev.TR = TRr;
    c->dpg_rev.n_groups = (uint32_t)rgrp.size();
    c->dpg_rev.grp_off = c->d_rgrp_off.p;
    c->dpg_rev.grp = c->d_rgrp.p;
    c->dpg_rev.grp_mask = c->d_rgrp_mask.p;
    c->dpg_rev.nwp = c->d_rv_nwp.p;
    uint32_t rring = 2;
    while (rring <= f.rv_max_lookback) rring <<= 1;
    c->dpg_rev.ring = rring;
    c->dpg_rev.max_groups = rmax_groups;
    c->has_path_graph = true;
    return RG_OK;
}

static int upload_graph(rg_ctx* c) {
    FlatGraph& f = c->fg;
    cudaStream_t st = c->stream;
    c->has_graph = false;   // stays false if the upload fails half-way
    std::vector<uint8_t> rowflags(f.n, 0);
    // RF_SINGLE_PREV: segment start whose only predecessor is the row right above it (and not the '$' row): the
    // kernels treat it like a row inside a segment, so its predecessor need not be kept in the ring either.
    for (uint32_t i = 0; i < f.n; i++) rowflags[i] = (f.nwp[i] ? RF_NWP : 0);
    for (uint32_t i = 2; i + 1 < f.n; i++) {
        if (!f.nwp[i]) continue;
        uint32_t b = f.pred_off[i], e = f.pred_off[i + 1];
        if (e - b == 1 && f.pred_idx[b] == i - 1)
            rowflags[i] |= RF_SINGLE_PREV;
        else
            for (uint32_t k = b; k < e; k++) rowflags[f.pred_idx[k]] |= RF_IS_PRED;
    }
    if (f.n > 2 && f.nwp[1])
        for (uint32_t k = f.pred_off[1]; k < f.pred_off[2]; k++) rowflags[f.pred_idx[k]] |= RF_IS_PRED;
    for (uint32_t k = f.pred_off[f.n - 1]; k < f.pred_off[f.n]; k++) rowflags[f.pred_idx[k]] |= RF_F_PRED;
    std::vector<uint32_t> nwp_ord(f.n, 0);
    c->n_gather = 0;
    for (uint32_t i = 0; i < f.n; i++)
        if ((rowflags[i] & RF_NWP) && !(rowflags[i] & RF_SINGLE_PREV)) nwp_ord[i] = c->n_gather++;
    std::vector<RowInfo> rowinfo(f.n);
    for (uint32_t i = 0; i < f.n; i++) {
        uint32_t np = f.nwp[i] ? f.pred_off[i + 1] - f.pred_off[i] : 0;
        RowInfo ri;
        ri.r_value = f.r_values[i];
        ri.min_pred = f.min_pred[i];
        ri.pred_off = f.pred_off[i];
        ri.lnz = f.lnz[i];
        ri.flags = rowflags[i];
        ri.min_pred_slot = f.min_pred_slot[i];
        ri.npred = (uint8_t)std::min<uint32_t>(np, 255);
        rowinfo[i] = ri;
    }
    bool ok = c->d_rowinfo.upload(rowinfo, st) && c->d_lnz.upload(f.lnz, st) && c->d_rowflags.upload(rowflags, st) &&
              c->d_pred_off.upload(f.pred_off, st) && c->d_pred_idx.upload(f.pred_idx, st) &&
              c->d_min_pred.upload(f.min_pred, st) && c->d_min_pred_slot.upload(f.min_pred_slot, st) &&
              c->d_prev_slot.upload(f.prev_slot, st) && c->d_r_values.upload(f.r_values, st) &&
              c->d_nwp_ord.upload(nwp_ord, st);
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) return c->cuda_fail("graph upload");
    uint32_t ring = 2;
    while (ring <= f.max_lookback) ring <<= 1;
    c->dg.n = f.n;
    c->dg.lnz = c->d_lnz.p;
    c->dg.rowflags = c->d_rowflags.p;
    c->dg.pred_off = c->d_pred_off.p;
    c->dg.pred_idx = c->d_pred_idx.p;
    c->dg.min_pred = c->d_min_pred.p;
    c->dg.min_pred_slot = c->d_min_pred_slot.p;
    c->dg.prev_slot = c->d_prev_slot.p;
    c->dg.r_values = c->d_r_values.p;
    c->dg.rowinfo = c->d_rowinfo.p;
    c->dg.ring = ring;
    c->dg.nwp_ord = c->d_nwp_ord.p;
    c->dg.n_gather = c->n_gather;
    c->has_graph = true;
    c->results_ready = false;
    return RG_OK;
}

extern "C" {

const char* rg_strerror(int s) {
    switch (s) {
        case RG_OK: return "ok";
        case RG_ERR_INVALID: return "invalid argument or call order";
        case RG_ERR_CUDA: return "CUDA error";
        case RG_ERR_NO_DEVICE: return "no CUDA device (recgraph_b200 has no CPU fallback)";
        case RG_ERR_IO: return "I/O or format error";
        case RG_ERR_BAD_CHAR: return "character outside A,C,G,T,N";
        case RG_ERR_UNSUPPORTED: return "input outside the supported domain";
        case RG_ERR_REF_PANIC: return "the reference implementation panics on this input";
        case RG_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

const char* rg_last_error(const rg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }

int rg_init(int device, rg_ctx** out) {
    if (!out) return RG_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return RG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return RG_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return RG_ERR_CUDA;
    }
    rg_ctx* c = new (std::nothrow) rg_ctx();
    if (!c) return RG_ERR_NOMEM;
    c->device = device;
    cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return RG_ERR_CUDA;
    }
    rg_default_scoring(&c->scoring);
    rg_set_scoring(c, &c->scoring);
    if (!c->h_counters.ensure(4)) {
        delete c;
        return RG_ERR_NOMEM;
    }
    *out = c;
    return RG_OK;
}

void rg_destroy(rg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rg_load_gfa_text(rg_ctx* c, const char* text, size_t len) {
    if (!c || !text) return RG_ERR_INVALID;
    cudaSetDevice(c->device);
    GfaGraph g;
    std::string err;
    if (!parse_gfa(text, len, g, err)) return c->fail(err.find("outside the supported") != std::string::npos ? RG_ERR_UNSUPPORTED : RG_ERR_IO, err);
    FlatGraph tmp;   // a failed load must leave the ctx's graph untouched
    int rc = flatten_graph(g, tmp, err);
    if (rc != RG_OK) return c->fail(rc, err);
    c->fg = std::move(tmp);
    rc = upload_graph(c);
    if (rc != RG_OK) return rc;
    return upload_path_graph(c);
}

int rg_load_gfa_file(rg_ctx* c, const char* path) {
    if (!c || !path) return RG_ERR_INVALID;
    std::ifstream f(path, std::ios::binary);
    if (!f) return c->fail(RG_ERR_IO, std::string("cannot open ") + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    return rg_load_gfa_text(c, s.data(), s.size());
}

int rg_set_lnz_graph(rg_ctx* c, uint32_t n, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                     const uint32_t* pred_idx, const uint64_t* seg_id) {
    if (!c || !lnz_codes || !nwp || !pred_off || !pred_idx) return RG_ERR_INVALID;
    cudaSetDevice(c->device);
    std::string err;
    FlatGraph tmp;
    int rc = flat_from_lnz(n, lnz_codes, nwp, pred_off, pred_idx, seg_id, tmp, err);
    if (rc != RG_OK) return c->fail(rc, err);
    c->fg = std::move(tmp);
    // a path graph left by an earlier rg_load_gfa_* does not describe this graph
    c->has_path_graph = false;
    c->path_graph_error = "graph set without paths (rg_set_lnz_graph): pathwise modes need rg_load_gfa_* or rg_set_path_graph";
    return upload_graph(c);
}

int rg_set_path_graph(rg_ctx* c, uint32_t n, uint32_t n_paths, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                      const uint32_t* pred_idx, const uint32_t* edge_path_bits, const uint32_t* node_path_bits,
                      const uint32_t* alphas, const uint64_t* seg_id) {
    if (!c || !lnz_codes || !nwp || !pred_off || !pred_idx || !edge_path_bits || !node_path_bits || !alphas) return RG_ERR_INVALID;
    cudaSetDevice(c->device);
    std::string err;
    FlatGraph tmp;
    int rc = flat_from_path_graph(n, n_paths, lnz_codes, nwp, pred_off, pred_idx, edge_path_bits, node_path_bits, alphas, seg_id, tmp, err);
    if (rc != RG_OK) return c->fail(rc, err);
    c->fg = std::move(tmp);
    rc = upload_graph(c);
    if (rc != RG_OK) return rc;
    return upload_path_graph(c);
}

int rg_graph_info(const rg_ctx* c, uint32_t* n, uint32_t* n_segments, uint32_t* n_paths) {
    if (!c || !c->has_graph) return RG_ERR_INVALID;
    if (n) *n = c->fg.n;
    if (n_segments) *n_segments = c->fg.n_segments;
    if (n_paths) *n_paths = c->fg.P;
    return RG_OK;
}

int rg_make_score_matrix(int kind, int32_t match, int32_t mismatch, rg_scoring* s) {
    return make_score_matrix(kind, match, mismatch, s);
}

void rg_default_scoring(rg_scoring* s) {
    if (!s) return;
    make_score_matrix(0, 2, -4, s);
    s->gap_open = -4;
    s->gap_ext = -2;
    s->base_rec_cost = 4;
    s->multi_rec_cost = 0.1f;
    s->rec_band_width = 1.0f;
    s->extra_b = 1.0f;
    s->extra_f = 0.01f;
    s->fixed_bta = -1;
}

int rg_set_scoring(rg_ctx* c, const rg_scoring* s) {
    if (!c || !s) return RG_ERR_INVALID;
    c->scoring = *s;
    memset(&c->ds, 0, sizeof c->ds);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) c->ds.sc[i][j] = s->score[i][j];
    c->ds.o = s->gap_open;
    c->ds.e = s->gap_ext;
    c->ds.b = s->extra_b;
    c->ds.f = s->extra_f;
    c->ds.fixed_bta = s->fixed_bta;
    c->ds.R = s->base_rec_cost;
    c->ds.r = s->multi_rec_cost;
    c->ds.rbw = s->rec_band_width;
    c->has_scoring = true;
    return RG_OK;
}

int rg_upload_reads(rg_ctx* c, int32_t n_reads, const uint8_t* read_codes, const uint64_t* read_off) {
    if (!c || n_reads < 0 || !read_off || (n_reads > 0 && !read_codes)) return RG_ERR_INVALID;
    cudaSetDevice(c->device);
    c->results_ready = false;
    c->n_reads = n_reads;
    c->h_off.assign(read_off, read_off + n_reads + 1);
    c->total_len = read_off[n_reads] - read_off[0];
    uint32_t mx = 0;
    for (int32_t i = 0; i < n_reads; i++) {
        uint64_t l = read_off[i + 1] - read_off[i];
        if (l == 0) return c->fail(RG_ERR_INVALID, "empty read (the reference's FASTA reader never yields one)");
        if (l > 0x0fffffff) return c->fail(RG_ERR_UNSUPPORTED, "read too long");
        mx = std::max<uint32_t>(mx, (uint32_t)l);
    }
    c->max_len = mx;
    for (uint64_t k = read_off[0]; k < read_off[n_reads]; k++)
        if (read_codes[k] > RG_N) return c->fail(RG_ERR_BAD_CHAR, "read code outside 0..4 (A,C,G,T,N)");
    // longest reads first: the persistent kernel hands reads out in this order (load balance)
    std::vector<int32_t> order(n_reads);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return read_off[a + 1] - read_off[a] > read_off[b + 1] - read_off[b];
    });
    std::vector<uint64_t> off0(n_reads + 1);
    for (int32_t i = 0; i <= n_reads; i++) off0[i] = read_off[i] - read_off[0];
    if (!c->d_reads.ensure(c->total_len + 16) || !c->d_read_off.ensure(n_reads + 1) || !c->d_order.ensure(n_reads + 1))
        return c->fail(RG_ERR_NOMEM, "device allocation for reads failed");
    cudaStream_t st = c->stream;
    if (c->total_len)
        cudaMemcpyAsync(c->d_reads.p, read_codes + read_off[0], c->total_len, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(c->d_read_off.p, off0.data(), (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (n_reads) cudaMemcpyAsync(c->d_order.p, order.data(), n_reads * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return c->cuda_fail("read upload");
    return RG_OK;
}

// Work-spaces are sized from the free memory of the device: a family that starts after another one first returns that
// family's buffers (they are re-created on demand).
static void release_pathwise_ws(rg_ctx* c) {
    c->d_tr_tables.release(), c->d_tr_ring_lead.release(), c->d_tr_lastcol.release(), c->d_tr_colmax.release(), c->d_tr_stage.release();
    c->d_tr_ring_org.release(), c->d_tr_ring_meta.release(), c->d_tr_mv_f.release(), c->d_tr_mv_r.release(), c->d_tr_own.release();
    c->d_tr_own_pred.release(), c->d_tr_cb_f.release(), c->d_tr_cb_r.release(), c->d_gapT.release();
    c->d_pwS.release(), c->d_pwLead.release(), c->d_pwTrace.release(), c->d_rS.release(), c->d_rLead.release(), c->d_rTrace.release();
    c->d_fm.release(), c->d_rwbest.release(), c->d_lastcol.release();
    c->tr_sig = 0;
    c->tr_slots = 0;
}
static void release_poa_ws(rg_ctx* c) {
    c->d_rowmeta.release(), c->d_ring_m.release(), c->d_ring_y.release(), c->d_trace.release();
}

static int align_poa(rg_ctx* c, int mode) {
    release_pathwise_ws(c);
    const FlatGraph& f = c->fg;
    const uint32_t n = f.n;
    const uint32_t Lmax = c->max_len + 1;
    const bool lin = mode == RG_MODE_GLOBAL || mode == RG_MODE_LOCAL || mode == RG_MODE_GAP_LOCAL || mode == RG_MODE_LOCAL_SCALAR;
    const int trace_bytes = lin ? poa_lin_trace_bytes(f.max_indeg) : (f.max_indeg <= 4 ? 1 : 2);
    if (f.max_indeg > (lin ? 31u : 64u)) return c->fail(RG_ERR_UNSUPPORTED, "in-degree above the trace-code domain (31 for modes 0/1/3, 64 for mode 2)");
    if (mode == RG_MODE_GLOBAL || mode == RG_MODE_LOCAL)
        for (int k = 1; k < 5; k++)
            if (c->scoring.score[k][5] != c->scoring.score[0][5])
                return c->fail(RG_ERR_UNSUPPORTED, "modes 0/1 need one gap score for all characters (true for every matrix the reference builds)");
    if (mode == RG_MODE_GLOBAL_SCALAR || mode == RG_MODE_LOCAL_SCALAR)
        for (int k = 0; k < 5; k++)
            if (c->scoring.score[k][5] != c->scoring.score[0][5] || c->scoring.score[5][k] != c->scoring.score[0][5])
                return c->fail(RG_ERR_UNSUPPORTED, "scalar mode 0 needs one gap score for all characters (true for every matrix the reference builds)");
    uint32_t wstride = (Lmax + 31) & ~31u;
    int ws_cols = 64, blocks_per_sm = 1;
    // register-blocked kernel (lane-owned column blocks) whenever the longest read fits 32*C columns
    const int blkC = lin ? gap_blk_cols(Lmax) : ((mode == RG_MODE_GAP_GLOBAL && !c->force_striped) ? gap_blk_cols(Lmax) : 0);
    if (lin && !blkC) return c->fail(RG_ERR_UNSUPPORTED, "modes 0/1/3: reads longer than 1023 bases are not supported on the device yet");
    int lc;
    if (blkC) {
        wstride = 32u * blkC;
        if (lin) {
            lc = 0;
            blocks_per_sm = 2;
        } else {
            lc = gap_blk_blocks_per_sm(blkC, trace_bytes, &blocks_per_sm);
        }
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    } else {
        lc = poa_launch_config(mode, trace_bytes, Lmax, &ws_cols, &blocks_per_sm);
    }
    if (lc == -2) return c->fail(RG_ERR_UNSUPPORTED, "alignment mode not implemented on the device yet");
    if (lc != 0) return c->cuda_fail("kernel configuration");
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // memory already held by this ctx's work-space is reusable
    free_b += (c->d_rowmeta.cap * sizeof(RowMeta)) + (c->d_ring_m.cap + c->d_ring_y.cap) * 4 + c->d_trace.cap +
              (c->d_slot_runs.cap + c->d_out_runs.cap) * sizeof(rg_run);
    // blocked mode-2 kernel: 4 bit planes per cell (32 * C/8 words per row) + predecessor-slot planes of the gathering rows
    const bool blk2 = blkC && !lin;
    const int SBk = trace_bytes == 1 ? 2 : 6;
    const uint64_t plane_bytes = (((uint64_t)n * 128u * std::max(1, blkC / 8)) + 255) & ~255ull;
    const uint64_t side_bytes = (uint64_t)c->n_gather * 2u * SBk * 128u;
    const int tb_alloc = blk2 ? 1 : trace_bytes;
    const uint64_t full = blk2 ? plane_bytes + side_bytes
                          : blkC ? (uint64_t)n * wstride  // fixed row stride, absolute columns
                              : (uint64_t)n * (Lmax + 15);  // the band can open to the whole row; rows start at multiples of 16 cells
    const uint32_t run_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4096, (uint64_t)(n / 4 + 2 * Lmax)), 1u << 22);
    const size_t per_slot_fixed = (size_t)n * sizeof(RowMeta) + (size_t)c->dg.ring * wstride * 8 + (size_t)run_cap * sizeof(rg_run);
    const size_t budget_all = (size_t)(free_b * 0.85);
    size_t out_runs_cap = std::min<uint64_t>((uint64_t)c->n_reads * std::min<uint32_t>(run_cap, 16384), (budget_all / 8) / sizeof(rg_run));
    out_runs_cap = std::max<size_t>(out_runs_cap, 1024);
    const size_t budget = budget_all - out_runs_cap * sizeof(rg_run);
    const uint32_t reads8 = (uint32_t)((c->n_reads + 7) / 8 * 8);
    const uint32_t wpb = blk2 ? (uint32_t)gap_blk_warps_per_block() : 8u;   // warps (= slots) per CTA of the kernel that will run
    uint32_t slots = std::min<uint32_t>((uint32_t)c->sms * blocks_per_sm * wpb, reads8);
    for (int attempt = 0; attempt < 8; attempt++) {
        // Reads in flight are limited by trace memory: prefer the full n x L trace per slot (never overflows),
        // shrink the number of slots down to one block per SM before shrinking the trace.
        uint64_t trace_cap = full;
        size_t per_slot = per_slot_fixed + (size_t)trace_cap * tb_alloc;
        if ((size_t)slots * per_slot > budget) {
            uint32_t min_slots = std::min<uint32_t>(slots, std::max<uint32_t>(8, (uint32_t)c->sms * 8 >> attempt));
            uint32_t fit = (uint32_t)std::min<size_t>(budget / per_slot, 1u << 20) / 8 * 8;
            if (blkC) {
                // the blocked kernels address their trace by row: a slot needs the whole n x 32C trace, so large graphs
                // simply run with fewer reads in flight
                if (fit < 8) return c->fail(RG_ERR_NOMEM, "not enough device memory for the blocked kernel's trace");
                slots = std::min(slots, fit);
            } else if (fit >= min_slots) {
                slots = fit;
            } else {
                slots = min_slots;
                size_t each = budget / slots;
                if (each <= per_slot_fixed + 4096) return c->fail(RG_ERR_NOMEM, "not enough device memory for the alignment work-space");
                trace_cap = std::min<uint64_t>(full, (each - per_slot_fixed) / tb_alloc);
            }
        }
        trace_cap = (trace_cap + 255) & ~255ull;  // round UP: `full` must always fit
        if (blkC && trace_cap < full) return c->fail(RG_ERR_NOMEM, "not enough device memory for the blocked kernel's trace");
        bool ok = c->d_rowmeta.ensure((size_t)slots * n) && c->d_ring_m.ensure((size_t)slots * c->dg.ring * wstride) &&
                  c->d_ring_y.ensure((size_t)slots * c->dg.ring * wstride) &&
                  c->d_trace.ensure((size_t)slots * trace_cap * tb_alloc) &&
                  c->d_slot_runs.ensure((size_t)slots * run_cap) && c->d_out_runs.ensure(out_runs_cap) &&
                  c->d_results.ensure(c->n_reads + 1) && c->d_counters.ensure(4);
        if (!ok) return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
        PoaWorkspace ws{};
        ws.rowmeta = c->d_rowmeta.p;
        ws.ring_m = c->d_ring_m.p;
        ws.ring_y = c->d_ring_y.p;
        ws.trace = c->d_trace.p;
        ws.runs = c->d_slot_runs.p;
        ws.trace_cap = trace_cap;
        ws.run_cap = run_cap;
        ws.wstride = wstride;
        ws.slots = slots;
        ws.use16 = c->no_s16 ? 0u : 1u;
        ws.side_off = plane_bytes;
        PoaBatch b{};
        b.reads = c->d_reads.p;
        b.read_off = c->d_read_off.p;
        b.n_reads = c->n_reads;
        b.order = c->d_order.p;
        b.results = c->d_results.p;
        b.out_runs = c->d_out_runs.p;
        b.out_run_cap = out_runs_cap;
        b.counters = c->d_counters.p;
        cudaMemsetAsync(c->d_counters.p, 0, 4 * sizeof(unsigned long long), c->stream);
        cudaEventRecord(c->ev0, c->stream);
        if (blkC && trace_cap < full) return c->fail(RG_ERR_NOMEM, "not enough device memory for the blocked kernel's trace");
        int rc = lin ? launch_poa_lin(mode, blkC, c->dg, c->ds, ws, b, trace_bytes, (int)(slots / 8), c->stream)
                 : blkC ? launch_gap_global_blk(blkC, c->dg, c->ds, ws, b, trace_bytes,
                                                (int)std::max<uint32_t>((slots + wpb - 1) / wpb, std::min<uint32_t>((uint32_t)c->sms * blocks_per_sm, slots)), c->stream)
                      : launch_poa(mode, c->dg, c->ds, ws, b, trace_bytes, (int)(slots / 8), ws_cols, c->stream);
        cudaEventRecord(c->ev1, c->stream);
        if (rc != 0 || cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("alignment kernel");
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->kernel_ms += ms;
        c->launches += 1;
        c->slots_used = slots;
        // overflow check needs the statuses: cheap D2H of the records
        if (!c->h_results.ensure(c->n_reads + 1)) return c->fail(RG_ERR_NOMEM, "pinned allocation failed");
        cudaMemcpyAsync(c->h_results.p, c->d_results.p, (size_t)c->n_reads * sizeof(rg_read_result), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("result copy");
        bool overflow = false;
        for (int32_t i = 0; i < c->n_reads; i++)
            if (c->h_results.p[i].status & RG_READ_TRACE_OVERFLOW) overflow = true;
        c->n_runs_total = std::min<uint64_t>(c->h_counters.p[1], out_runs_cap);
        if (!overflow) return RG_OK;
        // rare: the whole batch is re-run with fewer reads in flight / larger run buffers
        if (c->h_counters.p[1] > out_runs_cap) out_runs_cap = std::min<size_t>(c->h_counters.p[1] * 2, (budget_all / 2) / sizeof(rg_run));
        if (slots <= 8 && trace_cap >= full) return c->fail(RG_ERR_NOMEM, "trace buffers overflow at maximum size");
        slots = std::max<uint32_t>(8, slots / 2 / 8 * 8);
    }
    return c->fail(RG_ERR_NOMEM, "trace buffers still overflow after retries");
}

static int align_pathwise_v1(rg_ctx* c, int mode);

// Modes 4 / 5 / 8 / 9 through the score-transport kernel (pathwise_tr.cu).
static int align_pathwise(rg_ctx* c, int mode) {
    if (!c->has_path_graph)
        return c->fail(c->path_graph_error.find("panics") != std::string::npos ? RG_ERR_REF_PANIC : RG_ERR_INVALID,
                       c->path_graph_error.empty() ? "no path graph" : c->path_graph_error);
    const FlatGraph& f = c->fg;
    for (int k = 1; k < 5; k++)
        if (c->scoring.score[k][5] != c->scoring.score[0][5] || c->scoring.score[5][k] != c->scoring.score[0][5])
            return c->fail(RG_ERR_UNSUPPORTED, "pathwise modes need one gap score for all characters (true for every matrix the reference builds)");
    const uint32_t n = f.n, PW = f.PW;
    const bool rec = mode == RG_MODE_REC_GLOBAL || mode == RG_MODE_REC_SEMIGLOBAL;
    if (f.P > 128) return c->fail(RG_ERR_UNSUPPORTED, "more than 128 paths are not supported on the device yet");
    if (n >= (1u << 21)) return c->fail(RG_ERR_UNSUPPORTED, "graph too large for the pathwise kernels");
    release_poa_ws(c);
    if (c->pw_v1) return align_pathwise_v1(c, mode);
    // ---- reads are grouped by the column block they need (256 x {4, 8, 16, 32} columns, 384 x 32 for the longest): the
    // processing order is longest-first, so every class is a contiguous range of it; one launch per class, each with a
    // work-space sized for its own read length
    struct Cls {
        int32_t lo, hi;     // range of the processing order
        uint32_t LP, cpt, NT;
    };
    std::vector<Cls> classes;
    {
        std::vector<int32_t> order(c->n_reads);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            return c->h_off[a + 1] - c->h_off[a] > c->h_off[b + 1] - c->h_off[b];
        });   // the same order rg_upload_reads stored on the device
        for (int32_t k = 0; k < c->n_reads; k++) {
            const uint32_t L = (uint32_t)(c->h_off[order[k] + 1] - c->h_off[order[k]]) + 1;
            uint32_t cpt = (uint32_t)pathwise_tr_cpt(L), nt = 256;
            if (!cpt) {
                cpt = (uint32_t)pathwise_tr_cpt_wide(L);
                nt = 384;
            }
            if (!cpt) return align_pathwise_v1(c, mode);   // beyond 12 287 bases: the per-path kernel (bounded by shared memory)
            if (classes.empty() || classes.back().cpt != cpt || classes.back().NT != nt) classes.push_back({k, k, nt * cpt, cpt, nt});
            classes.back().hi = k + 1;
        }
    }
    if (!c->d_results.ensure(c->n_reads + 1) || !c->d_counters.ensure(4)) return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
    cudaMemsetAsync(c->d_counters.p, 0, 4 * sizeof(unsigned long long), c->stream);
    // one run buffer for the whole batch, sized before the per-class work-spaces
    size_t out_runs_cap;
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += c->d_out_runs.cap * sizeof(rg_run);
        const uint32_t rc_max = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4096, (uint64_t)n + 2 * classes.front().LP), 1u << 22);
        out_runs_cap = std::min<uint64_t>((uint64_t)c->n_reads * std::min<uint32_t>(rc_max, 16384), ((size_t)(free_b * 0.85) / 8) / sizeof(rg_run));
        out_runs_cap = std::max<size_t>(out_runs_cap, 1024);
        if (!c->d_out_runs.ensure(out_runs_cap)) return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
    }
    for (const Cls& cl : classes) {
        const bool wide = cl.NT == 384;
        PwtWorkspace ws{};
        ws.CPT = cl.cpt;
        ws.NT = cl.NT;
        ws.LP = cl.LP;
        ws.LT = ws.LP + 32;
        ws.Pp = PW * 32;
        ws.TRmax = std::max(c->dpg.TR, rec ? c->dpg_rev.TR : 2u);
        ws.ringmax = std::max(c->dpg.ring, rec ? c->dpg_rev.ring : 2u);
        ws.run_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4096, (uint64_t)n + 2 * ws.LP), 1u << 22);
        ws.diag = getenv("RG_PW_DIAG") ? 1u : 0u;
        ws.stage = wide ? (int32_t*)1 : nullptr;   // non-null: sizes the shared memory for the wide layout (pointer set below)
        int bps = 1;
        const bool mx = mode != RG_MODE_PATHWISE_GLOBAL;
        int lc = wide ? pathwise_tr_blocks_per_sm_wide(c->dpg, c->dpg_rev, c->ds, ws, mx, &bps)
                      : pathwise_tr_blocks_per_sm(c->dpg, c->dpg_rev, c->ds, ws, mx, &bps);
        if (lc == -3) return c->fail(RG_ERR_UNSUPPORTED, "read too long for the pathwise kernel's shared memory");
        if (lc != 0) return c->cuda_fail("kernel configuration");
        if (bps < 1) bps = 1;
        if (const char* e = getenv("RG_PW_BPS")) bps = std::max(1, std::min(bps, atoi(e)));  // testing: CTAs per SM
        // The work-space of a (mode, column class, batch size) is kept between calls; any other shape starts from a clean
        // slate, so that the sizes below are taken from what is really free (a buffer that stays larger than needed would be
        // counted as reusable without being so).
        const uint64_t sig = ((uint64_t)mode << 56) ^ ((uint64_t)cl.LP << 32) ^ ((uint64_t)cl.NT << 24) ^ (uint64_t)(cl.hi - cl.lo) ^
                             ((uint64_t)n << 8) ^ ((uint64_t)c->dpg.n_groups << 40);
        const bool same_shape = sig == c->tr_sig && c->tr_slots > 0;
        if (!same_shape) {
            release_pathwise_ws(c);
            c->d_slot_runs.release();
        }
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t sz_tables = (size_t)ws.TRmax * ws.Pp * ws.LT, sz_ring = (size_t)ws.ringmax * ws.LP;
        const size_t sz_mvf = (size_t)c->dpg.n_groups * (ws.LP / 4), sz_mvr = rec ? (size_t)c->dpg_rev.n_groups * (ws.LP / 4) : 0;
        const size_t sz_own = (size_t)n * (ws.LP / 4), sz_cb = rec ? (size_t)n * ws.LP : 0, sz_last = rec ? (size_t)n * ws.Pp : 0;
        const size_t sz_stage = wide ? (size_t)2 * ws.LP : 0;
        const size_t per_slot = sz_tables * 4 + sz_ring * 6 + (size_t)ws.ringmax * 16 + sz_mvf + sz_mvr + sz_own + (size_t)n * 4 +
                                sz_cb * 16 + sz_last * 4 + (size_t)2 * ws.LP * 4 + sz_stage * 4 + (size_t)ws.run_cap * sizeof(rg_run);
        const size_t budget = (size_t)(free_b * 0.9);
        uint32_t slots = std::min<uint32_t>((uint32_t)c->sms * bps, (uint32_t)(cl.hi - cl.lo));
        slots = same_shape ? c->tr_slots : (uint32_t)std::min<size_t>(slots, budget / per_slot);
        if (slots < 1) return c->fail(RG_ERR_NOMEM, "not enough device memory for one pathwise read in flight");
        ws.slots = slots;
        c->tr_sig = sig;
        c->tr_slots = slots;
        bool ok = c->d_tr_tables.ensure(slots * sz_tables) && c->d_tr_ring_lead.ensure(slots * sz_ring) &&
                  c->d_tr_ring_org.ensure(slots * sz_ring) && c->d_tr_ring_meta.ensure((size_t)slots * ws.ringmax) &&
                  c->d_tr_mv_f.ensure(slots * sz_mvf) && c->d_tr_mv_r.ensure(slots * sz_mvr) && c->d_tr_own.ensure(slots * sz_own) &&
                  c->d_tr_own_pred.ensure((size_t)slots * n) && c->d_tr_cb_f.ensure(slots * sz_cb) && c->d_tr_cb_r.ensure(slots * sz_cb) &&
                  c->d_tr_lastcol.ensure(slots * sz_last) && c->d_tr_colmax.ensure((size_t)slots * 2 * ws.LP) &&
                  c->d_tr_stage.ensure(slots * sz_stage) && c->d_slot_runs.ensure((size_t)slots * ws.run_cap);
        if (!ok) return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
        ws.tables = c->d_tr_tables.p;
        ws.ring_lead = c->d_tr_ring_lead.p;
        ws.ring_org = c->d_tr_ring_org.p;
        ws.ring_meta = c->d_tr_ring_meta.p;
        ws.mv_f = c->d_tr_mv_f.p;
        ws.mv_r = c->d_tr_mv_r.p;
        ws.own = c->d_tr_own.p;
        ws.own_pred = c->d_tr_own_pred.p;
        ws.cb_f = c->d_tr_cb_f.p;
        ws.cb_r = c->d_tr_cb_r.p;
        ws.lastcol = c->d_tr_lastcol.p;
        ws.colmax = c->d_tr_colmax.p;
        ws.stage = wide ? c->d_tr_stage.p : nullptr;
        ws.runs = c->d_slot_runs.p;
        PoaBatch b{};
        b.reads = c->d_reads.p;
        b.read_off = c->d_read_off.p;
        b.n_reads = cl.hi - cl.lo;
        b.order = c->d_order.p + cl.lo;
        b.results = c->d_results.p;
        b.out_runs = c->d_out_runs.p;
        b.out_run_cap = out_runs_cap;
        b.counters = c->d_counters.p;
        cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long), c->stream);   // the ticket restarts, the run count goes on
        cudaEventRecord(c->ev0, c->stream);
        int rc = wide ? launch_pathwise_tr_wide(mode, c->dpg, c->dpg_rev, c->ds, ws, b, (int)slots, c->stream)
                      : launch_pathwise_tr(mode, c->dpg, c->dpg_rev, c->ds, ws, b, (int)slots, c->stream);
        cudaEventRecord(c->ev1, c->stream);
        if (rc != 0 || cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("pathwise kernel");
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->kernel_ms += ms;
        c->launches += 1;
        c->slots_used = slots;
    }
    cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("result copy");
    c->n_runs_total = std::min<uint64_t>(c->h_counters.p[1], out_runs_cap);
    return RG_OK;
}

// The per-path kernel of round 1 (pathwise.cu): every path of every row computed; kept for A/B tests (RG_PW_V1) and for
// reads longer than the transport kernel's column blocks.
static int align_pathwise_v1(rg_ctx* c, int mode) {
    const FlatGraph& f = c->fg;
    const uint32_t n = f.n, PW = f.PW;
    const bool rec = mode == RG_MODE_REC_GLOBAL || mode == RG_MODE_REC_SEMIGLOBAL;
    PwWorkspace ws{};
    PwRecWorkspace rw{};
    ws.Lp = (c->max_len + 1 + 31) & ~31u;
    ws.Pp = PW * 32;
    ws.run_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4096, (uint64_t)n + 2 * ws.Lp), 1u << 22);
    int bps = 1;
    int lc = pathwise_blocks_per_sm(c->dpg, c->dpg_rev, ws, rec, &bps);
    if (lc == -3) return c->fail(RG_ERR_UNSUPPORTED, "read too long for the pathwise kernel's shared-memory move table");
    if (lc != 0) return c->cuda_fail("kernel configuration");
    if (bps < 1) bps = 1;
    if (const char* e = getenv("RG_PW_BPS")) bps = std::max(1, std::min(bps, atoi(e)));  // testing: CTAs per SM
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += (c->d_pwS.cap + c->d_pwLead.cap + c->d_pwTrace.cap + c->d_rS.cap + c->d_rLead.cap + c->d_rTrace.cap + c->d_lastcol.cap) * 4 +
              (c->d_fm.cap + c->d_rwbest.cap) * 8 + (c->d_slot_runs.cap + c->d_out_runs.cap) * sizeof(rg_run);
    size_t per_slot = (size_t)c->dpg.ring * ws.Lp * ws.Pp * 4 + (size_t)c->dpg.ring * ws.Lp * 4 +
                      (size_t)n * ws.Lp * PW * 8 + (size_t)ws.run_cap * sizeof(rg_run);
    if (rec)
        per_slot += (size_t)c->dpg_rev.ring * ws.Lp * ws.Pp * 4 + (size_t)c->dpg_rev.ring * ws.Lp * 4 + (size_t)n * ws.Lp * PW * 8 +
                    (size_t)n * ws.Lp * 16 + (size_t)n * ws.Pp * 4;
    const size_t budget_all = (size_t)(free_b * 0.85);
    size_t out_runs_cap = std::min<uint64_t>((uint64_t)c->n_reads * std::min<uint32_t>(ws.run_cap, 16384), (budget_all / 8) / sizeof(rg_run));
    out_runs_cap = std::max<size_t>(out_runs_cap, 1024);
    const size_t budget = budget_all - out_runs_cap * sizeof(rg_run);
    uint32_t slots = std::min<uint32_t>((uint32_t)c->sms * bps, (uint32_t)c->n_reads);
    slots = (uint32_t)std::min<size_t>(slots, budget / per_slot);
    if (slots < 1) return c->fail(RG_ERR_NOMEM, "not enough device memory for one pathwise read in flight");
    ws.slots = slots;
    bool ok = c->d_pwS.ensure((size_t)slots * c->dpg.ring * ws.Lp * ws.Pp) && c->d_pwLead.ensure((size_t)slots * c->dpg.ring * ws.Lp) &&
              c->d_pwTrace.ensure((size_t)slots * n * ws.Lp * PW * 2) && c->d_slot_runs.ensure((size_t)slots * ws.run_cap) &&
              c->d_out_runs.ensure(out_runs_cap) && c->d_results.ensure(c->n_reads + 1) && c->d_counters.ensure(4);
    if (ok && rec)
        ok = c->d_rS.ensure((size_t)slots * c->dpg_rev.ring * ws.Lp * ws.Pp) && c->d_rLead.ensure((size_t)slots * c->dpg_rev.ring * ws.Lp) &&
             c->d_rTrace.ensure((size_t)slots * n * ws.Lp * PW * 2) && c->d_fm.ensure((size_t)slots * n * ws.Lp) &&
             c->d_rwbest.ensure((size_t)slots * n * ws.Lp) && c->d_lastcol.ensure((size_t)slots * n * ws.Pp);
    if (!ok) return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
    rw.S = c->d_rS.p;
    rw.lead = c->d_rLead.p;
    rw.trace = c->d_rTrace.p;
    rw.fm = c->d_fm.p;
    rw.rw = c->d_rwbest.p;
    rw.lastcol = c->d_lastcol.p;
    ws.S = c->d_pwS.p;
    ws.lead = c->d_pwLead.p;
    ws.trace = c->d_pwTrace.p;
    ws.runs = c->d_slot_runs.p;
    PoaBatch b{};
    b.reads = c->d_reads.p;
    b.read_off = c->d_read_off.p;
    b.n_reads = c->n_reads;
    b.order = c->d_order.p;
    b.results = c->d_results.p;
    b.out_runs = c->d_out_runs.p;
    b.out_run_cap = out_runs_cap;
    b.counters = c->d_counters.p;
    cudaMemsetAsync(c->d_counters.p, 0, 4 * sizeof(unsigned long long), c->stream);
    cudaEventRecord(c->ev0, c->stream);
    int rc = launch_pathwise(mode, c->dpg, c->dpg_rev, c->ds, ws, rw, b, (int)slots, c->stream);
    cudaEventRecord(c->ev1, c->stream);
    if (rc != 0 || cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("pathwise kernel");
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->kernel_ms += ms;
    c->launches += 1;
    c->slots_used = slots;
    cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("result copy");
    c->n_runs_total = std::min<uint64_t>(c->h_counters.p[1], out_runs_cap);
    return RG_OK;
}

// Modes 6 / 7: affine-gap pathwise alignment (experimental in the reference), pathwise_gap.cu.
static int align_pathwise_gap(rg_ctx* c, int mode) {
    if (!c->has_path_graph)
        return c->fail(c->path_graph_error.find("panics") != std::string::npos ? RG_ERR_REF_PANIC : RG_ERR_INVALID,
                       c->path_graph_error.empty() ? "no path graph" : c->path_graph_error);
    const FlatGraph& f = c->fg;
    const uint32_t n = f.n;
    if (f.P > 128) return c->fail(RG_ERR_UNSUPPORTED, "more than 128 paths are not supported on the device yet");
    release_poa_ws(c);
    PwGapWorkspace ws{};
    ws.Lp = c->max_len + 1;
    ws.Pp = f.PW * 32;
    ws.run_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4096, (uint64_t)n + 2 * ws.Lp), 1u << 22);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += c->d_gapT.cap * 4 + (c->d_slot_runs.cap + c->d_out_runs.cap) * sizeof(rg_run);
    const size_t per_slot = (size_t)3 * n * ws.Lp * ws.Pp * 4 + (size_t)ws.run_cap * sizeof(rg_run);
    const size_t budget_all = (size_t)(free_b * 0.85);
    size_t out_runs_cap = std::min<uint64_t>((uint64_t)c->n_reads * std::min<uint32_t>(ws.run_cap, 16384), (budget_all / 8) / sizeof(rg_run));
    out_runs_cap = std::max<size_t>(out_runs_cap, 1024);
    const size_t budget = budget_all - out_runs_cap * sizeof(rg_run);
    uint32_t slots = std::min<uint32_t>((uint32_t)c->sms * 4, (uint32_t)c->n_reads);
    slots = (uint32_t)std::min<size_t>(slots, budget / per_slot);
    if (slots < 1) return c->fail(RG_ERR_NOMEM, "not enough device memory for the three n x L x P tensors of one read (modes 6 / 7 keep the reference's footprint)");
    ws.slots = slots;
    if (!(c->d_gapT.ensure((size_t)slots * 3 * n * ws.Lp * ws.Pp) && c->d_slot_runs.ensure((size_t)slots * ws.run_cap) &&
          c->d_out_runs.ensure(out_runs_cap) && c->d_results.ensure(c->n_reads + 1) && c->d_counters.ensure(4)))
        return c->fail(RG_ERR_NOMEM, "device workspace allocation failed");
    ws.T = c->d_gapT.p;
    ws.runs = c->d_slot_runs.p;
    PoaBatch b{};
    b.reads = c->d_reads.p;
    b.read_off = c->d_read_off.p;
    b.n_reads = c->n_reads;
    b.order = c->d_order.p;
    b.results = c->d_results.p;
    b.out_runs = c->d_out_runs.p;
    b.out_run_cap = out_runs_cap;
    b.counters = c->d_counters.p;
    cudaMemsetAsync(c->d_counters.p, 0, 4 * sizeof(unsigned long long), c->stream);
    cudaEventRecord(c->ev0, c->stream);
    int rc = launch_pathwise_gap(mode, c->dpg, c->ds, ws, b, (int)slots, c->stream);
    cudaEventRecord(c->ev1, c->stream);
    if (rc != 0 || cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("pathwise (affine) kernel");
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->kernel_ms += ms;
    c->launches += 1;
    c->slots_used = slots;
    cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("result copy");
    c->n_runs_total = std::min<uint64_t>(c->h_counters.p[1], out_runs_cap);
    return RG_OK;
}

int rg_align_staged(rg_ctx* c, int mode) {
    if (!c) return RG_ERR_INVALID;
    if (!c->has_graph) return c->fail(RG_ERR_INVALID, "no graph loaded");
    cudaSetDevice(c->device);
    c->results_ready = false;
    c->kernel_ms = 0;
    c->launches = 0;
    c->cells = 0;
    c->n_runs_total = 0;
    c->last_mode = mode;
    if (c->n_reads == 0) {
        c->results_ready = true;
        return RG_OK;
    }
    int rc;
    switch (mode) {
        case RG_MODE_GLOBAL:
        case RG_MODE_LOCAL:
        case RG_MODE_GAP_LOCAL:
        case RG_MODE_GLOBAL_SCALAR:
        case RG_MODE_LOCAL_SCALAR:
        case RG_MODE_GAP_GLOBAL: rc = align_poa(c, mode); break;
        case RG_MODE_PATHWISE_GLOBAL:
        case RG_MODE_PATHWISE_SEMIGLOBAL:
        case RG_MODE_REC_GLOBAL:
        case RG_MODE_REC_SEMIGLOBAL: rc = align_pathwise(c, mode); break;
        case RG_MODE_PATHWISE_GAP_GLOBAL:
        case RG_MODE_PATHWISE_GAP_SEMIGLOBAL: rc = align_pathwise_gap(c, mode); break;
        default: return c->fail(RG_ERR_UNSUPPORTED, "alignment mode not implemented on the device yet");
    }
    if (rc != RG_OK) return rc;
    c->results_ready = true;
    return RG_OK;
}

int rg_fetch_results(rg_ctx* c, rg_batch_result* out) {
    if (!c || !out) return RG_ERR_INVALID;
    if (!c->results_ready) return c->fail(RG_ERR_INVALID, "no results to fetch");
    cudaSetDevice(c->device);
    memset(out, 0, sizeof *out);
    out->n_reads = c->n_reads;
    if (c->n_reads) {
        if (!c->h_results.ensure(c->n_reads + 1) || !c->h_runs.ensure(c->n_runs_total + 1))
            return c->fail(RG_ERR_NOMEM, "pinned allocation failed");
        cudaMemcpyAsync(c->h_results.p, c->d_results.p, (size_t)c->n_reads * sizeof(rg_read_result), cudaMemcpyDeviceToHost, c->stream);
        if (c->n_runs_total)
            cudaMemcpyAsync(c->h_runs.p, c->d_out_runs.p, c->n_runs_total * sizeof(rg_run), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return c->cuda_fail("result fetch");
        uint64_t cells = 0;
        for (int32_t i = 0; i < c->n_reads; i++) cells += c->h_results.p[i].cells;
        c->cells = cells;
    }
    out->reads = c->h_results.p;
    out->runs = c->h_runs.p;
    out->n_runs_total = c->n_runs_total;
    out->kernel_ms = c->kernel_ms;
    out->gpu_launches = c->launches;
    return RG_OK;
}

int rg_align_batch(rg_ctx* c, int mode, int32_t n_reads, const uint8_t* read_codes, const uint64_t* read_off,
                   rg_batch_result* out) {
    int rc = rg_upload_reads(c, n_reads, read_codes, read_off);
    if (rc != RG_OK) return rc;
    rc = rg_align_staged(c, mode);
    if (rc != RG_OK) return rc;
    return rg_fetch_results(c, out);
}

int rg_last_kernel_stats(const rg_ctx* c, double* kernel_ms, uint64_t* launches, uint64_t* cells) {
    if (!c) return RG_ERR_INVALID;
    if (kernel_ms) *kernel_ms = c->kernel_ms;
    if (launches) *launches = c->launches;
    if (cells) *cells = c->cells;
    return RG_OK;
}

int64_t rg_format_gaf(rg_ctx* c, int mode, const rg_batch_result* res, int32_t read_index, const char* read_name,
                      uint32_t read_len, int amb_mode, char* buf, size_t cap) {
    if (!c || !res || read_index < 0 || read_index >= res->n_reads || !c->has_graph) return RG_ERR_INVALID;
    std::string s;
    format_gaf(c->fg, mode, res->reads[read_index], res->runs, read_name ? read_name : "", read_len, amb_mode == 1 ? (RG_AMB_STRAND | RG_AMB_HANDLES) : amb_mode, s);
    if (buf && cap) {
        size_t k = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), k);
        buf[k] = 0;
    }
    return (int64_t)s.size();
}

int rg_format_gaf_all(rg_ctx* c, int mode, const rg_batch_result* res, const char* const* names, int64_t first_index,
                      const uint64_t* read_off, int amb_mode, char** out_text, size_t* out_len) {
    if (!c || !res || !read_off || !out_text || !c->has_graph) return RG_ERR_INVALID;
    // records are independent: contiguous chunks of reads are formatted by one host thread each and concatenated in order
    const int32_t n = res->n_reads;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hw, (int64_t)32, (int64_t)(n / 64 + 1)}));
    std::vector<std::string> part(T);
    const int flags = amb_mode == 1 ? (RG_AMB_STRAND | RG_AMB_HANDLES) : amb_mode;
    auto work = [&](int t) {
        const int32_t lo = (int32_t)((int64_t)n * t / T), hi = (int32_t)((int64_t)n * (t + 1) / T);
        std::string& s = part[t];
        s.reserve((size_t)(hi - lo) * 512);
        char nm[32];
        for (int32_t i = lo; i < hi; i++) {
            const char* name = names ? names[i] : nm;
            if (!names) snprintf(nm, sizeof nm, "read%lld", (long long)(first_index + i));
            format_gaf(c->fg, mode, res->reads[i], res->runs, name, (uint32_t)(read_off[i + 1] - read_off[i]), flags, s);
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    size_t total = 0;
    for (auto& s : part) total += s.size();
    char* p = (char*)malloc(total + 1);
    if (!p) return RG_ERR_NOMEM;
    size_t o = 0;
    for (auto& s : part) {
        memcpy(p + o, s.data(), s.size());
        o += s.size();
    }
    p[total] = 0;
    *out_text = p;
    if (out_len) *out_len = total;
    return RG_OK;
}

static int fill_reads(std::vector<std::string>& names, std::vector<uint8_t>& codes, std::vector<uint64_t>& off, rg_reads* out) {
    out->n_reads = (int32_t)names.size();
    out->codes = (uint8_t*)malloc(codes.size() + 1);
    out->off = (uint64_t*)malloc(off.size() * sizeof(uint64_t));
    out->names = (char**)malloc((names.size() + 1) * sizeof(char*));
    if (!out->codes || !out->off || !out->names) return RG_ERR_NOMEM;
    memcpy(out->codes, codes.data(), codes.size());
    memcpy(out->off, off.data(), off.size() * sizeof(uint64_t));
    for (size_t i = 0; i < names.size(); i++) out->names[i] = strdup(names[i].c_str());
    return RG_OK;
}

int rg_read_fasta_text(const char* text, size_t len, rg_reads* out, char* errbuf, size_t errcap) {
    if (!text || !out) return RG_ERR_INVALID;
    memset(out, 0, sizeof *out);
    std::vector<std::string> names;
    std::vector<uint8_t> codes;
    std::vector<uint64_t> off;
    std::string err;
    int status = RG_OK;
    if (!parse_fasta(text, len, names, codes, off, err, &status)) {
        if (errbuf && errcap) snprintf(errbuf, errcap, "%s", err.c_str());
        return status;
    }
    return fill_reads(names, codes, off, out);
}

int rg_read_fasta_file(const char* path, rg_reads* out, char* errbuf, size_t errcap) {
    if (!path || !out) return RG_ERR_INVALID;
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        if (errbuf && errcap) snprintf(errbuf, errcap, "cannot open %s", path);
        return RG_ERR_IO;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    return rg_read_fasta_text(s.data(), s.size(), out, errbuf, errcap);
}

void rg_free_reads(rg_reads* r) {
    if (!r) return;
    free(r->codes);
    free(r->off);
    if (r->names) {
        for (int32_t i = 0; i < r->n_reads; i++) free(r->names[i]);
        free(r->names);
    }
    memset(r, 0, sizeof *r);
}

void rg_free(void* p) { free(p); }

// ---- host-side diagnostics (no device needed): the flattened graph as a text dump (layout documented in the header),
// so that the host builders (graph.rs:31-123, utils.rs:103-165, pathwise_graph.rs:135-354) can be checked on a CPU box
static char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}
static bool flat_from_text(const char* text, size_t len, FlatGraph& f, std::string& err) {
    GfaGraph g;
    if (!parse_gfa(text, len, g, err)) return false;
    return flatten_graph(g, f, err) == RG_OK;
}

char* rg_debug_dump_lnz(const char* gfa_text, size_t len) {
    if (!gfa_text) return nullptr;
    FlatGraph f;
    std::string err;
    if (!flat_from_text(gfa_text, len, f, err)) return dup_string("ERROR " + err);
    std::string o = "lnz=";
    for (uint32_t i = 0; i < f.n; i++) o += CODE_CHARS[f.lnz[i]];
    o += "\nnwp=";
    for (uint32_t i = 0; i < f.n; i++) o += f.nwp[i] ? '1' : '0';
    o += "\n";
    for (uint32_t i = 0; i < f.n; i++)
        if (f.pred_off[i + 1] > f.pred_off[i]) {
            o += "pred " + std::to_string(i) + ":";
            for (uint32_t k = f.pred_off[i]; k < f.pred_off[i + 1]; k++) o += " " + std::to_string(f.pred_idx[k]);
            o += "\n";
        }
    for (uint32_t i = 0; i + 1 < f.n; i++)
        o += "hofp " + std::to_string(i) + ": " + (i == 0 ? std::string("-1") : std::to_string(f.row_seg_id[i])) + "\n";
    o += "r_values=";
    for (uint32_t i = 0; i < f.n; i++) o += (i ? "," : "") + std::to_string((long)f.r_values[i]);
    o += "\n";
    return dup_string(o);
}

char* rg_debug_dump_pathgraph(const char* gfa_text, size_t len, int reverse_graph) {
    if (!gfa_text) return nullptr;
    FlatGraph f;
    std::string err;
    if (!flat_from_text(gfa_text, len, f, err)) return dup_string("ERROR " + err);
    if (!f.has_paths) return dup_string("ERROR the graph has no paths");
    const std::vector<uint8_t>& nwp = reverse_graph ? f.rv_nwp : f.pw_nwp;
    const std::vector<uint32_t>& poff = reverse_graph ? f.rv_pred_off : f.pw_pred_off;
    const std::vector<uint32_t>& pidx = reverse_graph ? f.rv_pred_idx : f.pw_pred_idx;
    const std::vector<uint32_t>& ebits = reverse_graph ? f.rv_edge_bits : f.pw_edge_bits;
    auto bits = [&](const uint32_t* w) {
        std::string b;
        for (uint32_t k = 0; k < f.P; k++) b += ((w[k / 32] >> (k % 32)) & 1u) ? '1' : '0';
        return b;
    };
    std::string o = "paths_number=" + std::to_string(f.P) + "\nlnz=";
    for (uint32_t i = 0; i < f.n; i++) o += CODE_CHARS[f.lnz[i]];
    o += "\nnwp=";
    for (uint32_t i = 0; i < f.n; i++) o += nwp[i] ? '1' : '0';
    o += "\n";
    for (uint32_t i = 0; i < f.n; i++) {
        o += "node " + std::to_string(i) + ": id=" + std::to_string(f.row_seg_id[i]) + " alpha=" + std::to_string(f.alphas[i]) +
             " paths=" + bits(&f.node_bits[(size_t)i * f.PW]) + "\n";
        for (uint32_t k = poff[i]; k < poff[i + 1]; k++)
            o += "pred " + std::to_string(i) + " " + std::to_string(pidx[k]) + " " + bits(&ebits[(size_t)k * f.PW]) + "\n";
    }
    o += "dfs=";
    for (uint32_t i = 0; i < f.n; i++) o += (i ? "," : "") + std::to_string(f.dfs[i]);
    o += "\ndfe=";
    for (uint32_t i = 0; i < f.n; i++) o += (i ? "," : "") + std::to_string(f.dfe[i]);
    o += "\n";
    return dup_string(o);
}

int rg_int_peak(rg_ctx* c, double* iadd, double* imnmx, double* viaddmnmx) {
    if (!c || !iadd || !imnmx || !viaddmnmx) return RG_ERR_INVALID;
    cudaSetDevice(c->device);
    return launch_int_peak(iadd, imnmx, viaddmnmx, c->stream) == 0 ? RG_OK : RG_ERR_CUDA;
}

}  // extern "C"
