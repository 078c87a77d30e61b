// GFA1 ingestion and flattening into device-ready arrays.
// What is computed follows graph.rs:31-123 (LnzGraph), utils.rs:103-126,144-165 (r values, row -> segment id),
// pathwise_graph.rs:135-354 (PathGraph, reverse graph, distances); how it is laid out is ours: flat CSR arrays
// in topological row order, predecessor list order preserved, per-edge path bitsets packed in 32-bit words.
#include <algorithm>
#include <charconv>
#include <cstring>
#include <unordered_map>

#include "host.h"

namespace rg {

static inline int code_of(char c) {
    switch (c) {
        case 'A': return CODE_A;
        case 'C': return CODE_C;
        case 'G': return CODE_G;
        case 'T': return CODE_T;
        case 'N': return CODE_N;
        default: return -1;
    }
}

namespace {
struct Cursor {
    const char* p;
    const char* end;
};
// next tab-separated field of the current line
inline bool field(const char*& p, const char* eol, const char*& fb, const char*& fe) {
    if (p > eol) return false;
    fb = p;
    const char* t = (const char*)memchr(p, '\t', (size_t)(eol - p));
    fe = t ? t : eol;
    p = fe + 1;
    return true;
}
inline bool parse_u64(const char* b, const char* e, uint64_t& v) {
    if (b == e) return false;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}
}  // namespace

bool parse_gfa(const char* text, size_t len, GfaGraph& g, std::string& err) {
    struct Seg {
        uint64_t id;
        const char* sb;
        const char* se;
    };
    struct Link {
        uint64_t from, to;
    };
    struct PathRef {
        const char* nb;
        const char* ne;
        const char* sb;
        const char* se;
    };
    std::vector<Seg> segs;
    std::vector<Link> links;
    std::vector<PathRef> paths;
    const char* p = text;
    const char* end = text + len;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* le = eol;
        if (le > p && le[-1] == '\r') le--;
        if (le > p) {
            const char* q = p;
            const char *fb, *fe;
            field(q, le, fb, fe);
            if (fe - fb == 1) {
                char t = *fb;
                if (t == 'S') {
                    const char *nb, *ne, *sb, *se;
                    if (!field(q, le, nb, ne) || !field(q, le, sb, se)) {
                        err = "GFA: malformed S line";
                        return false;
                    }
                    uint64_t id;
                    if (!parse_u64(nb, ne, id)) {
                        err = "GFA: segment name is not an unsigned integer: " + std::string(nb, ne);
                        return false;
                    }
                    segs.push_back({id, sb, se});
                } else if (t == 'L') {
                    const char *a, *ae, *o1, *o1e, *b, *be, *o2, *o2e;
                    if (!field(q, le, a, ae) || !field(q, le, o1, o1e) || !field(q, le, b, be) ||
                        !field(q, le, o2, o2e)) {
                        err = "GFA: malformed L line";
                        return false;
                    }
                    uint64_t from, to;
                    if (!parse_u64(a, ae, from) || !parse_u64(b, be, to)) {
                        err = "GFA: link endpoint is not an unsigned integer";
                        return false;
                    }
                    if (*o1 != '+' || *o2 != '+') {
                        err = "GFA: reverse-orientation links are outside the supported domain";
                        return false;
                    }
                    links.push_back({from, to});
                } else if (t == 'P') {
                    const char *nb, *ne, *sb, *se;
                    if (!field(q, le, nb, ne) || !field(q, le, sb, se)) {
                        err = "GFA: malformed P line";
                        return false;
                    }
                    paths.push_back({nb, ne, sb, se});
                }
            }
        }
        p = eol + 1;
    }
    // handles_iter().collect(); sort()  (graph.rs:32-33)
    std::vector<uint32_t> order(segs.size());
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return segs[a].id < segs[b].id; });
    std::unordered_map<uint64_t, uint32_t> index;
    index.reserve(segs.size() * 2);
    g.seg_id.clear();
    g.seg_seq.clear();
    for (uint32_t k = 0; k < order.size(); k++) {
        const Seg& s = segs[order[k]];
        if (index.count(s.id)) {  // HashMap insert: the later S line replaces the earlier one
            g.seg_seq[index[s.id]] = std::string(s.sb, s.se);
            continue;
        }
        index[s.id] = (uint32_t)g.seg_id.size();
        g.seg_id.push_back(s.id);
        g.seg_seq.emplace_back(s.sb, s.se);
    }
    g.seg_preds.assign(g.seg_id.size(), {});
    std::vector<std::vector<uint32_t>> succ(g.seg_id.size());
    for (auto& l : links) {
        auto fi = index.find(l.from), ti = index.find(l.to);
        if (fi == index.end() || ti == index.end()) {
            err = "GFA: link references an unknown segment";
            return false;
        }
        auto& sv = succ[fi->second];
        if (std::find(sv.begin(), sv.end(), ti->second) != sv.end()) continue;  // HashGraph::create_edge de-dups
        sv.push_back(ti->second);
        g.seg_preds[ti->second].push_back(fi->second);
    }
    g.paths.clear();
    g.path_names.clear();
    for (auto& pr : paths) {
        std::vector<uint32_t> steps;
        const char* s = pr.sb;
        while (s < pr.se) {
            const char* c = (const char*)memchr(s, ',', (size_t)(pr.se - s));
            const char* e = c ? c : pr.se;
            if (e > s) {
                char o = e[-1];
                uint64_t id;
                if ((o != '+' && o != '-') || !parse_u64(s, e - 1, id)) {
                    err = "GFA: malformed path step";
                    return false;
                }
                if (o == '-') {
                    err = "GFA: reverse-orientation path steps are outside the supported domain";
                    return false;
                }
                auto it = index.find(id);
                if (it == index.end()) {
                    err = "GFA: path step on an unknown segment";
                    return false;
                }
                steps.push_back(it->second);
            }
            s = e + 1;
        }
        g.paths.push_back(std::move(steps));
        g.path_names.emplace_back(pr.nb, pr.ne);
    }
    return true;
}

// r values (utils.rs:103-126) and the per-row helper arrays shared by both builders.
static void finish_lnz(FlatGraph& f) {
    const uint32_t n = f.n;
    f.min_pred.assign(n, 0);
    f.min_pred_slot.assign(n, 0);
    f.prev_slot.assign(n, PREV_ALWAYS);
    f.is_pred_row.assign(n, 0);
    f.max_indeg = 1;
    f.max_lookback = 1;
    for (uint32_t i = 1; i < n; i++) {
        uint32_t b = f.pred_off[i], e = f.pred_off[i + 1];
        if (!f.nwp[i] || b == e) {
            f.min_pred[i] = i - 1;
            f.prev_slot[i] = f.nwp[i] ? PREV_NONE : PREV_ALWAYS;
            continue;
        }
        uint32_t mp = f.pred_idx[b], ms = 0;
        uint8_t ps = PREV_NONE;
        for (uint32_t k = b; k < e; k++) {
            uint32_t p = f.pred_idx[k];
            if (p < mp) {
                mp = p;
                ms = k - b;
            }
            if (p == i - 1 && ps == PREV_NONE && k - b < 0xFE) ps = (uint8_t)(k - b);
            if (i != n - 1) f.is_pred_row[p] = 1;
        }
        f.min_pred[i] = mp;
        f.min_pred_slot[i] = (uint8_t)std::min<uint32_t>(ms, 255);
        f.prev_slot[i] = ps;
        f.max_indeg = std::max(f.max_indeg, e - b);
        if (i != n - 1 && i > mp) f.max_lookback = std::max(f.max_lookback, i - mp);
    }
    // set_r_values
    std::vector<int64_t> r(n, -1);
    r[n - 1] = 0;
    for (uint32_t k = f.pred_off[n - 1]; k < f.pred_off[n]; k++) r[f.pred_idx[k]] = 0;
    for (uint32_t i = n - 2; i >= 1; i--) {
        if (r[i] == -1 || r[i] > r[i + 1] + 1) r[i] = r[i + 1] + 1;
        if (f.nwp[i])
            for (uint32_t k = f.pred_off[i]; k < f.pred_off[i + 1]; k++) {
                uint32_t p = f.pred_idx[k];
                if (r[p] == -1 || r[p] > r[i] + 1) r[p] = r[i] + 1;
            }
    }
    f.r_values.resize(n);
    for (uint32_t i = 0; i < n; i++) f.r_values[i] = (int32_t)r[i];
}

int flat_from_lnz(uint32_t n, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                  const uint32_t* pred_idx, const uint64_t* seg_id, FlatGraph& f, std::string& err) {
    if (n < 3) {
        err = "graph needs at least one character between '$' and 'F'";
        return RG_ERR_INVALID;
    }
    f = FlatGraph();
    f.n = n;
    f.lnz.assign(lnz_codes, lnz_codes + n);
    f.lnz[0] = CODE_START;
    f.lnz[n - 1] = CODE_END;
    for (uint32_t i = 1; i + 1 < n; i++)
        if (f.lnz[i] > CODE_N) {
            err = "graph character outside A,C,G,T,N";
            return RG_ERR_BAD_CHAR;
        }
    f.nwp.assign(nwp, nwp + n);
    f.pred_off.assign(pred_off, pred_off + n + 1);
    f.pred_idx.assign(pred_idx, pred_idx + pred_off[n]);
    for (uint32_t i = 0; i < n; i++) {
        if (f.pred_off[i + 1] < f.pred_off[i]) {
            err = "pred_off is not monotone";
            return RG_ERR_INVALID;
        }
        for (uint32_t k = f.pred_off[i]; k < f.pred_off[i + 1]; k++)
            if (f.pred_idx[k] >= i) {
                err = "predecessor index is not smaller than its node (rows must be topologically ordered)";
                return RG_ERR_INVALID;
            }
        if (f.nwp[i] && f.pred_off[i + 1] == f.pred_off[i] && i != 0) {
            err = "nwp row without predecessors (reference: pred_hash.get().unwrap() panic)";
            return RG_ERR_REF_PANIC;
        }
    }
    if (!f.nwp[n - 1] || f.pred_off[n] == f.pred_off[n - 1]) {
        err = "last row ('F') needs predecessors";
        return RG_ERR_REF_PANIC;
    }
    if (!f.nwp[1]) {
        err = "row 1 must start a segment (nwp)";
        return RG_ERR_INVALID;
    }
    // utils.rs:152-161: a new handle starts at every nwp row
    f.row_seg.assign(n, UINT32_MAX);
    f.row_seg_id.assign(n, 0);
    f.seg_first_row.clear();
    uint32_t cur = UINT32_MAX;
    for (uint32_t i = 1; i + 1 < n; i++) {
        if (f.nwp[i]) {
            cur = (uint32_t)f.seg_first_row.size();
            f.seg_first_row.push_back(i);
        }
        f.row_seg[i] = cur;
        f.row_seg_id[i] = seg_id ? seg_id[i] : (uint64_t)cur + 1;
    }
    f.n_segments = (uint32_t)f.seg_first_row.size();
    f.seg_ids.resize(f.n_segments);
    for (uint32_t k = 0; k < f.n_segments; k++) f.seg_ids[k] = f.row_seg_id[f.seg_first_row[k]];
    finish_lnz(f);
    return RG_OK;
}

// Reverse graph (pathwise_graph.rs:250-282: every (node, pred, paths) entry transposed), look-backs and the distance vectors of
// pathwise_graph.rs:306-354, from the forward PathGraph CSR (entries of a node in ascending predecessor order).
static void finish_path_graph(FlatGraph& f) {
    const uint32_t n = f.n, PW = f.PW;
    {
        std::vector<uint32_t> cnt(n + 1, 0);
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t k = f.pw_pred_off[i]; k < f.pw_pred_off[i + 1]; k++) cnt[f.pw_pred_idx[k] + 1]++;
        f.rv_pred_off.assign(n + 1, 0);
        for (uint32_t i = 0; i < n; i++) f.rv_pred_off[i + 1] = f.rv_pred_off[i] + cnt[i + 1];
        f.rv_pred_idx.assign(f.pw_pred_idx.size(), 0);
        f.rv_edge_bits.assign(f.pw_edge_bits.size(), 0);
        std::vector<uint32_t> fill(f.rv_pred_off.begin(), f.rv_pred_off.end() - 1);
        for (uint32_t i = 0; i < n; i++)   // ascending node => ascending "predecessor" inside every reverse list
            for (uint32_t k = f.pw_pred_off[i]; k < f.pw_pred_off[i + 1]; k++) {
                const uint32_t p = f.pw_pred_idx[k], dst = fill[p]++;
                f.rv_pred_idx[dst] = i;
                for (uint32_t w = 0; w < PW; w++) f.rv_edge_bits[(size_t)dst * PW + w] = f.pw_edge_bits[(size_t)k * PW + w];
            }
    }
    f.rv_nwp.assign(n, 0);
    for (uint32_t i = 0; i < n; i++)
        if (f.rv_pred_off[i + 1] > f.rv_pred_off[i]) f.rv_nwp[i] = 1;
    f.pw_max_lookback = 1;
    f.rv_max_lookback = 1;
    for (uint32_t i = 1; i + 1 < n; i++) {
        for (uint32_t k = f.pw_pred_off[i]; k < f.pw_pred_off[i + 1]; k++)
            f.pw_max_lookback = std::max(f.pw_max_lookback, i - f.pw_pred_idx[k]);
        for (uint32_t k = f.rv_pred_off[i]; k < f.rv_pred_off[i + 1]; k++)
            f.rv_max_lookback = std::max(f.rv_max_lookback, f.rv_pred_idx[k] - i);
    }
    // distances (pathwise_graph.rs:306-354)
    {
        std::vector<int64_t> r(n, -1);
        r[0] = 0;
        for (uint32_t k = f.rv_pred_off[0]; k < f.rv_pred_off[1]; k++) r[f.rv_pred_idx[k]] = 1;
        for (uint32_t i = 1; i + 1 < n; i++) {
            if (r[i] == -1 || r[i] > r[i - 1] + 1) r[i] = r[i - 1] + 1;
            if (f.rv_nwp[i])
                for (uint32_t k = f.rv_pred_off[i]; k < f.rv_pred_off[i + 1]; k++) {
                    uint32_t p = f.rv_pred_idx[k];
                    if (r[p] == -1 || r[p] > r[i] + 1) r[p] = r[i] + 1;
                }
        }
        f.dfs.resize(n);
        for (uint32_t i = 0; i < n; i++) f.dfs[i] = (int32_t)r[i];
    }
    {
        std::vector<int64_t> r(n, -1);
        r[n - 1] = 0;
        for (uint32_t k = f.pw_pred_off[n - 1]; k < f.pw_pred_off[n]; k++) r[f.pw_pred_idx[k]] = 1;
        for (uint32_t i = n - 2; i >= 1; i--) {
            if (r[i] == -1 || r[i] > r[i + 1] + 1) r[i] = r[i + 1] + 1;
            if (f.pw_nwp[i])
                for (uint32_t k = f.pw_pred_off[i]; k < f.pw_pred_off[i + 1]; k++) {
                    uint32_t p = f.pw_pred_idx[k];
                    if (r[p] == -1 || r[p] > r[i] + 1) r[p] = r[i] + 1;
                }
        }
        f.dfe.resize(n);
        for (uint32_t i = 0; i < n; i++) f.dfe[i] = (int32_t)r[i];
    }
}

// A prebuilt PathGraph (pathwise_graph.rs:10-18): lnz, nwp, PredHash as CSR with one path bitset per (node, predecessor)
// entry, paths_nodes, alphas, nodes_id_pos. The LnzGraph side of the FlatGraph is filled from the same predecessor lists.
int flat_from_path_graph(uint32_t n, uint32_t P, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                         const uint32_t* pred_idx, const uint32_t* edge_bits, const uint32_t* node_bits, const uint32_t* alphas,
                         const uint64_t* seg_id, FlatGraph& f, std::string& err) {
    if (P == 0) {
        err = "a PathGraph needs at least one path";
        return RG_ERR_INVALID;
    }
    int rc = flat_from_lnz(n, lnz_codes, nwp, pred_off, pred_idx, seg_id, f, err);
    if (rc != RG_OK) return rc;
    const uint32_t PW = (P + 31) / 32;
    f.P = P;
    f.PW = PW;
    f.has_paths = true;
    f.pw_nwp.assign(nwp, nwp + n);
    f.node_bits.assign(node_bits, node_bits + (size_t)n * PW);
    f.alphas.assign(alphas, alphas + n);
    // entries of a node in ascending predecessor order (the fixed order of this implementation)
    f.pw_pred_off.assign(pred_off, pred_off + n + 1);
    f.pw_pred_idx.clear();
    f.pw_edge_bits.clear();
    for (uint32_t i = 0; i < n; i++) {
        std::vector<uint32_t> ord;
        for (uint32_t k = pred_off[i]; k < pred_off[i + 1]; k++) ord.push_back(k);
        std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return pred_idx[a] < pred_idx[b]; });
        for (uint32_t k : ord) {
            f.pw_pred_idx.push_back(pred_idx[k]);
            for (uint32_t w = 0; w < PW; w++) f.pw_edge_bits.push_back(edge_bits[(size_t)k * PW + w]);
        }
    }
    finish_path_graph(f);
    return RG_OK;
}

int flatten_graph(const GfaGraph& g, FlatGraph& f, std::string& err) {
    f = FlatGraph();
    const uint32_t S = (uint32_t)g.seg_id.size();
    if (S == 0) {
        err = "GFA has no segments";
        return RG_ERR_IO;
    }
    uint64_t total = 0;
    for (auto& s : g.seg_seq) {
        if (s.empty()) {
            err = "GFA: empty segment sequence";
            return RG_ERR_IO;
        }
        total += s.size();
    }
    if (total + 2 > 0x7fffffffull) {
        err = "graph too large";
        return RG_ERR_UNSUPPORTED;
    }
    const uint32_t n = (uint32_t)total + 2;
    f.n = n;
    f.lnz.resize(n);
    f.nwp.assign(n, 0);
    f.row_seg.assign(n, UINT32_MAX);
    f.row_seg_id.assign(n, 0);
    f.seg_first_row.resize(S);
    f.n_segments = S;
    f.seg_ids.assign(g.seg_id.begin(), g.seg_id.end());
    std::vector<uint32_t> seg_last(S);
    f.lnz[0] = CODE_START;
    uint32_t row = 1;
    for (uint32_t s = 0; s < S; s++) {
        f.seg_first_row[s] = row;
        for (char c : g.seg_seq[s]) {
            int code = code_of(c);
            if (code < 0) {
                err = std::string("graph character outside A,C,G,T,N: '") + c + "'";
                return RG_ERR_BAD_CHAR;
            }
            f.lnz[row] = (uint8_t)code;
            f.row_seg[row] = s;
            f.row_seg_id[row] = g.seg_id[s];
            row++;
        }
        seg_last[s] = row - 1;
    }
    f.lnz[n - 1] = CODE_END;
    // predecessors (graph.rs:62-88) in CSR; F's predecessors = segments that precede nobody, ascending (graph.rs:112-123)
    std::vector<uint8_t> precedes(S, 0);
    f.pred_off.assign(n + 1, 0);
    for (uint32_t s = 0; s < S; s++) {
        uint32_t st = f.seg_first_row[s];
        f.nwp[st] = 1;
        f.pred_off[st + 1] = g.seg_preds[s].empty() ? 1 : (uint32_t)g.seg_preds[s].size();
        for (uint32_t p : g.seg_preds[s]) {
            if (p >= s) {
                err = "GFA: link goes from a higher to a lower (or equal) segment id — segment ids must be topologically ordered";
                return RG_ERR_UNSUPPORTED;
            }
            precedes[p] = 1;
        }
    }
    uint32_t n_last = 0;
    for (uint32_t s = 0; s < S; s++) n_last += !precedes[s];
    f.nwp[n - 1] = 1;
    f.pred_off[n] = n_last;
    for (uint32_t i = 0; i < n; i++) f.pred_off[i + 1] += f.pred_off[i];
    f.pred_idx.resize(f.pred_off[n]);
    for (uint32_t s = 0; s < S; s++) {
        uint32_t o = f.pred_off[f.seg_first_row[s]];
        if (g.seg_preds[s].empty())
            f.pred_idx[o] = 0;
        else
            for (uint32_t p : g.seg_preds[s]) f.pred_idx[o++] = seg_last[p];
    }
    {
        uint32_t o = f.pred_off[n - 1];
        for (uint32_t s = 0; s < S; s++)
            if (!precedes[s]) f.pred_idx[o++] = seg_last[s];
    }
    finish_lnz(f);

    // ---- PathGraph
    const uint32_t P = (uint32_t)g.paths.size();
    f.P = P;
    f.has_paths = P > 0;
    if (P == 0) return RG_OK;
    const uint32_t PW = (P + 31) / 32;
    f.PW = PW;
    f.node_bits.assign((size_t)n * PW, 0);
    f.alphas.assign(n, P + 1);
    f.pw_nwp.assign(n, 0);
    struct Edge {
        uint32_t node, pred, path;
    };
    std::vector<Edge> edges;
    std::vector<std::vector<uint32_t>> seg_paths(S);
    for (uint32_t p = 0; p < P; p++) {
        const auto& st = g.paths[p];
        for (size_t pos = 0; pos < st.size(); pos++) {
            uint32_t s = st[pos];
            if (pos > 0 && st[pos - 1] >= s) {
                err = "GFA: path is not increasing in segment id (graph must be a topologically ordered DAG)";
                return RG_ERR_UNSUPPORTED;
            }
            seg_paths[s].push_back(p);
            uint32_t hs = f.seg_first_row[s];
            f.pw_nwp[hs] = 1;
            if (pos == 0) {
                edges.push_back({hs, 0, p});
            } else {
                edges.push_back({hs, seg_last[st[pos - 1]], p});
                if (pos == st.size() - 1) edges.push_back({n - 1, seg_last[s], p});
            }
        }
    }
    for (uint32_t s = 0; s < S; s++)
        for (uint32_t p : seg_paths[s])
            for (uint32_t r = f.seg_first_row[s]; r <= seg_last[s]; r++) {
                f.node_bits[(size_t)r * PW + p / 32] |= 1u << (p % 32);
                if (f.alphas[r] == P + 1 || p < f.alphas[r]) f.alphas[r] = p;
            }
    for (uint32_t p = 0; p < P; p++) {
        f.node_bits[p / 32] |= 1u << (p % 32);
        f.node_bits[(size_t)(n - 1) * PW + p / 32] |= 1u << (p % 32);
    }
    f.alphas[0] = 0;
    f.alphas[n - 1] = 0;
    f.pw_nwp[n - 1] = 1;
    auto build_csr = [&](std::vector<Edge>& ev, std::vector<uint32_t>& off, std::vector<uint32_t>& idx,
                         std::vector<uint32_t>& bits) {
        std::sort(ev.begin(), ev.end(), [](const Edge& a, const Edge& b) {
            if (a.node != b.node) return a.node < b.node;
            if (a.pred != b.pred) return a.pred < b.pred;
            return a.path < b.path;
        });
        off.assign(n + 1, 0);
        idx.clear();
        bits.clear();
        size_t k = 0;
        while (k < ev.size()) {
            uint32_t node = ev[k].node, pred = ev[k].pred;
            idx.push_back(pred);
            size_t b0 = bits.size();
            bits.resize(b0 + PW, 0);
            while (k < ev.size() && ev[k].node == node && ev[k].pred == pred) {
                bits[b0 + ev[k].path / 32] |= 1u << (ev[k].path % 32);
                k++;
            }
            off[node + 1]++;
        }
        for (uint32_t i = 0; i < n; i++) off[i + 1] += off[i];
    };
    build_csr(edges, f.pw_pred_off, f.pw_pred_idx, f.pw_edge_bits);
    finish_path_graph(f);
    return RG_OK;
}

}  // namespace rg
