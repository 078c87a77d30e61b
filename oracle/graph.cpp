// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). Ingestion side of the restatement:
// graph.rs, pathwise_graph.rs, sequences.rs, score_matrix.rs, utils.rs (band + r values).
#include <algorithm>
#include <charconv>
#include <fstream>
#include <sstream>

#include "oracle.hpp"

namespace rgo {

// ------------------------------------------------------------------------------- score_matrix.rs
// score_matrix.rs:35-51
ScoreMatrix create_score_matrix_match_mis(int m, int x) {
    ScoreMatrix sm;
    const char al[6] = {'A', 'C', 'G', 'T', 'N', '-'};
    for (char i : al)
        for (char j : al) {
            if (i == j)
                sm.insert(i, j, m);
            else if (i == '-' || j == '-')
                sm.insert(i, j, x * 2);
            else
                sm.insert(i, j, x);
        }
    sm.insert('N', 'N', x);
    sm.remove('-', '-');
    return sm;
}
// score_matrix.rs:52-66 (the api.rs f32 builder: gap-vs-char is x, not 2x)
ScoreMatrix create_score_matrix_match_mis_f32(int m, int x) {
    ScoreMatrix sm;
    const char al[6] = {'A', 'C', 'G', 'T', 'N', '-'};
    for (char i : al)
        for (char j : al) sm.insert(i, j, i == j ? m : x);
    sm.insert('N', 'N', x);
    sm.remove('-', '-');
    return sm;
}
// score_matrix.rs:67-105; HOXD55.mtx / HOXD70.mtx contents embedded (rows = first key char, cols = second).
ScoreMatrix create_score_matrix_hoxd(const std::string& name) {
    static const int h55[5][5] = {{91, -90, -25, -100, 0},
                                  {-90, 100, -100, -25, 0},
                                  {-25, -100, 100, -90, 0},
                                  {-100, -25, -90, 91, 0},
                                  {0, 0, 0, 0, 0}};
    static const int h70[5][5] = {{91, -114, -31, -123, 0},
                                  {-114, 100, -125, -31, 0},
                                  {-31, -125, 100, -114, 0},
                                  {-123, -31, -144, 91, 0},
                                  {0, 0, 0, 0, 0}};
    const int(*t)[5];
    if (name == "HOXD55.mtx" || name == "HOXD55")
        t = h55;
    else if (name == "HOXD70.mtx" || name == "HOXD70")
        t = h70;
    else
        throw RefPanic("wrong matrix type");
    const char al[5] = {'A', 'C', 'G', 'T', 'N'};
    ScoreMatrix sm;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) sm.insert(al[i], al[j], t[i][j]);
    for (char c : al) {
        sm.insert(c, '-', -200);
        sm.insert('-', c, -200);
    }
    sm.remove('-', '-');
    return sm;
}

// ------------------------------------------------------------------------------- HashGraph stand-in
Handle HashGraph::append_handle(const std::string& seq) { return create_handle(seq, max_id + 1); }
Handle HashGraph::create_handle(const std::string& seq, uint64_t id) {
    if (seq.empty()) throw RefPanic("Tried to add empty handle");
    graph[id] = HGNode{seq, {}, {}};
    max_id = std::max(max_id, id);
    return Handle{id << 1};
}
// handlegraph hashgraph.rs create_edge: duplicate edges (same left->right) are ignored.
void HashGraph::create_edge(Handle left, Handle right) {
    auto li = graph.find(left.id());
    auto ri = graph.find(right.id());
    if (li == graph.end() || ri == graph.end()) throw RefPanic("Node doesn't exist for the given handle");
    auto& edges_of_left = left.is_reverse() ? li->second.left_edges : li->second.right_edges;
    Handle stored_r = left.is_reverse() ? right.flip() : right;
    if (std::find(edges_of_left.begin(), edges_of_left.end(), stored_r) != edges_of_left.end()) return;
    edges_of_left.push_back(stored_r);
    if (!(left == right.flip())) {
        if (right.is_reverse())
            ri->second.right_edges.push_back(left.flip());
        else
            ri->second.left_edges.push_back(left);
    }
}
size_t HashGraph::create_path_handle(const std::string& name) {
    paths.push_back(HGPath{name, {}});
    return paths.size() - 1;
}
void HashGraph::append_step(size_t path, Handle h) { paths[path].nodes.push_back(h); }
std::string HashGraph::sequence(Handle h) const {
    const std::string& s = graph.at(h.id()).sequence;
    if (!h.is_reverse()) return s;
    std::string r(s.rbegin(), s.rend());
    for (char& c : r) {
        switch (c) {
            case 'A': c = 'T'; break;
            case 'C': c = 'G'; break;
            case 'G': c = 'C'; break;
            case 'T': c = 'A'; break;
            case 'a': c = 't'; break;
            case 'c': c = 'g'; break;
            case 'g': c = 'c'; break;
            case 't': c = 'a'; break;
            default: break;
        }
    }
    return r;
}
std::vector<Handle> HashGraph::left_neighbours(Handle h) const {
    const HGNode& n = graph.at(h.id());
    std::vector<Handle> r;
    if (h.is_reverse())
        for (Handle x : n.right_edges) r.push_back(x.flip());
    else
        r = n.left_edges;
    return r;
}
// graph.rs:32-40 / graph.rs:128-143: handles_iter().collect(); sort(); if amb_mode { reverse(); flip }
std::vector<Handle> HashGraph::handles_sorted(bool amb_mode) const {
    std::vector<Handle> hs;
    for (auto& kv : graph) hs.push_back(Handle{kv.first << 1});
    std::sort(hs.begin(), hs.end());
    if (amb_mode) {
        std::reverse(hs.begin(), hs.end());
        for (auto& h : hs) h = h.flip();
    }
    return hs;
}

static std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t st = 0;
    while (true) {
        size_t p = s.find(sep, st);
        if (p == std::string::npos) {
            out.push_back(s.substr(st));
            break;
        }
        out.push_back(s.substr(st, p - st));
        st = p + 1;
    }
    return out;
}
static uint64_t parse_usize(const std::string& s) {
    uint64_t v = 0;
    auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    if (r.ec != std::errc() || r.ptr != s.data() + s.size()) throw RefPanic("GFA: segment name is not usize: " + s);
    return v;
}
// gfa ^0.8.0 GFAParser::parse_file into GFA<usize,()> then HashGraph::from_gfa: segments, then links
// in file order, then paths in file order. Unknown / malformed-optional lines are skipped.
HashGraph parse_gfa_text(const std::string& text) {
    struct Link {
        uint64_t from, to;
        bool frev, trev;
    };
    struct PathL {
        std::string name;
        std::vector<Handle> steps;
    };
    std::vector<std::pair<uint64_t, std::string>> segs;
    std::vector<Link> links;
    std::vector<PathL> paths;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        auto f = split(line, '\t');
        if (f[0] == "S") {
            if (f.size() < 3) throw RefPanic("GFA: bad S line");
            segs.push_back({parse_usize(f[1]), f[2]});
        } else if (f[0] == "L") {
            if (f.size() < 6) throw RefPanic("GFA: bad L line");
            links.push_back(Link{parse_usize(f[1]), parse_usize(f[3]), f[2] == "-", f[4] == "-"});
        } else if (f[0] == "P") {
            if (f.size() < 3) throw RefPanic("GFA: bad P line");
            PathL p;
            p.name = f[1];
            for (auto& st : split(f[2], ',')) {
                if (st.empty()) continue;
                char o = st.back();
                if (o != '+' && o != '-') throw RefPanic("GFA: bad path step");
                p.steps.push_back(Handle{(parse_usize(st.substr(0, st.size() - 1)) << 1) | (o == '-' ? 1u : 0u)});
            }
            paths.push_back(p);
        }
    }
    HashGraph g;
    for (auto& s : segs) g.create_handle(s.second, s.first);
    for (auto& l : links)
        g.create_edge(Handle{(l.from << 1) | (l.frev ? 1u : 0u)}, Handle{(l.to << 1) | (l.trev ? 1u : 0u)});
    for (auto& p : paths) {
        size_t id = g.create_path_handle(p.name);
        for (Handle h : p.steps) g.append_step(id, h);
    }
    return g;
}
HashGraph parse_gfa_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw RefPanic("GFA file open failed: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_gfa_text(ss.str());
}

// ------------------------------------------------------------------------------- graph.rs
// graph.rs:31-123
LnzGraph create_graph_struct(const HashGraph& graph, bool amb_mode) {
    std::vector<Handle> sorted_handles = graph.handles_sorted(amb_mode);
    long last_index = 1;
    std::map<uint64_t, long> visited_node, last_nodes;
    std::vector<char> linearization{'$'};
    for (Handle h : sorted_handles) {
        for (char c : graph.sequence(h)) {
            linearization.push_back(c);
            last_index++;
        }
        visited_node[h.id()] = last_index - 1;
        last_nodes[h.id()] = last_index - 1;
    }
    size_t n = linearization.size() + 1;
    LnzGraph lg;
    lg.nwp.assign(n, 0);
    lg.pred_hash.assign(n, {});
    for (Handle h : sorted_handles) {
        auto lefts = graph.left_neighbours(h);
        long h_last_idx = visited_node.at(h.id());
        size_t handle_start_pos = (size_t)h_last_idx - graph.sequence(h).size() + 1;
        if (lefts.empty()) {
            lg.nwp[handle_start_pos] = 1;
            lg.pred_hash[handle_start_pos].push_back(0);
        }
        for (Handle p : lefts) {
            long pred_last_idx = visited_node.at(p.id());
            last_nodes.erase(p.id());
            lg.nwp[handle_start_pos] = 1;
            lg.pred_hash[handle_start_pos].push_back((size_t)pred_last_idx);
        }
    }
    // set_last_node (graph.rs:112-123). HashMap order fixed to ascending node id (oracle.hpp header).
    linearization.push_back('F');
    lg.nwp[linearization.size() - 1] = 1;
    for (auto& kv : last_nodes) lg.pred_hash[linearization.size() - 1].push_back((size_t)kv.second);
    lg.lnz = linearization;
    return lg;
}

// utils.rs:144-198: lnz row -> segment id string, row 0 -> "-1"; no entry for the last row.
std::vector<std::string> handle_pos_in_lnz(const LnzGraph& lg, const HashGraph& g, bool amb_mode) {
    std::vector<Handle> sorted_handles = g.handles_sorted(amb_mode);
    std::vector<std::string> hofp(lg.nwp.size());
    long curr = 0;
    for (size_t i = 1; i + 1 < lg.nwp.size(); i++) {
        if (lg.nwp[i]) curr++;
        hofp[i] = std::to_string(sorted_handles.at((size_t)(curr - 1)).id());
    }
    hofp[0] = "-1";
    return hofp;
}

// ------------------------------------------------------------------------------- pathwise_graph.rs
static void set_preds_and_paths(PathGraph& pg, size_t curr_node, size_t pred_pos, size_t path_id) {
    auto& v = pg.pred_hash[curr_node];
    auto it = std::lower_bound(v.begin(), v.end(), pred_pos,
                               [](const std::pair<size_t, BitVec>& a, size_t b) { return a.first < b; });
    if (it == v.end() || it->first != pred_pos) it = v.insert(it, {pred_pos, BitVec(pg.paths_number, 0)});
    it->second[path_id] = 1;
}
// pathwise_graph.rs:135-248
PathGraph create_path_graph(const HashGraph& graph, bool is_reversed) {
    std::vector<Handle> sorted_handles = graph.handles_sorted(is_reversed);
    PathGraph pg;
    long last_index = 1;
    std::map<uint64_t, std::pair<long, long>> handles_id_position;
    pg.lnz.push_back('$');
    pg.nodes_id_pos.push_back(0);
    for (Handle h : sorted_handles) {
        long start_position = last_index;
        for (char ch : graph.sequence(h)) {
            pg.lnz.push_back(ch);
            pg.nodes_id_pos.push_back(h.id());
            last_index++;
        }
        handles_id_position[h.id()] = {start_position, last_index - 1};
    }
    pg.lnz.push_back('F');
    pg.nodes_id_pos.push_back(0);
    size_t n = pg.lnz.size();
    pg.nwp.assign(n, 0);
    pg.pred_hash.assign(n, {});
    size_t P = graph.paths.size();
    pg.paths_number = P;
    pg.alphas.assign(n, P + 1);
    pg.paths_nodes.assign(n, BitVec(P, 0));
    pg.paths_nodes[0] = BitVec(P, 1);
    pg.alphas[0] = 0;
    pg.alphas[n - 1] = 0;
    for (size_t path_id = 0; path_id < P; path_id++) {
        std::vector<Handle> path_nodes = graph.paths[path_id].nodes;
        if (is_reversed) std::reverse(path_nodes.begin(), path_nodes.end());
        for (size_t pos = 0; pos < path_nodes.size(); pos++) {
            auto it = handles_id_position.find(path_nodes[pos].id());
            if (it == handles_id_position.end()) throw RefPanic("path step on unknown segment");
            size_t handle_start = (size_t)it->second.first, handle_end = (size_t)it->second.second;
            for (size_t idx = handle_start; idx <= handle_end; idx++) {
                pg.paths_nodes[idx][path_id] = 1;
                if (pg.alphas[idx] == P + 1) pg.alphas[idx] = path_id;
            }
            pg.nwp[handle_start] = 1;
            if (pos == 0) {
                set_preds_and_paths(pg, handle_start, 0, path_id);
            } else {
                size_t pred_end = (size_t)handles_id_position.at(path_nodes[pos - 1].id()).second;
                set_preds_and_paths(pg, handle_start, pred_end, path_id);
                if (pos == path_nodes.size() - 1) set_preds_and_paths(pg, n - 1, handle_end, path_id);
            }
        }
    }
    pg.nwp[n - 1] = 1;
    pg.paths_nodes[n - 1] = BitVec(P, 1);
    return pg;
}
// pathwise_graph.rs:250-282
PathGraph create_reverse_path_graph(const PathGraph& fwd) {
    PathGraph r;
    r.lnz = fwd.lnz;
    r.paths_nodes = fwd.paths_nodes;
    r.alphas = fwd.alphas;
    r.paths_number = fwd.paths_number;
    r.nodes_id_pos = fwd.nodes_id_pos;
    r.nwp.assign(fwd.lnz.size(), 0);
    r.pred_hash.assign(fwd.lnz.size(), {});
    for (size_t node = 0; node < fwd.pred_hash.size(); node++)
        for (auto& pp : fwd.pred_hash[node]) {
            r.nwp[pp.first] = 1;
            for (size_t path_id = 0; path_id < pp.second.size(); path_id++)
                if (pp.second[path_id]) set_preds_and_paths(r, pp.first, node, path_id);
        }
    return r;
}
// pathwise_graph.rs:306-329 (called with the reverse graph)
std::vector<long> get_distance_from_start(const PathGraph& graph) {
    size_t n = graph.lnz.size();
    std::vector<long> r(n, -1);
    r[0] = 0;
    for (auto& pp : graph.get_preds_and_paths(0)) r[pp.first] = 1;
    for (size_t i = 1; i + 1 < n; i++) {
        if (r[i] == -1 || r[i] > r[i - 1] + 1) r[i] = r[i - 1] + 1;
        if (graph.nwp[i])
            for (auto& pp : graph.get_preds_and_paths(i))
                if (r[pp.first] == -1 || r[pp.first] > r[i] + 1) r[pp.first] = r[i] + 1;
    }
    return r;
}
// pathwise_graph.rs:330-354 (called with the forward graph)
std::vector<long> get_distance_from_end(const PathGraph& graph) {
    size_t n = graph.lnz.size();
    std::vector<long> r(n, -1);
    r[n - 1] = 0;
    for (auto& pp : graph.get_preds_and_paths(n - 1)) r[pp.first] = 1;
    for (size_t i = n - 2; i >= 1; i--) {
        if (r[i] == -1 || r[i] > r[i + 1] + 1) r[i] = r[i + 1] + 1;
        if (graph.nwp[i])
            for (auto& pp : graph.get_preds_and_paths(i))
                if (r[pp.first] == -1 || r[pp.first] > r[i] + 1) r[pp.first] = r[i] + 1;
    }
    return r;
}
// pathwise_graph.rs:284-305
Displacement nodes_displacement_matrix(const PathGraph& g, const PathGraph& rev) {
    Displacement d;
    d.dfe = get_distance_from_end(g);
    d.dfs = get_distance_from_start(rev);
    return d;
}

// ------------------------------------------------------------------------------- sequences.rs
static char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
// sequences.rs:5-45
void get_sequences_text(const std::string& text, std::vector<std::vector<char>>& sequences,
                        std::vector<std::string>& names) {
    std::istringstream in(text);
    std::string line;
    std::vector<char> sequence;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines strips "\r\n"
        if (!line.empty() && line[0] != '>') {
            for (char c : line) sequence.push_back(c == '-' ? 'N' : up(c));
        } else if (!line.empty() && line[0] == '>') {
            names.push_back(line.substr(1));
            if (!sequence.empty()) {
                sequence.insert(sequence.begin(), '$');
                sequences.push_back(sequence);
            }
            sequence.clear();
        }
    }
    if (!sequence.empty()) {
        sequence.insert(sequence.begin(), '$');
        sequences.push_back(sequence);
    }
    if (sequences.size() != names.size()) throw RefPanic("wrong fasta file format");
}
// sequences.rs:48-61
std::vector<char> build_align_string(const std::string& line) {
    std::vector<char> seq{'$'};
    for (char c : line) seq.push_back(c == '-' ? 'N' : up(c));
    return seq;
}
// sequences.rs:65-82
std::vector<char> rev_and_compl(const std::vector<char>& seq) {
    std::vector<char> r;
    for (size_t k = 1; k < seq.size(); k++) {
        switch (seq[k]) {
            case 'A': r.push_back('T'); break;
            case 'C': r.push_back('G'); break;
            case 'G': r.push_back('C'); break;
            case 'T': r.push_back('A'); break;
            case 'N': r.push_back('N'); break;
            default: throw RefPanic(std::string("wrong char: ") + seq[k] + ", unable to rev&compl");
        }
    }
    std::reverse(r.begin(), r.end());
    r.insert(r.begin(), '$');
    return r;
}

// ------------------------------------------------------------------------------- utils.rs (band)
// utils.rs:74-98
static std::pair<size_t, size_t> set_left_right_x64(size_t left, size_t right, size_t seq_len) {
    size_t new_right = right, new_left = left;
    while ((new_right - new_left) % 8 != 0) {
        if ((new_right - new_left) % 2 == 0 && new_right < seq_len)
            new_right += 1;
        else if (new_left > 0)
            new_left -= 1;
        else
            break;
    }
    if (new_left == 0)
        while ((new_right - 1) % 8 != 0 && new_right < seq_len) new_right += 1;
    if (new_right == seq_len)
        while ((new_right - new_left) % 8 != 0 && new_left > 1) new_left -= 1;
    return {new_left, new_right};
}
// utils.rs:17-72
std::pair<size_t, size_t> set_ampl_for_row(size_t i, const std::vector<size_t>& p_arr, size_t r_val,
                                           const std::vector<size_t>& best_scoring_pos, size_t seq_len,
                                           size_t bta, bool simd_version) {
    size_t ms, me;
    if (i == 0) {
        ms = 0;
        me = 0;
    } else if (p_arr.empty()) {
        size_t pl = best_scoring_pos[i - 1];
        ms = pl + 1;
        me = pl + 1;
    } else {
        size_t pl = 0, pr = 0;
        bool first = true;
        for (size_t p : p_arr) {
            size_t current_best = best_scoring_pos[p];
            if (first) {
                pl = current_best;
                pr = current_best;
                first = false;
            }
            if (current_best < pl) pl = current_best;
            if (current_best > pr) pr = current_best;
        }
        ms = pl + 1;
        me = pr + 1;
    }
    int32_t tmp_bs = std::min((int32_t)ms, ((int32_t)seq_len - (int32_t)r_val) - (int32_t)bta);
    size_t band_start = tmp_bs < 0 ? 0 : (size_t)tmp_bs;
    size_t band_end = seq_len > r_val ? std::min(seq_len, std::max(me, seq_len - r_val) + bta)
                                      : std::min(seq_len, me + bta);
    if (simd_version) return set_left_right_x64(band_start, band_end, seq_len);
    return {band_start, band_end};
}
// utils.rs:103-126
std::vector<size_t> set_r_values(const LnzGraph& g) {
    size_t lnz_len = g.lnz.size();
    std::vector<long> r(lnz_len, -1);
    r[lnz_len - 1] = 0;
    for (size_t p : g.preds(lnz_len - 1)) r[p] = 0;
    for (size_t i = lnz_len - 2; i >= 1; i--) {
        if (r[i] == -1 || r[i] > r[i + 1] + 1) r[i] = r[i + 1] + 1;
        if (g.nwp[i])
            for (size_t p : g.preds(i))
                if (r[p] == -1 || r[p] > r[i] + 1) r[p] = r[i] + 1;
    }
    std::vector<size_t> out(lnz_len);
    for (size_t i = 0; i < lnz_len; i++) out[i] = (size_t)r[i];  // `*x as usize`: -1 wraps to usize::MAX
    return out;
}
// main.rs:57,175: `(b + f * seq.len() as f32) as usize` — f32 arithmetic, saturating cast.
size_t bases_to_add(float b, float f, size_t seq_len) {
    volatile float prod = f * (float)seq_len;
    volatile float v = b + prod;
    float vv = v;
    if (!(vv > 0.0f)) return 0;
    if (vv >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)vv;
}

}  // namespace rgo
