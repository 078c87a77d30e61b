// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). Pathwise (modes 4/5) and recombination (modes 8/9):
// pathwise_alignment.rs, pathwise_alignment_semiglobal.rs, pathwise_alignment_recombination.rs,
// pathwise_alignment_output.rs, recombination_output.rs, utils.rs:221-323.
//
// The DP is restated literally in the reference's delta encoding (scores of non-leader paths are stored
// relative to the node's `alpha` path). The forward (align / exec) and reverse (rev_align) passes of the
// reference are the same cell code mirrored in i and j, so one routine parameterised by direction covers
// pathwise_alignment.rs:18-304, pathwise_alignment_semiglobal.rs:19-225 and
// pathwise_alignment_recombination.rs:146-431,450-741.
#include <algorithm>

#include "oracle.hpp"

namespace rgo {

struct Dpm {
    size_t n, L, P;
    std::vector<int> v;
    Dpm(size_t n_, size_t L_, size_t P_) : n(n_), L(L_), P(P_), v(n_ * L_ * P_, 0) {}
    inline int& at(size_t i, size_t j, size_t p) {
        if (p >= P) throw RefPanic("index out of bounds: path index");
        return v[(i * L + j) * P + p];
    }
    inline int at(size_t i, size_t j, size_t p) const {
        if (p >= P) throw RefPanic("index out of bounds: path index");
        return v[(i * L + j) * P + p];
    }
};
static inline bool bit(const BitVec& b, size_t k) {
    if (k >= b.size()) throw RefPanic("BitVec index out of bounds");
    return b[k];
}
static BitVec band(const BitVec& a, const BitVec& b) {
    BitVec r(a.size());
    for (size_t k = 0; k < a.size(); k++) r[k] = a[k] & b[k];
    return r;
}
static inline int max3(int d, int u, int l) { return std::max(d, std::max(u, l)); }

// One delta-encoded DP pass. rev=false: rows 0..n-2 ascending, columns 0..L-1 ascending, neighbours i-1 / j-1.
// rev=true : rows n-1..1 descending, columns L-1..1 descending, neighbours i+1 / j+1 (rev graph preds).
// free_border: border column (0 fwd / L-1 rev) is all zeros (modes 5 / 9) instead of a per-path gap chain.
static Dpm pathwise_dp(const std::vector<char>& sequence, const PathGraph& graph, const ScoreMatrix& sm, bool rev,
                       bool free_border) {
    const auto& lnz = graph.lnz;
    const auto& nwp = graph.nwp;
    const auto& path_node = graph.paths_nodes;
    const auto& alphas = graph.alphas;
    const size_t P = graph.paths_number, n = lnz.size(), L = sequence.size();
    Dpm dpm(n, L, P);
    const size_t base_row = rev ? n - 1 : 0;
    const size_t border_col = rev ? L - 1 : 0;

    auto cell_border = [&](size_t i, size_t j) {  // (_, 0) fwd  /  j == last_char_pos rev
        const size_t ip = rev ? i + 1 : i - 1;
        if (!nwp[i]) {
            BitVec common = band(path_node[i], path_node[ip]);
            if (bit(common, alphas[ip])) {
                for (size_t path = 0; path < P; path++)
                    if (common[path]) {
                        if (path == alphas[i])
                            dpm.at(i, j, path) = dpm.at(ip, j, path) + sm.get(lnz[i], '-');
                        else
                            dpm.at(i, j, path) = dpm.at(ip, j, path);
                    }
            } else {
                dpm.at(i, j, alphas[i]) = dpm.at(ip, j, alphas[i]) + dpm.at(ip, j, alphas[ip]) + sm.get(lnz[i], '-');
                for (size_t path = 0; path < P; path++)
                    if (common[path] && path != alphas[i])
                        dpm.at(i, j, path) = dpm.at(ip, j, path) - dpm.at(ip, j, alphas[i]);
            }
        } else {
            std::vector<std::pair<size_t, std::vector<size_t>>> alphas_deltas;
            for (auto& pp : graph.get_preds_and_paths(i)) {
                size_t p = pp.first;
                BitVec common = band(path_node[i], pp.second);
                std::vector<size_t> paths;
                for (size_t k = 0; k < P; k++)
                    if (common[k]) paths.push_back(k);
                if (bit(common, alphas[p])) {
                    alphas_deltas.push_back({alphas[p], paths});
                    dpm.at(i, j, alphas[p]) = dpm.at(p, j, alphas[p]) + sm.get(lnz[i], '-');
                    for (size_t path : paths)
                        if (path != alphas[p]) dpm.at(i, j, path) = dpm.at(p, j, path);
                } else {
                    size_t temp_alpha;
                    if (bit(common, alphas[i]))
                        temp_alpha = alphas[i];
                    else {
                        if (paths.empty()) throw RefPanic("position(|is_in| is_in).unwrap() on None");
                        temp_alpha = paths[0];
                    }
                    alphas_deltas.push_back({temp_alpha, paths});
                    dpm.at(i, j, temp_alpha) = dpm.at(p, j, alphas[p]) + dpm.at(p, j, temp_alpha) + sm.get(lnz[i], '-');
                    for (size_t path : paths)
                        if (path != temp_alpha) dpm.at(i, j, path) = dpm.at(p, j, path) - dpm.at(p, j, temp_alpha);
                }
            }
            for (auto& ad : alphas_deltas) {
                size_t a = ad.first;
                if (a != alphas[i]) {
                    dpm.at(i, j, a) -= dpm.at(i, j, alphas[i]);
                    for (size_t path : ad.second)
                        if (path != a) dpm.at(i, j, path) += dpm.at(i, j, a);
                }
            }
        }
    };

    auto cell_general = [&](size_t i, size_t j) {
        const size_t ip = rev ? i + 1 : i - 1;
        const size_t jp = rev ? j + 1 : j - 1;
        if (!nwp[i]) {
            BitVec common = band(path_node[i], path_node[ip]);
            if (bit(common, alphas[ip])) {
                int u = dpm.at(ip, j, alphas[ip]) + sm.get(lnz[i], '-');
                int d = dpm.at(ip, jp, alphas[ip]) + sm.get(lnz[i], sequence[j]);
                int l = dpm.at(i, jp, alphas[i]) + sm.get(sequence[j], '-');
                int best = max3(d, u, l);
                dpm.at(i, j, alphas[i]) = best;
                for (size_t path = 0; path < P; path++)
                    if (common[path] && path != alphas[i]) {
                        if (best == d)
                            dpm.at(i, j, path) = dpm.at(ip, jp, path);
                        else if (best == u)
                            dpm.at(i, j, path) = dpm.at(ip, j, path);
                        else
                            dpm.at(i, j, path) = dpm.at(i, jp, path);
                    }
            } else {
                int u = dpm.at(ip, j, alphas[ip]) + dpm.at(ip, j, alphas[i]) + sm.get(lnz[i], '-');
                int d = dpm.at(ip, jp, alphas[ip]) + dpm.at(ip, jp, alphas[i]) + sm.get(lnz[i], sequence[j]);
                int l = dpm.at(i, jp, alphas[i]) + sm.get(sequence[j], '-');
                int best = max3(d, u, l);
                dpm.at(i, j, alphas[i]) = best;
                for (size_t path = 0; path < P; path++)
                    if (common[path] && path != alphas[i]) {
                        if (best == d)
                            dpm.at(i, j, path) = dpm.at(ip, jp, path) - dpm.at(ip, jp, alphas[i]);
                        else if (best == u)
                            dpm.at(i, j, path) = dpm.at(ip, j, path) - dpm.at(ip, j, alphas[i]);
                        else
                            dpm.at(i, j, path) = dpm.at(i, jp, path);
                    }
            }
        } else {
            std::vector<std::pair<size_t, std::vector<size_t>>> alphas_deltas;
            for (auto& pp : graph.get_preds_and_paths(i)) {
                size_t p = pp.first;
                BitVec common = band(path_node[i], pp.second);
                std::vector<size_t> paths;
                for (size_t k = 0; k < P; k++)
                    if (common[k]) paths.push_back(k);
                if (bit(common, alphas[p])) {
                    size_t ap = alphas[p];
                    alphas_deltas.push_back({ap, paths});
                    int u = dpm.at(p, j, ap) + sm.get(lnz[i], '-');
                    int d = dpm.at(p, jp, ap) + sm.get(lnz[i], sequence[j]);
                    int l = alphas[i] == ap ? dpm.at(i, jp, ap) + sm.get(sequence[j], '-')
                                            : dpm.at(i, jp, ap) + dpm.at(i, jp, alphas[i]) + sm.get(sequence[j], '-');
                    int best = max3(d, u, l);
                    dpm.at(i, j, ap) = best;
                    for (size_t path : paths)
                        if (path != ap) {
                            if (best == d)
                                dpm.at(i, j, path) = dpm.at(p, jp, path);
                            else if (best == u)
                                dpm.at(i, j, path) = dpm.at(p, j, path);
                            else if (ap == alphas[i])
                                dpm.at(i, j, path) = dpm.at(i, jp, path);
                            else
                                dpm.at(i, j, path) = dpm.at(i, jp, path) - dpm.at(i, jp, ap);
                        }
                } else {
                    size_t temp_alpha;
                    if (bit(common, alphas[i]))
                        temp_alpha = alphas[i];
                    else {
                        if (paths.empty()) throw RefPanic("position(|is_in| is_in).unwrap() on None");
                        temp_alpha = paths[0];
                    }
                    alphas_deltas.push_back({temp_alpha, paths});
                    int u = dpm.at(p, j, alphas[p]) + dpm.at(p, j, temp_alpha) + sm.get(lnz[i], '-');
                    int d = dpm.at(p, jp, alphas[p]) + dpm.at(p, jp, temp_alpha) + sm.get(lnz[i], sequence[j]);
                    int l = alphas[i] == temp_alpha
                                ? dpm.at(i, jp, temp_alpha) + sm.get(sequence[j], '-')
                                : dpm.at(i, jp, temp_alpha) + dpm.at(i, jp, alphas[i]) + sm.get(sequence[j], '-');
                    int best = max3(d, u, l);
                    dpm.at(i, j, temp_alpha) = best;
                    for (size_t path : paths)
                        if (path != temp_alpha) {
                            if (best == d)
                                dpm.at(i, j, path) = dpm.at(p, jp, path) - dpm.at(p, jp, temp_alpha);
                            else if (best == u)
                                dpm.at(i, j, path) = dpm.at(p, j, path) - dpm.at(p, j, temp_alpha);
                            else if (temp_alpha == alphas[i])
                                dpm.at(i, j, path) = dpm.at(i, jp, path);
                            else
                                dpm.at(i, j, path) = dpm.at(i, jp, path) - dpm.at(i, jp, temp_alpha);
                        }
                }
            }
            for (auto& ad : alphas_deltas) {
                size_t a = ad.first;
                if (a != alphas[i]) {
                    dpm.at(i, j, a) -= dpm.at(i, j, alphas[i]);
                    for (size_t path : ad.second)
                        if (path != a) dpm.at(i, j, path) += dpm.at(i, j, a);
                }
            }
        }
    };

    auto cell = [&](size_t i, size_t j) {
        if (i == base_row && j == border_col) {
            // zeros
        } else if (i == base_row) {
            const size_t jp = rev ? j + 1 : j - 1;
            size_t a = alphas[base_row];
            dpm.at(i, j, a) = dpm.at(i, jp, a) + sm.get(sequence[j], '-');
            for (size_t k = a + 1; k < P; k++) dpm.at(i, j, k) = dpm.at(i, jp, k);
        } else if (j == border_col) {
            if (!free_border) cell_border(i, j);
        } else {
            cell_general(i, j);
        }
    };
    if (!rev) {
        for (size_t i = 0; i + 1 < n; i++)
            for (size_t j = 0; j < L; j++) cell(i, j);
    } else {
        // rev_align: `if i == last && j == last {0} else if i == last {...} else if j == last {...}`
        for (size_t i = n - 1; i >= 1; i--)
            for (size_t j = L - 1; j >= 1; j--) cell(i, j);
    }
    return dpm;
}

// pathwise_alignment_recombination.rs:747-757
static void absolute_scores(Dpm& dpm, const std::vector<size_t>& alphas, const std::vector<BitVec>& paths_nodes) {
    for (size_t i = 0; i + 1 < dpm.n; i++)
        for (size_t j = 0; j < dpm.L; j++)
            for (size_t path = 0; path < dpm.P; path++)
                if (path != alphas[i] && paths_nodes[i][path]) dpm.at(i, j, path) += dpm.at(i, j, alphas[i]);
}

static std::vector<size_t> dedup_u64(const std::vector<uint64_t>& v) {
    std::vector<size_t> out;
    for (uint64_t x : v)
        if (out.empty() || out.back() != (size_t)x) out.push_back((size_t)x);
    return out;
}

// utils.rs:221-254
static void get_path_len_start_end(const std::vector<uint64_t>& ids, size_t start, size_t end, size_t path_len,
                                   size_t& o_len, size_t& o_start, size_t& o_end) {
    size_t path_start = 0;
    if (start > 0) {
        uint64_t first_node_id = ids.at(start);
        size_t counter = start - 1;
        while (counter > 0 && ids[counter] == first_node_id) {
            counter -= 1;
            path_start += 1;
        }
    }
    size_t path_end = path_len > 0 ? path_start + path_len - 1 : 0;
    size_t end_offset = 0;
    if (end > 0) {
        uint64_t last_node_id = ids.at(end);
        size_t counter = end + 1;
        while (counter < ids.size() - 1 && ids[counter] == last_node_id) {
            counter += 1;
            end_offset += 1;
        }
    }
    o_len = path_end + end_offset + 1;
    o_start = path_start;
    o_end = path_end;
}
// utils.rs:256-323
static void get_rec_path_len_start_end(const std::vector<uint64_t>& ids, size_t fen, size_t rsn, size_t start,
                                       size_t end, size_t forw_path_length, size_t rev_path_length, size_t& o_len,
                                       size_t& o_start, size_t& o_end) {
    size_t path_start = 0;
    if (start > 0) {
        uint64_t first_node_id = ids.at(start);
        size_t counter = start - 1;
        while (counter > 0 && ids[counter] == first_node_id) {
            counter -= 1;
            path_start += 1;
        }
    }
    size_t forw_path_end = forw_path_length > 0 ? path_start + forw_path_length - 1 : 0;
    size_t forw_end_offset = 0;
    if (fen > 0) {
        uint64_t last_node_id = ids.at(fen);
        size_t counter = fen + 1;
        while (counter < ids.size() - 1 && ids[counter] == last_node_id) {
            counter += 1;
            forw_end_offset += 1;
        }
    }
    size_t forw_path_len = forw_path_end + forw_end_offset + 1;
    size_t rev_path_start = 0;
    if (rsn > 0) {
        uint64_t first_node_id = ids.at(rsn);
        size_t counter = rsn - 1;
        while (counter > 0 && ids[counter] == first_node_id) {
            counter -= 1;
            rev_path_start += 1;
        }
    }
    size_t rev_path_end = rev_path_length > 0 ? rev_path_start + rev_path_length - 1 : 0;
    size_t path_end = forw_path_len + rev_path_end;
    size_t end_offset = 0;
    if (end > 0) {
        uint64_t last_node_id = ids.at(end);
        size_t counter = end + 1;
        while (counter < ids.size() - 1 && ids[counter] == last_node_id) {
            counter += 1;
            end_offset += 1;
        }
    }
    size_t rev_path_len = rev_path_end + end_offset + 1;
    o_len = forw_path_len + rev_path_len;
    o_start = path_start;
    o_end = path_end;
}

// pathwise_alignment_output.rs:7-184. `absolute`: scores already absolute (modes 8/9 no_rec builders,
// recombination_output.rs:239-361,633-782 — same walk without the alpha corrections).
static GAFStruct build_alignment(const Dpm& dpm, const PathGraph& g, const std::vector<char>& seq,
                                 const ScoreMatrix& sm, size_t best_path, size_t ending_node, bool global_align,
                                 bool absolute) {
    const auto& lnz = g.lnz;
    const auto& alphas = g.alphas;
    const auto& nwp = g.nwp;
    const auto& ids = g.nodes_id_pos;
    auto absval = [&](size_t i, size_t j) {
        if (absolute || alphas[i] == best_path) return dpm.at(i, j, best_path);
        return dpm.at(i, j, best_path) + dpm.at(i, j, alphas[i]);
    };
    std::vector<char> cigar, path_sequence;
    std::vector<uint64_t> handle_id_alignment;
    size_t path_length = 0;
    size_t i = ending_node, j = dpm.L - 1;
    int score = absval(i, j);
    while (i > 0 && j > 0) {
        bool has_pred = false;
        size_t predecessor = 0;
        int d = 0, u = 0, l = 0;
        if (!nwp[i]) {
            d = absval(i - 1, j - 1) + sm.get(lnz[i], seq[j]);
            u = absval(i - 1, j) + sm.get(lnz[i], '-');
            l = absval(i, j - 1) + sm.get('-', seq[j]);
        } else {
            for (auto& pp : g.get_preds_and_paths(i))
                if (bit(pp.second, best_path)) {
                    has_pred = true;
                    predecessor = pp.first;
                    d = absval(pp.first, j - 1) + sm.get(lnz[i], seq[j]);
                    u = absval(pp.first, j) + sm.get(lnz[i], '-');
                    l = absval(i, j - 1) + sm.get('-', seq[j]);
                }
        }
        int mx = max3(d, u, l);
        if (mx == d) {
            cigar.push_back(lnz[i] != seq[j] ? 'd' : 'D');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i - 1;
            j -= 1;
            path_length += 1;
        } else if (mx == u) {
            cigar.push_back('U');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i - 1;
            path_length += 1;
        } else {
            cigar.push_back('L');
            j -= 1;
        }
    }
    while (j > 0) {
        cigar.push_back('L');
        j -= 1;
    }
    if (global_align) {
        while (i > 0) {
            cigar.push_back('U');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            path_length += 1;
            size_t predecessor;
            if (!nwp[i]) {
                predecessor = i - 1;
            } else if (absolute) {
                // recombination_output.rs:721-734: None => i - 1
                bool hp = false;
                size_t p = 0;
                for (auto& pp : g.get_preds_and_paths(i))
                    if (bit(pp.second, best_path)) {
                        hp = true;
                        p = pp.first;
                    }
                predecessor = hp ? p : i - 1;
            } else {
                // pathwise_alignment_output.rs:123-134: default 0
                size_t p = 0;
                for (auto& pp : g.get_preds_and_paths(i))
                    if (bit(pp.second, best_path)) p = pp.first;
                predecessor = p;
            }
            i = predecessor;
        }
    }
    std::reverse(cigar.begin(), cigar.end());
    std::reverse(path_sequence.begin(), path_sequence.end());
    GAFStruct gaf;
    gaf.query_name = "Temp";
    gaf.query_length = dpm.L - 1;
    gaf.query_start = 0;
    gaf.query_end = dpm.L - 2;
    gaf.strand = '+';
    gaf.path = dedup_u64(handle_id_alignment);
    std::reverse(gaf.path.begin(), gaf.path.end());
    get_path_len_start_end(ids, i == 0 ? i : i + 1, ending_node, path_length, gaf.path_length, gaf.path_start,
                           gaf.path_end);
    gaf.residue_matches_number = 0;
    gaf.alignment_block_length = "*";
    gaf.mapping_quality = "*";
    gaf.comments = build_cigar(cigar) + ", best path: " + std::to_string(best_path) + ", score: " +
                   std::to_string(score) + "\t" + std::string(path_sequence.begin(), path_sequence.end());
    return gaf;
}

// pathwise_alignment.rs:5-340
GAFStruct pathwise_alignment_exec(const std::vector<char>& sequence, const PathGraph& g, const ScoreMatrix& sm) {
    Dpm dpm = pathwise_dp(sequence, g, sm, false, false);
    const size_t P = g.paths_number, n = g.lnz.size();
    std::vector<size_t> ending_nodes(P, 0);
    std::vector<int> results(P, 0);
    for (auto& pp : g.get_preds_and_paths(n - 1)) {
        size_t pred = pp.first;
        for (size_t path = 0; path < P; path++)
            if (pp.second[path]) {
                if (path == g.alphas[pred])
                    results[path] = dpm.at(pred, dpm.L - 1, path);
                else
                    results[path] = dpm.at(pred, dpm.L - 1, path) + dpm.at(pred, dpm.L - 1, g.alphas[pred]);
                ending_nodes[path] = pred;
            }
    }
    if (P == 0) throw RefPanic("best_path.unwrap() on None");
    size_t best_path = 0;  // max of (score, path): last maximum, i.e. highest path id on ties
    for (size_t path = 1; path < P; path++)
        if (results[path] >= results[best_path]) best_path = path;
    return build_alignment(dpm, g, sequence, sm, best_path, ending_nodes[best_path], true, false);
}

// pathwise_alignment_semiglobal.rs:244-277
static std::pair<size_t, size_t> best_ending_node(const Dpm& dpm, const PathGraph& g) {
    bool has_max = false;
    int mx = 0;
    size_t ending_node = 0, chosen_path = 0;
    for (size_t i = 1; i + 1 < dpm.n; i++) {
        const BitVec& paths = g.paths_nodes[i];
        bool has_best = false;
        int best = 0;
        size_t bp = 0;
        for (size_t path = 0; path < dpm.P; path++) {
            if (!paths[path]) continue;
            int score = dpm.at(i, dpm.L - 1, path);
            if (path != g.alphas[i]) score += dpm.at(i, dpm.L - 1, g.alphas[i]);
            if (!has_best || best < score) {
                has_best = true;
                best = score;
                bp = path;
            }
        }
        if (!has_best) throw RefPanic("best_path.unwrap() on None (node on no path)");
        if (!has_max || best > mx) {
            has_max = true;
            mx = best;
            ending_node = i;
            chosen_path = bp;
        }
    }
    return {ending_node, chosen_path};
}
// pathwise_alignment_semiglobal.rs:6-242
GAFStruct pathwise_alignment_semiglobal_exec(const std::vector<char>& sequence, const PathGraph& g,
                                             const ScoreMatrix& sm) {
    Dpm dpm = pathwise_dp(sequence, g, sm, false, true);
    auto fb = best_ending_node(dpm, g);
    return build_alignment(dpm, g, sequence, sm, fb.second, fb.first, false, false);
}

// pathwise_alignment_recombination.rs:875-883
static std::vector<char> get_rev_sequence(const std::vector<char>& seq) {
    std::vector<char> r(seq.begin() + 1, seq.end());
    r.push_back('F');
    return r;
}
// pathwise_alignment_recombination.rs:9-22
static int get_node_offset(const std::vector<uint64_t>& ids, size_t curr_node) {
    uint64_t handle = ids[curr_node];
    if (handle == 0) return 0;
    size_t counter = curr_node;
    int offset = 0;
    while (ids[counter - 1] == handle) {
        counter -= 1;
        offset += 1;
    }
    return offset;
}

struct BestAln {
    size_t fen = 0, rsn = 0, fp = 0, rp = 0, col = 0;
    float score = 0;
    int displ = 0;
};
// pathwise_alignment_recombination.rs:759-873
static BestAln best_alignment(const Dpm& m, const Dpm& w, const Displacement& dms, int brc, float mrc, int aln_mode,
                              const PathGraph& g, float rbw) {
    const auto& nodes_path = g.paths_nodes;
    const auto& ids = g.nodes_id_pos;
    const size_t n = m.n, L = m.L, P = m.P;
    BestAln b;
    bool has_max = false;
    int mx = 0;
    size_t best_path = 0;
    if (aln_mode == 8) {
        for (auto& pp : g.get_preds_and_paths(n - 1))
            for (size_t path = 0; path < P; path++)
                if (pp.second[path]) {
                    int v = m.at(pp.first, L - 1, path);
                    if (!has_max || mx < v) {
                        has_max = true;
                        mx = v;
                        best_path = path;
                    }
                }
    } else {
        for (size_t i = 0; i + 1 < n; i++)
            for (size_t path = 0; path < P; path++)
                if (nodes_path[i][path]) {
                    int v = m.at(i, L - 1, path);
                    if (!has_max || mx < v) {
                        has_max = true;
                        mx = v;
                        best_path = path;
                    }
                }
    }
    if (!has_max) throw RefPanic("max.unwrap() on None");
    float curr_best_score = (float)mx;
    b.fp = best_path;
    b.rp = best_path;
    bool onedge = false;
    volatile float t1 = (float)L * (1.0f - rbw);
    volatile float t2 = t1 / 2.0f;
    float t2v = t2;
    int oob_i = t2v != t2v ? 0 : (t2v >= 2147483648.0f ? INT32_MAX : (t2v <= -2147483648.0f ? INT32_MIN : (int)t2v));
    int out_of_band = std::max(oob_i, 1);
    int rec_penalty = 0;
    std::vector<size_t> forw_paths(n), rev_paths(n);
    if ((size_t)out_of_band > L) throw RefPanic("attempt to subtract with overflow");
    for (size_t j = (size_t)out_of_band; j < L - (size_t)out_of_band; j++) {
        for (size_t i = 0; i < n; i++) {
            // max of (score, path) over ALL P slots (non-members hold 0): last maximum wins
            size_t bf = 0, br = 0;
            for (size_t path = 1; path < P; path++) {
                if (m.at(i, j, path) >= m.at(i, j, bf)) bf = path;
                if (w.at(i, j, path) >= w.at(i, j, br)) br = path;
            }
            forw_paths[i] = bf;
            rev_paths[i] = br;
        }
        for (size_t i = 1; i + 1 < n; i++) {
            size_t forw_path = forw_paths[i];
            if (!nodes_path[i][forw_path]) continue;
            for (size_t rev_i = 1; rev_i + 1 < n; rev_i++) {
                if (ids[i] != ids[rev_i]) {
                    size_t rev_path = rev_paths[rev_i];
                    if (forw_path != rev_path && nodes_path[rev_i][rev_path]) {
                        int dd = dms.at(i, rev_i);
                        volatile float mul = mrc * (float)dd;
                        volatile float penalty = (float)brc + mul;
                        volatile float new_score_v = (float)(m.at(i, j, forw_path) + w.at(rev_i, j, rev_path)) - penalty;
                        float new_score = new_score_v;
                        bool edge = (i + 1 == n || ids[i] != ids[i + 1]) && ids[rev_i] != ids[rev_i - 1];
                        if (new_score > curr_best_score || (new_score == curr_best_score && !onedge && edge)) {
                            onedge = edge;
                            curr_best_score = new_score;
                            b.fen = i;
                            b.rsn = rev_i;
                            b.fp = forw_path;
                            b.rp = rev_path;
                            b.col = j;
                            rec_penalty = dd;
                        }
                    }
                }
            }
        }
    }
    b.score = curr_best_score;
    b.displ = rec_penalty;
    return b;
}

// pathwise_alignment_recombination.rs:885-897
static size_t ending_node_of(const Dpm& dpm, size_t best_path, const std::vector<BitVec>& paths_nodes) {
    bool has = false;
    int best_score = 0;
    size_t best_node = 0;
    for (size_t i = 1; i + 1 < dpm.n; i++)
        if (paths_nodes[i][best_path]) {
            int v = dpm.at(i, dpm.L - 1, best_path);
            if (!has || v > best_score) {
                has = true;
                best_score = v;
                best_node = i;
            }
        }
    return best_node;
}

// recombination_output.rs:12-237 (semiglobal) and 363-631 (global: additionally pads U to both graph ends)
static GAFStruct gaf_output_rec(const Dpm& dpm, const Dpm& rev_dpm, const PathGraph& g, const PathGraph& rg,
                                const std::vector<char>& seq, const ScoreMatrix& sm, const BestAln& b,
                                bool global) {
    const auto& lnz = g.lnz;
    const auto& ids = g.nodes_id_pos;
    const size_t n = dpm.n, L = dpm.L;
    const size_t best_path = b.fp, rev_best_path = b.rp;
    std::vector<char> cigar, path_sequence;
    std::vector<uint64_t> handle_id_alignment;
    size_t rev_path_length = 0;
    size_t i = b.rsn, j = b.col;
    size_t rev_ending_node = i;
    std::vector<char> r_seq = get_rev_sequence(seq);
    while (i > 0 && i < n - 1 && j < L - 1) {
        bool has_pred = false;
        size_t predecessor = 0;
        int d = 0, u = 0, l = 0;
        if (!rg.nwp[i]) {
            d = rev_dpm.at(i + 1, j + 1, rev_best_path) + sm.get(lnz[i], r_seq[j]);
            u = rev_dpm.at(i + 1, j, rev_best_path) + sm.get(lnz[i], '-');
            l = rev_dpm.at(i, j + 1, rev_best_path) + sm.get('-', r_seq[j]);
        } else {
            for (auto& pp : rg.get_preds_and_paths(i))
                if (bit(pp.second, rev_best_path)) {
                    has_pred = true;
                    predecessor = pp.first;
                    d = rev_dpm.at(pp.first, j + 1, rev_best_path) + sm.get(lnz[i], r_seq[j]);
                    u = rev_dpm.at(pp.first, j, rev_best_path) + sm.get(lnz[i], '-');
                    l = rev_dpm.at(i, j + 1, rev_best_path) + sm.get('-', r_seq[j]);
                }
        }
        int mx = max3(d, u, l);
        rev_ending_node = i;
        if (mx == d) {
            cigar.push_back(lnz[i] != r_seq[j] ? 'd' : 'D');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i + 1;
            j += 1;
            rev_path_length += 1;
        } else if (mx == u) {
            cigar.push_back('U');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i + 1;
            rev_path_length += 1;
        } else {
            cigar.push_back('L');
            j += 1;
        }
    }
    while (j < L - 1) {
        cigar.push_back('L');
        j += 1;
    }
    if (global) {
        while (i < n - 1) {
            cigar.push_back('U');
            handle_id_alignment.push_back(ids[i]);
            path_sequence.push_back(lnz[i]);
            bool hp = false;
            size_t p = 0;
            if (rg.nwp[i])
                for (auto& pp : rg.get_preds_and_paths(i))
                    if (bit(pp.second, rev_best_path)) {
                        hp = true;
                        p = pp.first;
                    }
            i = hp ? p : i + 1;
            rev_path_length += 1;
        }
    }
    size_t path_length = 0;
    std::vector<char> temp_cigar, temp_path_sequence;
    std::vector<uint64_t> temp_handle_id_alignment;
    i = b.fen;
    j = b.col;
    while (i > 0 && j > 0) {
        bool has_pred = false;
        size_t predecessor = 0;
        int d = 0, u = 0, l = 0;
        if (!g.nwp[i]) {
            d = dpm.at(i - 1, j - 1, best_path) + sm.get(lnz[i], seq[j]);
            u = dpm.at(i - 1, j, best_path) + sm.get(lnz[i], '-');
            l = dpm.at(i, j - 1, best_path) + sm.get('-', seq[j]);
        } else {
            for (auto& pp : g.get_preds_and_paths(i))
                if (bit(pp.second, best_path)) {
                    has_pred = true;
                    predecessor = pp.first;
                    d = dpm.at(pp.first, j - 1, best_path) + sm.get(lnz[i], seq[j]);
                    u = dpm.at(pp.first, j, best_path) + sm.get(lnz[i], '-');
                    l = dpm.at(i, j - 1, best_path) + sm.get('-', seq[j]);
                }
        }
        int mx = max3(d, u, l);
        if (mx == d) {
            temp_cigar.push_back(lnz[i] != seq[j] ? 'd' : 'D');
            temp_handle_id_alignment.push_back(ids[i]);
            temp_path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i - 1;
            j -= 1;
            path_length += 1;
        } else if (mx == u) {
            temp_cigar.push_back('U');
            temp_handle_id_alignment.push_back(ids[i]);
            temp_path_sequence.push_back(lnz[i]);
            i = has_pred ? predecessor : i - 1;
            path_length += 1;
        } else {
            temp_cigar.push_back('L');
            j -= 1;
        }
    }
    while (j > 0) {
        temp_cigar.push_back('L');
        j -= 1;
    }
    if (global) {
        while (i > 0) {
            temp_cigar.push_back('U');
            temp_handle_id_alignment.push_back(ids[i]);
            temp_path_sequence.push_back(lnz[i]);
            bool hp = false;
            size_t p = 0;
            if (g.nwp[i])
                for (auto& pp : g.get_preds_and_paths(i))
                    if (bit(pp.second, best_path)) {
                        hp = true;
                        p = pp.first;
                    }
            i = hp ? p : i - 1;
            path_length += 1;
        }
    }
    uint64_t rec_edge = (uint64_t)temp_path_sequence.size() - 1;  // usize wrap in a release build
    std::reverse(temp_cigar.begin(), temp_cigar.end());
    temp_cigar.insert(temp_cigar.end(), cigar.begin(), cigar.end());
    std::reverse(temp_handle_id_alignment.begin(), temp_handle_id_alignment.end());
    temp_handle_id_alignment.insert(temp_handle_id_alignment.end(), handle_id_alignment.begin(),
                                    handle_id_alignment.end());
    std::reverse(temp_path_sequence.begin(), temp_path_sequence.end());
    temp_path_sequence.insert(temp_path_sequence.end(), path_sequence.begin(), path_sequence.end());
    std::string path_sequence_string(temp_path_sequence.begin(), temp_path_sequence.end());
    GAFStruct gaf;
    gaf.query_name = "Temp";
    gaf.query_length = L - 1;
    gaf.query_start = 0;
    gaf.query_end = L - 2;
    gaf.strand = '+';
    gaf.path = dedup_u64(temp_handle_id_alignment);
    size_t start = i == 0 ? i : i + 1;
    get_rec_path_len_start_end(ids, b.fen, b.rsn, start, rev_ending_node, path_length, rev_path_length,
                               gaf.path_length, gaf.path_start, gaf.path_end);
    gaf.residue_matches_number = 0;
    gaf.alignment_block_length = "*";
    gaf.mapping_quality = "*";
    std::string recombination;
    if (best_path == rev_best_path) {
        recombination = "No recombination, best path: " + std::to_string(best_path);
    } else {
        int fen_offset = get_node_offset(ids, b.fen), rsn_offset = get_node_offset(ids, b.rsn);
        recombination = "recombination path " + std::to_string(best_path) + " " + std::to_string(rev_best_path) +
                        ", nodes " + std::to_string(ids[b.fen]) + "[" + std::to_string(fen_offset) + "] " +
                        std::to_string(ids[b.rsn]) + "[" + std::to_string(rsn_offset) + "], score: " +
                        f32_display(b.score) + ", displacement: " + std::to_string(b.displ) + "\t" +
                        path_sequence_string + "\t" + std::to_string(rec_edge);
    }
    gaf.comments = build_cigar(temp_cigar) + ", " + recombination;
    return gaf;
}

// pathwise_alignment_recombination.rs:23-127
GAFStruct pathwise_alignment_recombination_exec(int aln_mode, const std::vector<char>& sequence, const PathGraph& g,
                                                const PathGraph& rev_g, const ScoreMatrix& sm, int base_rec_cost,
                                                float multi_rec_cost, const Displacement& displ, float rbw) {
    Dpm forward_matrix = pathwise_dp(sequence, g, sm, false, aln_mode == 9);
    absolute_scores(forward_matrix, g.alphas, g.paths_nodes);
    std::vector<char> rev_sequence = get_rev_sequence(sequence);
    Dpm reverse_matrix = pathwise_dp(rev_sequence, rev_g, sm, true, aln_mode == 9);
    absolute_scores(reverse_matrix, rev_g.alphas, rev_g.paths_nodes);
    BestAln b = best_alignment(forward_matrix, reverse_matrix, displ, base_rec_cost, multi_rec_cost, aln_mode, g, rbw);
    if (aln_mode == 8) {
        if (b.fp == b.rp) {
            // gaf_output_global_no_rec (recombination_output.rs:633-782)
            size_t i = 0;
            for (auto& pp : g.get_preds_and_paths(forward_matrix.n - 1))
                if (bit(pp.second, b.fp)) i = pp.first;
            return build_alignment(forward_matrix, g, sequence, sm, b.fp, i, true, true);
        }
        return gaf_output_rec(forward_matrix, reverse_matrix, g, rev_g, sequence, sm, b, true);
    }
    if (b.fp == b.rp) {
        size_t en = ending_node_of(forward_matrix, b.fp, g.paths_nodes);
        return build_alignment(forward_matrix, g, sequence, sm, b.fp, en, false, true);
    }
    return gaf_output_rec(forward_matrix, reverse_matrix, g, rev_g, sequence, sm, b, false);
}

}  // namespace rgo
