// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). Affine-gap pathwise alignment, the reference's experimental
// modes 6 (global) and 7 (semiglobal): pathwise_alignment_gap.rs:4-574, pathwise_alignment_gap_semi.rs:5-473 and the
// CIGAR builders pathwise_alignment_output.rs:186-451. These modes print a CIGAR line (mode 7 appends
// "\t(start end)") and main.rs:277,286 then prints "Best path sequence {i}: {path}"; no GAF record.
//
// Literal restatement in the reference's delta encoding: three tensors dpm / x / y, scores of non-leader paths stored
// relative to the node's alpha path, every quirk kept (the builders compare RAW delta entries of dpm with x / y in
// their chain loops; `max == d` tests the predecessor's score without the substitution score). Mode 7 is mode 6
// without the first-column case (`(_, 0) => dpm[0][0] = vec![0; P]`, pathwise_alignment_gap_semi.rs:28) and with
// best_ending_node (:447-473). Parity unpinned: the reference ships no test or golden output for these modes; the
// self-checks in tests/test_oracle_golden.py are consistency properties (o = 0 equals the linear modes 4 / 5 scores).
// HashMap `alphas_deltas` (…_gap.rs:41,228) is walked in ascending key order; the entries touch disjoint path sets, so
// the order does not change the result.
#include <algorithm>
#include <map>

#include "oracle.hpp"

namespace rgo {

namespace {

struct T3 {  // Vec<Vec<Vec<i32>>> with Rust's bounds checks
    size_t n, L, P;
    std::vector<int> v;
    T3(size_t n_, size_t L_, size_t P_) : n(n_), L(L_), P(P_), v(n_ * L_ * P_, 0) {}
    inline int& at(size_t i, size_t j, size_t p) {
        if (i >= n || j >= L || p >= P) throw RefPanic("index out of bounds in the affine pathwise tensors");
        return v[(i * L + j) * P + p];
    }
    inline int at(size_t i, size_t j, size_t p) const {
        if (i >= n || j >= L || p >= P) throw RefPanic("index out of bounds in the affine pathwise tensors");
        return v[(i * L + j) * P + p];
    }
};

inline bool bit(const BitVec& b, size_t k) {
    if (k >= b.size()) throw RefPanic("BitVec index out of bounds");
    return b[k];
}
BitVec band(const BitVec& a, const BitVec& b) {
    BitVec r(a.size());
    for (size_t k = 0; k < a.size(); k++) r[k] = a[k] & b[k];
    return r;
}
std::vector<size_t> members(const BitVec& b) {
    std::vector<size_t> r;
    for (size_t k = 0; k < b.size(); k++)
        if (b[k]) r.push_back(k);
    return r;
}
size_t first_member(const BitVec& b) {
    for (size_t k = 0; k < b.size(); k++)
        if (b[k]) return k;
    throw RefPanic("position(|is_in| is_in).unwrap() on None");
}
inline int max3(int d, int u, int l) { return std::max(d, std::max(u, l)); }

// pathwise_alignment_gap.rs:19-541 / pathwise_alignment_gap_semi.rs:19-430
void fill(const std::vector<char>& sequence, const PathGraph& graph, const ScoreMatrix& sm, int o, int e, bool semi, T3& dpm,
          T3& x, T3& y) {
    const auto& lnz = graph.lnz;
    const auto& nwp = graph.nwp;
    const auto& path_node = graph.paths_nodes;
    const auto& alphas = graph.alphas;
    const size_t P = graph.paths_number, n = lnz.size(), L = sequence.size();

    for (size_t i = 0; i + 1 < n; i++) {
        for (size_t j = 0; j < L; j++) {
            if (i == 0 && j == 0) continue;  // (0, 0): zeros
            if (i == 0) {                    // (0, _)  …_gap.rs:26-34
                y.at(i, j, alphas.at(0)) = o + e * (int)j;
                dpm.at(i, j, alphas[0]) = y.at(i, j, alphas[0]);
                for (size_t k = alphas[0] + 1; k < P; k++) {
                    y.at(i, j, k) = y.at(i, j - 1, k);
                    dpm.at(i, j, k) = y.at(i, j, k);
                }
                continue;
            }
            if (j == 0) {
                if (semi) continue;  // `(_, 0) => dpm[0][0] = vec![0; path_number]`: the first column stays 0
                // (_, 0)  …_gap.rs:35-149
                if (!nwp[i]) {
                    BitVec common = band(path_node[i], path_node[i - 1]);
                    if (bit(common, alphas[i - 1])) {
                        for (size_t path = 0; path < P; path++)
                            if (common[path]) {
                                if (path == alphas[i])
                                    x.at(i, j, path) = (i == 1) ? o + e : x.at(i - 1, j, path) + e;
                                else
                                    x.at(i, j, path) = x.at(i - 1, j, path);
                                dpm.at(i, j, path) = x.at(i, j, path);
                            }
                    } else {
                        x.at(i, j, alphas[i]) = (i != 1) ? x.at(i - 1, j, alphas[i]) + x.at(i - 1, j, alphas[i - 1]) + e : o + e;
                        dpm.at(i, j, alphas[i]) = x.at(i, j, alphas[i]);
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != alphas[i]) {
                                x.at(i, j, path) = x.at(i - 1, j, path) - x.at(i - 1, j, alphas[i]);
                                dpm.at(i, j, path) = x.at(i, j, path);
                            }
                    }
                } else {
                    std::map<size_t, std::vector<size_t>> alphas_deltas;
                    for (const auto& pp : graph.get_preds_and_paths(i)) {
                        const size_t p = pp.first;
                        BitVec common = band(path_node[i], pp.second);
                        if (bit(common, alphas[p])) {
                            alphas_deltas[alphas[p]] = members(common);
                            x.at(i, j, alphas[p]) = (p == 0) ? o + e : x.at(p, j, alphas[p]) + e;
                            dpm.at(i, j, alphas[p]) = x.at(i, j, alphas[p]);
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != alphas[p]) {
                                    x.at(i, j, path) = x.at(p, j, path);
                                    dpm.at(i, j, path) = x.at(i, j, path);
                                }
                        } else {
                            const size_t ta = bit(common, alphas[i]) ? alphas[i] : first_member(common);
                            alphas_deltas[ta] = members(common);
                            x.at(i, j, ta) = (p == 0) ? o + e : x.at(p, j, ta) + x.at(p, j, alphas[p]) + e;
                            dpm.at(i, j, ta) = x.at(i, j, ta);
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ta) {
                                    x.at(i, j, path) = x.at(p, j, path) - x.at(p, j, ta);
                                    dpm.at(i, j, path) = x.at(i, j, path);
                                }
                        }
                    }
                    // remove multiple alpha  …_gap.rs:133-147
                    for (const auto& ad : alphas_deltas) {
                        const size_t a = ad.first;
                        if (a != alphas[i]) {
                            x.at(i, j, a) -= x.at(i, j, alphas[i]);
                            dpm.at(i, j, a) = x.at(i, j, a);
                            for (size_t path : ad.second)
                                if (path != a) {
                                    x.at(i, j, path) += x.at(i, j, a);
                                    dpm.at(i, j, path) = x.at(i, j, path);
                                }
                        }
                    }
                }
                continue;
            }
            // (_, _)
            const int sub = sm.get(lnz[i], sequence[j]);
            if (!nwp[i]) {
                BitVec common = band(path_node[i], path_node[i - 1]);
                const size_t ai = alphas[i], ap = alphas[i - 1];
                if (bit(common, ap)) {  // …_gap.rs:155-231
                    const int u_y = y.at(i - 1, j, ap) + e;
                    const int u_dpm = dpm.at(i - 1, j, ap) + o + e;
                    if (u_dpm >= u_y) {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) y.at(i, j, path) = dpm.at(i - 1, j, path);
                        y.at(i, j, ai) = u_dpm;
                    } else {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) y.at(i, j, path) = y.at(i - 1, j, path);
                        y.at(i, j, ai) = u_y;
                    }
                    const int u = y.at(i, j, ai);
                    const int l_x = x.at(i, j - 1, ai) + e;
                    const int l_dpm = dpm.at(i, j - 1, ai) + o + e;
                    if (l_dpm >= l_x) {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) x.at(i, j, path) = dpm.at(i, j - 1, path);
                        x.at(i, j, ai) = l_dpm;
                    } else {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) x.at(i, j, path) = x.at(i, j - 1, path);
                        x.at(i, j, ai) = l_x;
                    }
                    const int l = x.at(i, j, ai);
                    const int d = dpm.at(i - 1, j - 1, ap) + sub;
                    dpm.at(i, j, ai) = max3(d, u, l);
                    for (size_t path = 0; path < P; path++)
                        if (common[path] && path != ai) {
                            if (dpm.at(i, j, ai) == d)
                                dpm.at(i, j, path) = dpm.at(i - 1, j - 1, path);
                            else if (dpm.at(i, j, ai) == u)
                                dpm.at(i, j, path) = y.at(i, j, path);
                            else
                                dpm.at(i, j, path) = x.at(i, j, path);
                        }
                } else {  // …_gap.rs:232-304
                    const int u_y = y.at(i - 1, j, ap) + y.at(i - 1, j, ai) + e;
                    const int u_dpm = dpm.at(i - 1, j, ap) + dpm.at(i - 1, j, ai) + o + e;
                    if (u_dpm >= u_y) {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) y.at(i, j, path) = dpm.at(i - 1, j, path) - dpm.at(i - 1, j, ai);
                        y.at(i, j, ai) = u_dpm;
                    } else {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) y.at(i, j, path) = y.at(i - 1, j, path) - y.at(i - 1, j, ai);
                        y.at(i, j, ai) = u_y;
                    }
                    const int u = y.at(i, j, ai);
                    const int l_x = x.at(i, j - 1, ai) + e;
                    const int l_dpm = dpm.at(i, j - 1, ai) + o + e;
                    if (l_dpm >= l_x) {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) x.at(i, j, path) = dpm.at(i, j - 1, path);
                        x.at(i, j, ai) = l_dpm;
                    } else {
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ai) x.at(i, j, path) = x.at(i, j - 1, path);
                        x.at(i, j, ai) = l_x;
                    }
                    const int l = x.at(i, j, ai);
                    const int d = dpm.at(i - 1, j - 1, ap) + dpm.at(i - 1, j - 1, ai) + sub;
                    dpm.at(i, j, ai) = max3(d, u, l);
                    for (size_t path = 0; path < P; path++)
                        if (common[path] && path != ai) {
                            if (dpm.at(i, j, ai) == d)
                                dpm.at(i, j, path) = dpm.at(i - 1, j - 1, path) - dpm.at(i - 1, j - 1, ai);
                            else if (dpm.at(i, j, ai) == u)
                                dpm.at(i, j, path) = y.at(i, j, path);
                            else
                                dpm.at(i, j, path) = x.at(i, j, path);
                        }
                }
            } else {
                // multiple alphas possible  …_gap.rs:305-540
                std::map<size_t, std::vector<size_t>> alphas_deltas;
                const size_t ai = alphas[i];
                for (const auto& pp : graph.get_preds_and_paths(i)) {
                    const size_t p = pp.first;
                    BitVec common = band(path_node[i], pp.second);
                    const size_t ap = alphas[p];
                    if (bit(common, ap)) {  // :311-414
                        alphas_deltas[ap] = members(common);
                        const int u_y = y.at(p, j, ap) + e;
                        const int u_dpm = dpm.at(p, j, ap) + o + e;
                        if (u_dpm >= u_y) {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ap) y.at(i, j, path) = dpm.at(p, j, path);
                            y.at(i, j, ap) = u_dpm;
                        } else {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ai) y.at(i, j, path) = y.at(p, j, path);  // `path != alphas[i]`, :338
                            y.at(i, j, ap) = u_y;
                        }
                        const int u = y.at(i, j, ap);
                        const int l_x = (ap == ai) ? x.at(i, j - 1, ap) + e : x.at(i, j - 1, ap) + x.at(i, j - 1, ai) + e;
                        const int l_dpm = (ap == ai) ? dpm.at(i, j - 1, ap) + o + e : dpm.at(i, j - 1, ai) + dpm.at(i, j - 1, ap) + o + e;
                        if (l_dpm >= l_x) {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ap)
                                    x.at(i, j, path) = (ap == ai) ? dpm.at(i, j - 1, path) : dpm.at(i, j - 1, path) - dpm.at(i, j - 1, ap);
                            x.at(i, j, ap) = l_dpm;
                        } else {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ap)
                                    x.at(i, j, path) = (ap == ai) ? x.at(i, j - 1, path) : x.at(i, j - 1, path) - x.at(i, j - 1, ap);
                            x.at(i, j, ap) = l_x;
                        }
                        const int l = x.at(i, j, ap);
                        const int d = dpm.at(p, j - 1, ap) + sub;
                        dpm.at(i, j, ap) = max3(d, u, l);
                        for (size_t path = 0; path < P; path++)
                            if (common[path] && path != ap) {
                                if (dpm.at(i, j, ap) == d)
                                    dpm.at(i, j, path) = dpm.at(p, j - 1, path);
                                else if (dpm.at(i, j, ap) == u)
                                    dpm.at(i, j, path) = y.at(i, j, path);
                                else
                                    dpm.at(i, j, path) = x.at(i, j, path);
                            }
                    } else {  // set new alpha  :415-519
                        const size_t ta = bit(common, ai) ? ai : first_member(common);
                        alphas_deltas[ta] = members(common);
                        const int u_y = y.at(p, j, ap) + y.at(p, j, ta) + e;
                        const int u_dpm = dpm.at(p, j, ap) + dpm.at(p, j, ta) + o + e;
                        if (u_dpm >= u_y) {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ta) y.at(i, j, path) = dpm.at(p, j, path) - dpm.at(p, j, ta);
                            y.at(i, j, ta) = u_dpm;
                        } else {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ta) y.at(i, j, path) = y.at(p, j, path) - y.at(p, j, ta);
                            y.at(i, j, ta) = u_y;
                        }
                        const int u = y.at(i, j, ta);
                        const int l_x = (ai == ta) ? x.at(i, j - 1, ai) + e : x.at(i, j - 1, ai) + x.at(i, j - 1, ta) + e;
                        const int l_dpm = (ai == ta) ? dpm.at(i, j - 1, ai) + o + e : dpm.at(i, j - 1, ai) + dpm.at(i, j - 1, ta) + o + e;
                        if (l_dpm >= l_x) {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ta)
                                    x.at(i, j, path) = (ta == ai) ? dpm.at(i, j - 1, path) : dpm.at(i, j - 1, path) - dpm.at(i, j - 1, ta);
                            x.at(i, j, ta) = l_dpm;
                        } else {
                            for (size_t path = 0; path < P; path++)
                                if (common[path] && path != ta)
                                    x.at(i, j, path) = (ta == ai) ? x.at(i, j - 1, path) : x.at(i, j - 1, path) - x.at(i, j - 1, ta);
                            x.at(i, j, ta) = l_x;
                        }
                        const int l = x.at(i, j, ta);
                        const int d = dpm.at(p, j - 1, ap) + dpm.at(p, j - 1, ta) + sub;
                        dpm.at(i, j, ta) = max3(d, u, l);
                        for (size_t path = 0; path < P; path++)
                            if (path != ta && common[path]) {
                                if (dpm.at(i, j, ta) == d)
                                    dpm.at(i, j, path) = dpm.at(p, j - 1, path) - dpm.at(p, j - 1, ta);
                                else if (dpm.at(i, j, ta) == u)
                                    dpm.at(i, j, path) = y.at(i, j, path);
                                else
                                    dpm.at(i, j, path) = x.at(i, j, path);
                            }
                    }
                }
                // :520-538
                for (const auto& ad : alphas_deltas) {
                    const size_t a = ad.first;
                    if (a != ai) {
                        dpm.at(i, j, a) -= dpm.at(i, j, ai);
                        x.at(i, j, a) -= x.at(i, j, ai);
                        y.at(i, j, a) -= y.at(i, j, ai);
                        for (size_t path : ad.second)
                            if (path != a) {
                                dpm.at(i, j, path) += dpm.at(i, j, a);
                                x.at(i, j, path) += x.at(i, j, a);
                                y.at(i, j, path) += y.at(i, j, a);
                            }
                    }
                }
            }
        }
    }
}

// shared body of build_alignment_gap (pathwise_alignment_output.rs:186-318) and build_alignment_semiglobal_gap (:320-451):
// walks back from (i, L-1) and leaves i / the reversed op list for the caller's tail handling
void walk(const T3& dpm, const T3& x, const T3& y, const PathGraph& g, size_t best_path, size_t& i, size_t& j,
          std::vector<char>& cigar) {
    const auto& alphas = g.alphas;
    const auto& nwp = g.nwp;
    auto abs_at = [&](size_t ii, size_t jj) {
        return alphas.at(ii) == best_path ? dpm.at(ii, jj, best_path) : dpm.at(ii, jj, best_path) + dpm.at(ii, jj, alphas[ii]);
    };
    size_t guard = 0;
    const size_t guard_max = 8 * (dpm.n + dpm.L) + 64;
    auto tick = [&]() {
        if (++guard > guard_max) throw RefPanic("oracle: the reference would loop forever here");
    };
    while (i != 0 && j != 0) {
        tick();
        const int curr_score = abs_at(i, j);
        bool has_pred = false;
        size_t predecessor = 0;
        int d = 0, u = 0, l = 0;
        if (!nwp[i]) {
            d = abs_at(i - 1, j - 1);
            u = abs_at(i - 1, j);
            l = abs_at(i, j - 1);
        } else {
            for (const auto& pp : g.get_preds_and_paths(i))
                if (bit(pp.second, best_path)) {
                    has_pred = true;
                    predecessor = pp.first;
                    d = abs_at(pp.first, j - 1);
                    u = abs_at(pp.first, j);
                    l = abs_at(i, j - 1);
                }
        }
        const int mx = max3(d, u, l);
        if (mx == d) {
            cigar.push_back(curr_score < d ? 'd' : 'D');
            i = has_pred ? predecessor : i - 1;
            j -= 1;
        } else if (mx == u) {
            cigar.push_back('U');
            i = has_pred ? predecessor : i - 1;
            while (dpm.at(i, j, best_path) < y.at(i, j, best_path)) {  // raw delta entries, as in the reference
                tick();
                cigar.push_back('U');
                if (nwp.at(i)) {
                    for (const auto& pp : g.get_preds_and_paths(i))
                        if (bit(pp.second, best_path)) {
                            has_pred = true;
                            predecessor = pp.first;
                        }
                } else {
                    if (i == 0) throw RefPanic("attempt to subtract with overflow (i - 1)");
                    has_pred = true;
                    predecessor = i - 1;
                }
                if (!has_pred) throw RefPanic("predecessor.unwrap() on None");
                i = predecessor;
            }
        } else {
            cigar.push_back('L');
            j -= 1;
            while (dpm.at(i, j, best_path) < x.at(i, j, best_path)) {
                tick();
                cigar.push_back('L');
                if (j == 0) throw RefPanic("attempt to subtract with overflow (j - 1)");
                j -= 1;
            }
        }
    }
    while (j > 0) {
        cigar.push_back('L');
        j -= 1;
    }
}

size_t walk_to_source(const PathGraph& g, size_t i, size_t best_path) {
    size_t steps = 0;
    while (i > 0) {
        if (g.nwp[i]) {
            for (const auto& pp : g.get_preds_and_paths(i))
                if (bit(pp.second, best_path)) i = pp.first;
        } else {
            i -= 1;
        }
        steps++;
        if (steps > g.lnz.size() + 1) throw RefPanic("oracle: the reference would loop forever here (no predecessor on the best path)");
    }
    return steps;
}

}  // namespace

// pathwise_alignment_gap.rs:4-574 (mode 6): returns the best path; `line` = what exec println!s
size_t pathwise_alignment_gap_exec(const std::vector<char>& sequence, const PathGraph& g, const ScoreMatrix& sm, int o, int e,
                                   std::string& line, int* best_score) {
    const size_t n = g.lnz.size(), L = sequence.size(), P = g.paths_number;
    T3 dpm(n, L, P), x(n, L, P), y(n, L, P);
    fill(sequence, g, sm, o, e, false, dpm, x, y);
    std::vector<int> results(P, 0);
    for (const auto& pp : g.get_preds_and_paths(n - 1))
        for (size_t path = 0; path < P; path++)
            if (pp.second[path]) {
                const size_t pred = pp.first;
                results[path] = (path == g.alphas[pred]) ? dpm.at(pred, L - 1, path) : dpm.at(pred, L - 1, path) + dpm.at(pred, L - 1, g.alphas[pred]);
            }
    size_t best_path = 0;  // max of (score, path): the highest path id wins ties
    for (size_t path = 1; path < P; path++)
        if (results[path] >= results[best_path]) best_path = path;
    if (P == 0) throw RefPanic("best_path.unwrap() on None");
    if (best_score) *best_score = results[best_path];
    // build_alignment_gap
    std::vector<char> cigar;
    size_t i = 0;
    for (const auto& pp : g.get_preds_and_paths(n - 1))
        if (bit(pp.second, best_path)) i = pp.first;
    size_t j = L - 1;
    walk(dpm, x, y, g, best_path, i, j, cigar);
    while (i > 0) {
        cigar.push_back('U');
        i -= 1;
    }
    std::reverse(cigar.begin(), cigar.end());
    if (!cigar.empty()) cigar.pop_back();
    line = build_cigar(cigar);
    return best_path;
}

// pathwise_alignment_gap_semi.rs:5-473 (mode 7)
size_t pathwise_alignment_gap_semi_exec(const std::vector<char>& sequence, const PathGraph& g, const ScoreMatrix& sm, int o, int e,
                                        std::string& line, int* best_score) {
    const size_t n = g.lnz.size(), L = sequence.size(), P = g.paths_number;
    T3 dpm(n, L, P), x(n, L, P), y(n, L, P);
    fill(sequence, g, sm, o, e, true, dpm, x, y);
    // best_ending_node  :447-473
    bool have = false;
    int mx = 0;
    size_t ending_node = 0, chosen_path = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        std::vector<int> abs_scores(P);
        for (size_t path = 0; path < P; path++) abs_scores[path] = dpm.at(i, L - 1, path);
        for (size_t path = 0; path < P; path++)
            if (g.paths_nodes[i][path] && path != g.alphas.at(i)) abs_scores[path] = abs_scores[path] + abs_scores.at(g.alphas[i]);
        size_t bp = 0;
        for (size_t path = 1; path < P; path++)
            if (abs_scores[path] >= abs_scores[bp]) bp = path;
        if (!have || abs_scores[bp] > mx) {
            have = true;
            mx = abs_scores[bp];
            ending_node = i;
            chosen_path = bp;
        }
    }
    if (best_score) *best_score = mx;
    // build_alignment_semiglobal_gap
    std::vector<char> cigar;
    size_t i = ending_node, j = L - 1;
    walk(dpm, x, y, g, chosen_path, i, j, cigar);
    std::reverse(cigar.begin(), cigar.end());
    const size_t starting_node = walk_to_source(g, i, chosen_path);
    const size_t final_node = walk_to_source(g, ending_node, chosen_path);
    line = build_cigar(cigar) + "\t(" + std::to_string(starting_node) + " " + std::to_string(final_node) + ")";
    return chosen_path;
}

}  // namespace rgo
