// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). POA family (modes 0-3) and their tracebacks:
// global_abpoa.rs, local_poa.rs, gap_global_abpoa.rs, gap_local_poa.rs, bitfield_path.rs, gaf_output.rs.
#include <algorithm>
#include <charconv>
#include <cstdlib>

#include "oracle.hpp"

namespace rgo {

// ------------------------------------------------------------------------------- gaf_output.rs:70-94
std::string GAFStruct::to_string() const {
    std::string pm;
    for (size_t k = 0; k < path.size(); k++) {
        if (k) pm += ">";
        pm += std::to_string(path[k]);
    }
    std::string s = query_name + "\t" + std::to_string(query_length) + "\t" + std::to_string(query_start) + "\t" +
                    std::to_string(query_end) + "\t" + std::string(1, strand) + "\t>" + pm + "\t" +
                    std::to_string(path_length) + "\t" + std::to_string(path_start) + "\t" +
                    std::to_string(path_end) + "\t" + std::to_string(residue_matches_number) + "\t" +
                    alignment_block_length + "\t" + mapping_quality + "\t" + comments;
    return s;
}
// Rust `{}` on f32: shortest decimal that round-trips, never scientific.
std::string f32_display(float v) {
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

// pathwise_alignment_output.rs:471-556
std::string build_cigar(const std::vector<char>& cigar) {
    std::string out;
    size_t d_count = 0, u_count = 0, l_count = 0, mm_count = 0;
    auto flush = [&](size_t& c, char sym) {
        if (c != 0) {
            out += std::to_string(c) + sym;
            c = 0;
        }
    };
    for (char ch : cigar) {
        switch (ch) {
            case 'D':
                flush(u_count, 'I');
                flush(l_count, 'D');
                flush(mm_count, 'X');
                d_count++;
                break;
            case 'U':
                flush(d_count, 'M');
                flush(l_count, 'D');
                flush(mm_count, 'X');
                u_count++;
                break;
            case 'd':
                flush(d_count, 'M');
                flush(l_count, 'D');
                flush(u_count, 'I');
                mm_count++;
                break;
            default:
                flush(d_count, 'M');
                flush(u_count, 'I');
                flush(mm_count, 'X');
                l_count++;
                break;
        }
    }
    flush(d_count, 'M');
    flush(u_count, 'I');
    flush(l_count, 'D');
    flush(mm_count, 'X');
    return out;
}

// ------------------------------------------------------------------------------- bitfield_path.rs
// 32-bit cell: bits 31..16 predecessor (u16, TRUNCATING — F3), bits 15..0 direction code.
typedef uint64_t PathCell;
// RGO_PRED32=1 (tests on graphs above 65 535 rows, outside the reference's valid domain): keep the full
// predecessor instead of truncating it to u16. The cell then needs 64 bits.
static bool pred32_mode() {
    static const bool on = std::getenv("RGO_PRED32") != nullptr;
    return on;
}
static inline PathCell set_path_cell(size_t pred, char dir) {
    uint32_t d;
    switch (dir) {
        case 'O': d = 0; break;
        case 'D': d = 1; break;
        case 'd': d = 2; break;
        case 'L': d = 3; break;
        case 'U': d = 4; break;
        case 'X': d = 5; break;
        case 'Y': d = 6; break;
        case 'M': d = 7; break;
        default: throw RefPanic("impossible direction char");  // bitfield_path.rs:13
    }
    return ((uint64_t)(pred32_mode() ? (uint32_t)pred : (uint32_t)(uint16_t)pred) << 16) | d;
}
static inline size_t pred_from_bitvec(PathCell c) { return (size_t)(c >> 16); }
static inline char dir_from_bitvec(PathCell c) {
    static const char t[8] = {'O', 'D', 'd', 'L', 'U', 'X', 'Y', 'M'};
    uint32_t d = c & 0xffff;
    if (d > 7) throw RefPanic("impossible direction bitslice");
    return t[d];
}

// utils.rs:129-140
static inline std::pair<int, char> get_max_d_u_l(int d, int u, int l) {
    if (d < u) {
        if (u < l) return {l, 'L'};
        return {u, 'U'};
    }
    if (d < l) return {l, 'L'};
    return {d, 'D'};
}

// gaf_output.rs:867-874
static size_t node_start(const std::vector<std::string>& hofp, size_t row) {
    const std::string& handle_id = hofp.at(row);
    size_t i = row;
    while (hofp.at(i) == handle_id && i > 0) i -= 1;
    return row - i;
}
// gaf_output.rs:876-892
static std::string set_cigar_substring(int count_m, int count_i, int count_d, const std::string& cs) {
    if ((count_m * count_i) + (count_i * count_d) + (count_m * count_d) != 0)
        throw RefPanic("wrong format in cigar string");
    if (count_m > 0) return std::to_string(count_m) + "M" + cs;
    if (count_i > 0) return std::to_string(count_i) + "I" + cs;
    if (count_d > 0) return std::to_string(count_d) + "D" + cs;
    return cs;
}
static std::string join_but_last(const std::vector<std::string>& cigars) {
    std::string s;
    for (size_t k = 0; k + 1 < cigars.size(); k++) {
        if (k) s += ",";
        s += cigars[k];
    }
    return s;
}
static size_t parse_handle(const std::string& id) {
    size_t v = 0;
    auto r = std::from_chars(id.data(), id.data() + id.size(), v);
    if (r.ec != std::errc() || r.ptr != id.data() + id.size()) throw RefPanic("id.parse::<usize>().unwrap()");
    return v;
}
static char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

typedef std::vector<std::vector<PathCell>> PathMat;
typedef std::vector<std::pair<size_t, size_t>> Ampl;

// Shared state of the four bit-field tracebacks (gaf_output.rs:96-637), which differ only in the column remap
// (banded vs full rows) and in how L / U runs are consumed (affine vs linear).
struct TraceAcc {
    std::vector<const std::string*> handle_id_alignment;
    std::vector<std::string> cigars;
    std::string cigar;
    int count_m = 0, count_i = 0, count_d = 0;
    const std::string* curr_handle = nullptr;  // "" initially
    char last_dir = ' ';
    size_t path_length = 0, residue_matching = 0;
    void on_cell(const std::string& h, char dir) {
        if (curr_handle == nullptr ? !h.empty() : h != *curr_handle) {
            cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
            cigars.insert(cigars.begin(), cigar);
            cigar.clear();
            count_m = count_i = count_d = 0;
        }
        curr_handle = &h;
        if (upper(dir) != upper(last_dir)) {
            cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
            count_m = count_i = count_d = 0;
        }
        last_dir = dir;
    }
    GAFStruct finish(const std::vector<char>& sequence, const std::string& name, size_t query_start,
                     size_t query_end, bool amb_mode, const std::vector<std::string>& hofp, size_t row,
                     size_t last_row) {
        cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
        cigars.insert(cigars.begin(), cigar);
        // handle_id_alignment.dedup(); reverse()
        std::vector<const std::string*> hd;
        for (auto* h : handle_id_alignment)
            if (hd.empty() || *hd.back() != *h) hd.push_back(h);
        std::reverse(hd.begin(), hd.end());
        GAFStruct g;
        g.query_name = name;
        g.query_length = sequence.size() - 1;
        g.query_start = query_start;
        g.query_end = query_end;
        g.strand = amb_mode ? '-' : '+';
        g.path.clear();
        for (auto* h : hd) g.path.push_back(parse_handle(*h));
        g.path_length = path_length;
        g.path_start = node_start(hofp, row);
        g.path_end = node_start(hofp, last_row);
        g.residue_matches_number = residue_matching;
        g.alignment_block_length = "*";
        g.mapping_quality = "*";
        g.comments = join_but_last(cigars);
        return g;
    }
};

// gaf_output.rs:96-253
static GAFStruct gaf_of_gap_abpoa(const PathMat& path, const PathMat& path_x, const PathMat& path_y,
                                  const std::vector<char>& sequence, const std::string& name, const Ampl& ampl,
                                  size_t last_row, size_t last_col, bool amb_mode,
                                  const std::vector<std::string>& hofp) {
    size_t col = last_col, row = last_row;
    TraceAcc t;
    while (dir_from_bitvec(path[row][col]) != 'O') {
        PathCell curr = path[row][col];
        size_t pred = pred_from_bitvec(curr);
        char dir = dir_from_bitvec(curr);
        t.on_cell(hofp.at(row), dir);
        size_t p_left = ampl[pred].first;
        size_t j_pos = ampl[row].first < p_left ? col - (p_left - ampl[row].first) : col + (ampl[row].first - p_left);
        switch (dir) {
            case 'D':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col = j_pos - 1;
                t.count_m++;
                t.path_length++;
                t.residue_matching++;
                break;
            case 'd':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col = j_pos - 1;
                t.count_m++;
                t.path_length++;
                break;
            case 'L':
                if (dir_from_bitvec(path_x[row][col]) == 'X') {
                    while (dir_from_bitvec(path_x[row][col]) == 'X') {
                        t.count_d++;
                        col -= 1;
                    }
                } else {
                    t.count_d++;
                    col -= 1;
                }
                break;
            case 'U':
                if (dir_from_bitvec(path_y[row][col]) == 'Y') {
                    while (dir_from_bitvec(path_y[row][col]) == 'Y') {
                        size_t left_row = ampl[row].first;
                        size_t p = pred_from_bitvec(path_y[row][col]);
                        size_t left_p = ampl[p].first;
                        size_t jp = left_p < left_row ? col + (left_row - left_p) : col - (left_p - left_row);
                        t.handle_id_alignment.push_back(&hofp.at(row));
                        t.count_i++;
                        t.path_length++;
                        col = jp;
                        row = p;
                    }
                } else {
                    t.handle_id_alignment.push_back(&hofp.at(row));
                    t.count_i++;
                    t.path_length++;
                    row = pred;
                    col = j_pos;
                }
                break;
            default: throw RefPanic("impossible value in poa path");
        }
        if (row >= path.size() || col >= path[row].size()) throw RefPanic("index out of bounds in gaf_of_gap_abpoa");
    }
    return t.finish(sequence, name, col, last_col + ampl.at(last_row).first, amb_mode, hofp, row, last_row);
}

// gaf_output.rs:254-381
static GAFStruct gaf_of_global_abpoa(const PathMat& path, const std::vector<char>& sequence,
                                     const std::string& name, const Ampl& ampl, size_t last_row, size_t last_col,
                                     bool amb_mode, const std::vector<std::string>& hofp) {
    size_t col = last_col, row = last_row;
    TraceAcc t;
    while (dir_from_bitvec(path[row][col]) != 'O') {
        PathCell curr = path[row][col];
        size_t pred = pred_from_bitvec(curr);
        char dir = dir_from_bitvec(curr);
        t.on_cell(hofp.at(row), dir);
        size_t p_left = ampl[pred].first;
        size_t j_pos = ampl[row].first < p_left ? col - (p_left - ampl[row].first) : col + (ampl[row].first - p_left);
        switch (dir) {
            case 'D':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col = j_pos - 1;
                t.count_m++;
                t.path_length++;
                t.residue_matching++;
                break;
            case 'd':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col = j_pos - 1;
                t.count_m++;
                t.path_length++;
                break;
            case 'L':
                col -= 1;
                t.count_d++;
                break;
            case 'U':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col = j_pos;
                t.count_i++;
                t.path_length++;
                break;
            default: throw RefPanic("impossible value in poa path");
        }
        if (row >= path.size() || col >= path[row].size()) throw RefPanic("index out of bounds in gaf_of_global_abpoa");
    }
    return t.finish(sequence, name, col, last_col + ampl.at(last_row).first, amb_mode, hofp, row, last_row);
}

// gaf_output.rs:383-500
static GAFStruct gaf_of_local_poa(const PathMat& path, const std::vector<char>& sequence, const std::string& name,
                                  size_t last_row, size_t last_col, bool amb_mode,
                                  const std::vector<std::string>& hofp) {
    size_t col = last_col, row = last_row;
    TraceAcc t;
    while (dir_from_bitvec(path[row][col]) != 'O') {
        PathCell curr = path[row][col];
        size_t pred = pred_from_bitvec(curr);
        char dir = dir_from_bitvec(curr);
        t.on_cell(hofp.at(row), dir);
        switch (dir) {
            case 'D':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col -= 1;
                t.count_m++;
                t.path_length++;
                t.residue_matching++;
                break;
            case 'd':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col -= 1;
                t.count_m++;
                t.path_length++;
                break;
            case 'L':
                col -= 1;
                t.count_d++;
                break;
            case 'U':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                t.count_i++;
                t.path_length++;
                break;
            default: throw RefPanic("impossible value in poa path");
        }
    }
    return t.finish(sequence, name, col, last_col, amb_mode, hofp, row, last_row);
}

// gaf_output.rs:502-637
static GAFStruct gaf_of_gap_local_poa(const PathMat& path, const PathMat& path_x, const PathMat& path_y,
                                      const std::vector<char>& sequence, const std::string& name, size_t last_row,
                                      size_t last_col, bool amb_mode, const std::vector<std::string>& hofp) {
    size_t col = last_col, row = last_row;
    TraceAcc t;
    while (dir_from_bitvec(path[row][col]) != 'O') {
        PathCell curr = path[row][col];
        size_t pred = pred_from_bitvec(curr);
        char dir = dir_from_bitvec(curr);
        t.on_cell(hofp.at(row), dir);
        switch (dir) {
            case 'D':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col -= 1;
                t.count_m++;
                t.path_length++;
                t.residue_matching++;
                break;
            case 'd':
                t.handle_id_alignment.push_back(&hofp.at(row));
                row = pred;
                col -= 1;
                t.count_m++;
                t.path_length++;
                break;
            case 'L':
                if (dir_from_bitvec(path_x[row][col]) == 'X') {
                    while (dir_from_bitvec(path_x[row][col]) == 'X') {
                        t.count_d++;
                        col -= 1;
                    }
                } else {
                    t.count_d++;
                    col -= 1;
                }
                break;
            case 'U':
                if (dir_from_bitvec(path_y[row][col]) == 'Y') {
                    while (dir_from_bitvec(path_y[row][col]) == 'Y') {
                        size_t p = pred_from_bitvec(path_y[row][col]);
                        t.handle_id_alignment.push_back(&hofp.at(row));
                        row = p;
                        t.count_i++;
                        t.path_length++;
                    }
                } else {
                    t.handle_id_alignment.push_back(&hofp.at(row));
                    t.count_i++;
                    t.path_length++;
                    row = pred;
                }
                break;
            default: throw RefPanic("impossible value in poa path");
        }
    }
    return t.finish(sequence, name, col, last_col, amb_mode, hofp, row, last_row);
}

// f32 trace cell decode: `val.to_string().split('.')` (gaf_output.rs:664-669, 783-786)
static void decode_f32_cell(float val, size_t& pred, int& dir) {
    std::string s = f32_display(val);
    size_t dot = s.find('.');
    if (dot == std::string::npos) throw RefPanic("index out of bounds: pred_dir[1]");
    std::string a = s.substr(0, dot), b = s.substr(dot + 1);
    if (b.find('.') != std::string::npos) throw RefPanic("parse error");
    unsigned long long pv = 0;
    auto r = std::from_chars(a.data(), a.data() + a.size(), pv);
    if (a.empty() || r.ec != std::errc() || r.ptr != a.data() + a.size()) throw RefPanic("pred parse::<usize>() failed");
    int dv = 0;
    auto r2 = std::from_chars(b.data(), b.data() + b.size(), dv);
    if (b.empty() || r2.ec != std::errc() || r2.ptr != b.data() + b.size()) throw RefPanic("dir parse::<i32>() failed");
    pred = (size_t)pv;
    dir = dv;
}

// gaf_output.rs:639-751
static GAFStruct gaf_of_local_poa_simd(const std::vector<std::vector<float>>& path, const std::vector<char>& sequence,
                                       const std::string& name, size_t last_row, size_t last_col, bool amb_mode,
                                       const std::vector<std::string>& hofp) {
    size_t col = last_col, row = last_row;
    std::vector<const std::string*> handle_id_alignment;
    std::vector<std::string> cigars;
    std::string cigar;
    int count_m = 0, count_i = 0, count_d = 0;
    const std::string* curr_handle = nullptr;
    int last_dir = -1;
    size_t path_length = 0, residue_matching = 0;
    while (path[row][col] != 0.0f) {
        size_t pred;
        int dir;
        decode_f32_cell(path[row][col], pred, dir);
        const std::string& h = hofp.at(row);
        if (curr_handle == nullptr ? !h.empty() : h != *curr_handle) {
            cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
            cigars.insert(cigars.begin(), cigar);
            cigar.clear();
            count_m = count_i = count_d = 0;
        }
        curr_handle = &h;
        if (dir != last_dir) {
            cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
            count_m = count_i = count_d = 0;
        }
        last_dir = dir;
        switch (dir) {
            case 1:
                handle_id_alignment.push_back(&h);
                row = pred;
                col -= 1;
                count_m++;
                path_length++;
                residue_matching++;
                break;
            case 3:
                col -= 1;
                count_d++;
                break;
            case 2:
                handle_id_alignment.push_back(&h);
                row = pred;
                count_i++;
                path_length++;
                break;
            default: throw RefPanic("impossible value in poa path");
        }
    }
    cigar = set_cigar_substring(count_m, count_i, count_d, cigar);
    cigars.insert(cigars.begin(), cigar);
    std::vector<const std::string*> hd;
    for (auto* h : handle_id_alignment)
        if (hd.empty() || *hd.back() != *h) hd.push_back(h);
    std::reverse(hd.begin(), hd.end());
    GAFStruct g;
    g.query_name = name;
    g.query_length = sequence.size() - 1;
    g.query_start = col;
    g.query_end = last_col;
    g.strand = amb_mode ? '-' : '+';
    g.path.clear();
    for (auto* h : hd) g.path.push_back(parse_handle(*h));
    g.path_length = path_length;
    g.path_start = node_start(hofp, row);
    g.path_end = node_start(hofp, last_row);
    g.residue_matches_number = residue_matching;
    g.alignment_block_length = "*";
    g.mapping_quality = "*";
    g.comments = join_but_last(cigars);
    return g;
}

// gaf_output.rs:753-865
static GAFStruct gaf_of_global_abpoa_simd(const std::vector<std::vector<float>>& path,
                                          const std::vector<char>& sequence, const std::string& name,
                                          size_t last_row, size_t last_col, bool amb_mode,
                                          const std::vector<std::string>& hofp, const std::vector<char>& lnz,
                                          float best_score, std::string& out) {
    size_t col = last_col, row = last_row;
    std::vector<const std::string*> handle_id_alignment;
    std::vector<char> cigar, path_sequence;
    size_t path_length = 0, residue_matching = 0;
    bool out_ok = true;
    while (path[row][col] != 0.0f) {
        float val = path[row][col];
        if (val == -1.0f) {
            out_ok = false;
            break;
        }
        size_t pred;
        int dir;
        decode_f32_cell(val, pred, dir);
        switch (dir) {
            case 1:
                handle_id_alignment.push_back(&hofp.at(row));
                path_sequence.push_back(lnz[row]);
                row = pred;
                col -= 1;
                cigar.push_back(lnz[row] == sequence[col] ? 'D' : 'd');
                path_length++;
                residue_matching++;
                break;
            case 3:
                col -= 1;
                cigar.push_back('L');
                break;
            case 2:
                handle_id_alignment.push_back(&hofp.at(row));
                path_sequence.push_back(lnz[row]);
                row = pred;
                cigar.push_back('U');
                path_length++;
                break;
            default: throw RefPanic("impossible value in poa path");
        }
    }
    if (!out_ok) {
        out += "band not enough for correct output\n";
        return GAFStruct();
    }
    std::reverse(cigar.begin(), cigar.end());
    std::string cigar_out = build_cigar(cigar);
    std::reverse(path_sequence.begin(), path_sequence.end());
    std::vector<const std::string*> hd;
    for (auto* h : handle_id_alignment)
        if (hd.empty() || *hd.back() != *h) hd.push_back(h);
    std::reverse(hd.begin(), hd.end());
    GAFStruct g;
    g.query_name = name;
    g.query_length = sequence.size() - 1;
    g.query_start = col;
    g.query_end = last_col;
    g.strand = amb_mode ? '-' : '+';
    g.path.clear();
    for (auto* h : hd) g.path.push_back(parse_handle(*h));
    g.path_length = path_length;
    g.path_start = node_start(hofp, row);
    g.path_end = node_start(hofp, last_row);
    g.residue_matches_number = residue_matching;
    g.alignment_block_length = "*";
    g.mapping_quality = "*";
    g.comments = cigar_out + ", score: " + f32_display(best_score) + "\t" +
                 std::string(path_sequence.begin(), path_sequence.end());
    return g;
}

// ------------------------------------------------------------------------------- mode 0, AVX2 semantics
// global_abpoa.rs:10-257. 8-lane blocks are evaluated lane by lane; every f32 value is an exact small
// integer so evaluation order inside a block does not change results.
PoaResult global_abpoa_exec_simd(const std::vector<char>& read, const std::string& name, size_t number,
                                 const LnzGraph& graph, const ScoreMatrix& sm, size_t bta, bool amb_mode,
                                 const std::vector<std::string>& hofp, const std::vector<size_t>& r_values,
                                 std::string& out) {
    const size_t L = read.size(), n = graph.lnz.size();
    const auto& lnz = graph.lnz;
    PoaResult res;
    float min_score = 2.0f * (float)L * (float)sm.get(read.at(1), '-');
    std::vector<std::vector<float>> m(n, std::vector<float>(L, min_score));
    std::vector<std::vector<float>> path(n, std::vector<float>(L, -1.0f));
    std::vector<size_t> best_scoring_pos(n, 0);
    m[0][0] = 0.0f;
    path[0][0] = 0.0f;
    for (size_t i = 1; i + 1 < n; i++) {
        if (!graph.nwp[i]) {
            m[i][0] = m[i - 1][0] + (float)sm.get(lnz[i], '-');
            path[i][0] = (float)(i - 1) + 0.2f;
        } else {
            const auto& pred = graph.preds(i);
            size_t best_p = *std::min_element(pred.begin(), pred.end());
            m[i][0] = m[best_p][0] + (float)sm.get(lnz[i], '-');
            path[i][0] = (float)best_p + 0.2f;
        }
    }
    static const std::vector<size_t> empty;
    {
        auto lr = set_ampl_for_row(0, empty, r_values[0], best_scoring_pos, L, bta, true);
        for (size_t j = 1; j < lr.second; j++) {
            m[0][j] = m[0][j - 1] + (float)sm.get(read[j], '-');
            path[0][j] = 0.3f;
        }
        res.cells += lr.second - lr.first;
    }
    for (size_t i = 1; i + 1 < n; i++) {
        const std::vector<size_t>& p_arr = graph.nwp[i] ? graph.preds(i) : empty;
        auto lr = set_ampl_for_row(i, p_arr, r_values[i], best_scoring_pos, L, bta, true);
        size_t left = lr.first, right = lr.second;
        res.cells += right - left;
        size_t best_col = left;
        size_t start = left == 0 ? 1 : left;
        size_t end = right == L ? ((right - start) / 8) * 8 + start : right;
        float us_update = (float)sm.get(lnz[i], '-');
        for (size_t j = start; j < end; j += 8) {
            float ds_update[8];
            for (size_t k = 0; k < 8; k++) ds_update[k] = (float)sm.get(lnz[i], read.at(j + k));
            for (size_t k = 0; k < 8; k++) {
                size_t c = j + k;
                if (!graph.nwp[i]) {
                    float us = m[i - 1][c] + us_update;
                    float ds = m[i - 1][c - 1] + ds_update[k];
                    bool best_choice = ds > us;
                    m[i][c] = best_choice ? ds : us;
                    path[i][c] = (float)(i - 1) + (best_choice ? 0.1f : 0.2f);
                } else {
                    const auto& preds = graph.preds(i);
                    float best_us = m[preds[0]][c], best_ds = m[preds[0]][c - 1];
                    float pred_best_us = (float)preds[0], pred_best_ds = (float)preds[0];
                    for (size_t q = 1; q < preds.size(); q++) {
                        size_t p = preds[q];
                        float us = m[p][c], ds = m[p][c - 1];
                        if (us > best_us) {
                            best_us = us;
                            pred_best_us = (float)p;
                        }
                        if (ds > best_ds) {
                            best_ds = ds;
                            pred_best_ds = (float)p;
                        }
                    }
                    best_us = best_us + us_update;
                    best_ds = best_ds + ds_update[k];
                    bool best_choice = best_ds > best_us;
                    m[i][c] = best_choice ? best_ds : best_us;
                    pred_best_ds = pred_best_ds + 0.1f;
                    pred_best_us = pred_best_us + 0.2f;
                    path[i][c] = best_choice ? pred_best_ds : pred_best_us;
                }
            }
            for (size_t idx = j; idx < j + 8; idx++) {
                float l = m[i][idx - 1] + (float)sm.get(read[j], '-');  // read[j], not read[idx] (global_abpoa.rs:157)
                if (l > m[i][idx]) {
                    m[i][idx] = l;
                    path[i][idx] = (float)i + 0.3f;
                }
                if (m[i][idx] >= m[i][best_col]) best_col = idx;
            }
        }
        if (end < right) {
            for (size_t j = end; j < right; j++) {
                if (!graph.nwp[i]) {
                    float l = m[i][j - 1] + (float)sm.get(read[j], '-');
                    float u = m[i - 1][j] + (float)sm.get(lnz[i], '-');
                    float d = m[i - 1][j - 1] + (float)sm.get(lnz[i], read[j]);
                    m[i][j] = std::max(std::max(l, u), d);
                    if (m[i][j] == d)
                        path[i][j] = (float)(i - 1) + 0.1f;
                    else if (m[i][j] == u)
                        path[i][j] = (float)(i - 1) + 0.2f;
                    else
                        path[i][j] = (float)i + 0.3f;
                } else {
                    float u = 0, d = 0;
                    size_t u_pred = 0, d_pred = 0;
                    bool first = true;
                    for (size_t p : graph.preds(i)) {
                        if (first) {
                            u = m[p][j];
                            d = m[p][j - 1];
                            u_pred = p;
                            d_pred = p;
                            first = false;
                        }
                        if (m[p][j] > u) {
                            u = m[p][j];
                            u_pred = p;
                        }
                        if (m[p][j - 1] > d) {
                            d = m[p][j - 1];
                            d_pred = p;
                        }
                    }
                    u += (float)sm.get(lnz[i], '-');
                    d += (float)sm.get(read[j], lnz[i]);
                    float l = m[i][j - 1] + (float)sm.get(read[j], '-');
                    m[i][j] = std::max(std::max(l, u), d);
                    if (m[i][j] == d)
                        path[i][j] = (float)d_pred + 0.1f;
                    else if (m[i][j] == u)
                        path[i][j] = (float)u_pred + 0.2f;
                    else
                        path[i][j] = (float)i + 0.3f;
                }
                if (m[i][j] >= m[i][best_col]) best_col = j;
            }
        }
        best_scoring_pos[i] = best_col;
    }
    float best_result = 0;
    bool first = true;
    size_t last_row = 0;
    for (size_t p : graph.preds(n - 1)) {
        if (first) {
            best_result = m[p][L - 1];
            last_row = p;
            first = false;
        }
        if (m[p][L - 1] > best_result) {
            best_result = m[p][L - 1];
            last_row = p;
        }
    }
    res.score = (int)best_result;
    if (number != 0) {
        res.gaf = gaf_of_global_abpoa_simd(path, read, name, last_row, L - 1, amb_mode, hofp, lnz, best_result, out);
        res.has_gaf = true;
    }
    return res;
}

// ------------------------------------------------------------------------------- mode 0, scalar
// global_abpoa.rs:428-476
static bool band_ampl_enough_lin(const PathMat& path, const Ampl& ampl, size_t sequence_len, size_t start_row,
                                 size_t start_col) {
    size_t i = start_row, j = start_col;
    while (dir_from_bitvec(path[i][j]) != 'O') {
        size_t left = ampl[i].first, right = ampl[i].second;
        if (i == 0 || (j == 0 && left == 0)) return true;
        if ((j == 0 && left != 0) || (j == right - left - 1 && right != sequence_len)) return false;
        PathCell curr = path[i][j];
        size_t pred = pred_from_bitvec(curr);
        size_t left_p = ampl[pred].first;
        size_t j_pos = left_p < left ? j + (left - left_p) : j - (left_p - left);
        switch (dir_from_bitvec(curr)) {
            case 'D':
            case 'd':
                j = j_pos - 1;
                i = pred;
                break;
            case 'L': j -= 1; break;
            case 'U':
                i = pred;
                j = j_pos;
                break;
            default: throw RefPanic("explicit panic");
        }
        if (i >= path.size() || j >= path[i].size()) throw RefPanic("index out of bounds in band_ampl_enough");
    }
    return true;
}
// global_abpoa.rs:487-526 / gap_global_abpoa.rs:254-292
static bool get_best_d(const std::vector<size_t>& p_arr, const std::vector<std::vector<int>>& m, const Ampl& ampl,
                       size_t i, size_t j, int& d, size_t& d_idx) {
    bool first = true;
    size_t left_i = ampl[i].first;
    for (size_t p : p_arr) {
        size_t left_p = ampl[p].first;
        if (j + left_i > ampl[p].first && j + left_i <= ampl[p].second) {
            size_t j_pos = left_p < left_i ? j + (left_i - left_p) : j - (left_p - left_i);
            int curr_d = m[p][j_pos - 1];
            if (first) {
                d = curr_d;
                d_idx = p;
                first = false;
            }
            if (curr_d > d) {
                d = curr_d;
                d_idx = p;
            }
        }
    }
    return !first;
}
// global_abpoa.rs:529-566
static bool get_best_u_lin(const std::vector<size_t>& p_arr, const std::vector<std::vector<int>>& m,
                           const Ampl& ampl, size_t i, size_t j, int& u, size_t& u_idx) {
    bool first = true;
    size_t left_i = ampl[i].first;
    for (size_t p : p_arr) {
        size_t left_p = ampl[p].first;
        if (j + left_i >= ampl[p].first && j + left_i < ampl[p].second) {
            size_t j_pos = left_p < left_i ? j + (left_i - left_p) : j - (left_p - left_i);
            int current_u = m[p][j_pos];
            if (first) {
                first = false;
                u = current_u;
                u_idx = p;
            }
            if (current_u > u) {
                u = current_u;
                u_idx = p;
            }
        }
    }
    return !first;
}
static size_t min_pred_or_prev(const LnzGraph& g, size_t i) {
    if (!g.nwp[i]) return i - 1;
    const auto& p = g.preds(i);
    return *std::min_element(p.begin(), p.end());
}

// global_abpoa.rs:260-427
PoaResult global_abpoa_exec(const std::vector<char>& sequence, const std::string& name, size_t number,
                            const LnzGraph& g, const ScoreMatrix& sm, size_t bta, bool amb_mode,
                            const std::vector<std::string>& hofp, std::string& out) {
    const auto& lnz = g.lnz;
    const size_t n = lnz.size(), L = sequence.size();
    PoaResult res;
    std::vector<size_t> r_values = set_r_values(g);
    std::vector<size_t> best_scoring_pos(n, 0);
    std::vector<std::vector<int>> m(n);
    PathMat path(n);
    Ampl ampl(n, {0, 0});
    static const std::vector<size_t> empty;
    for (size_t i = 0; i + 1 < n; i++) {
        const std::vector<size_t>& p_arr0 = g.nwp[i] ? g.preds(i) : empty;
        auto lr = set_ampl_for_row(i, p_arr0, r_values[i], best_scoring_pos, L, bta, false);
        size_t left = lr.first, right = lr.second;
        ampl[i] = lr;
        if (right < left) throw RefPanic("attempt to subtract with overflow (right - left)");
        res.cells += right - left;
        size_t best_val_pos = 0;
        m[i].assign(right - left, 0);
        path[i].assign(right - left, 0);
        std::vector<size_t> prev1{i - 1};
        for (size_t j = 0; j < right - left; j++) {
            if (i == 0 && j == 0) {
                m[i][j] = 0;
                path[i][j] = set_path_cell(0, 'O');
            } else if (i == 0) {
                m[i][j] = m[i][j - 1] + sm.get('-', sequence[j + left]);
                path[i][j] = set_path_cell(i, 'L');
            } else if (j == 0 && left == 0) {
                size_t best_p = min_pred_or_prev(g, i);
                m[i][j] = m[best_p].at(j) + sm.get('-', lnz[i]);
                path[i][j] = set_path_cell(best_p, 'U');
            } else {
                const std::vector<size_t>& p_arr = g.nwp[i] ? g.preds(i) : prev1;
                int l, u, d;
                size_t l_pred, u_pred, d_pred;
                if (j > 0) {  // get_best_l, global_abpoa.rs:478-484
                    l = m[i][j - 1] + sm.get(sequence[j + left], '-');
                    l_pred = i;
                } else {
                    l = sm.get(sequence[j + left], '-') * (int)(i + left + j);
                    l_pred = min_pred_or_prev(g, i);
                }
                int uv = 0;
                if (get_best_u_lin(p_arr, m, ampl, i, j, uv, u_pred)) {
                    u = uv + sm.get(lnz[i], '-');
                } else {
                    u = sm.get(lnz[i], '-') * (int)(i + left + j);
                    u_pred = min_pred_or_prev(g, i);
                }
                int dv = 0;
                if (get_best_d(p_arr, m, ampl, i, j, dv, d_pred)) {
                    d = dv + sm.get(lnz[i], sequence[j + left]);
                } else {
                    d = sm.get(lnz[i], '-') * (int)(i + left);
                    d_pred = min_pred_or_prev(g, i);
                }
                auto bd = get_max_d_u_l(d, u, l);
                char dir = bd.second;
                if (dir == 'D' && sequence[j + left] != lnz[i]) dir = 'd';
                m[i][j] = bd.first;
                switch (dir) {
                    case 'D': path[i][j] = set_path_cell(d_pred, 'D'); break;
                    case 'd': path[i][j] = set_path_cell(d_pred, 'd'); break;
                    case 'U': path[i][j] = set_path_cell(u_pred, 'U'); break;
                    default: path[i][j] = set_path_cell(l_pred, 'L'); break;
                }
            }
            if (m[i][j] >= m[i][best_val_pos]) best_val_pos = j;
        }
        best_scoring_pos[i] = best_val_pos + left;
    }
    size_t last_row = n - 2;
    if (m[last_row].empty()) throw RefPanic("attempt to subtract with overflow (empty last row)");
    size_t last_col = m[last_row].size() - 1;
    for (size_t p : g.preds(n - 1)) {
        if (ampl[p].second == ampl[p].first) throw RefPanic("attempt to subtract with overflow (empty row)");
        size_t tmp_last_col = (ampl[p].second - ampl[p].first) - 1;
        if (m[p][tmp_last_col] > m[last_row][last_col]) {
            last_row = p;
            last_col = tmp_last_col;
        }
    }
    bool check = band_ampl_enough_lin(path, ampl, L, last_row, last_col);
    if (!check) out += "Band length probably too short, maybe try with larger b and f\n";
    res.score = m[last_row][last_col];
    if (number != 0) {
        res.gaf = gaf_of_global_abpoa(path, sequence, name, ampl, last_row, last_col, amb_mode, hofp);
        res.has_gaf = true;
    }
    return res;
}

// ------------------------------------------------------------------------------- mode 1
// local_poa.rs:10-179 (AVX2 semantics)
PoaResult local_poa_exec_simd(const std::vector<char>& read, const std::string& name, size_t number,
                              const LnzGraph& graph, const ScoreMatrix& sm, bool amb_mode,
                              const std::vector<std::string>& hofp, std::string& out) {
    (void)out;
    const size_t L = read.size(), n = graph.lnz.size();
    const auto& lnz = graph.lnz;
    PoaResult res;
    std::vector<std::vector<float>> m(n, std::vector<float>(L, 0.0f)), path(n, std::vector<float>(L, 0.0f));
    if (L < 8 && L % 8 == 0) throw RefPanic("attempt to subtract with overflow");
    size_t max_multiple = (L % 8 != 0) ? (L / 8) * 8 : L - 8;
    size_t best_row = 0, best_col = 0;
    for (size_t i = 1; i + 1 < n; i++) {
        float us_update = (float)sm.get(lnz[i], '-');
        for (size_t j = 1; j < max_multiple + 1; j += 8) {
            float ds_update[8];
            for (size_t k = 0; k < 8; k++) ds_update[k] = (float)sm.get(lnz[i], read.at(j + k));
            for (size_t k = 0; k < 8; k++) {
                size_t c = j + k;
                if (!graph.nwp[i]) {
                    float us = m[i - 1][c] + us_update;
                    float ds = m[i - 1][c - 1] + ds_update[k];
                    bool bc = ds > us;
                    m[i][c] = bc ? ds : us;
                    path[i][c] = (float)(i - 1) + (bc ? 0.1f : 0.2f);
                } else {
                    const auto& preds = graph.preds(i);
                    float best_us = m[preds[0]][c], best_ds = m[preds[0]][c - 1];
                    float pred_best_us = (float)preds[0], pred_best_ds = (float)preds[0];
                    for (size_t q = 1; q < preds.size(); q++) {
                        size_t p = preds[q];
                        float us = m[p][c], ds = m[p][c - 1];
                        if (us > best_us) {
                            best_us = us;
                            pred_best_us = (float)p;
                        }
                        if (ds > best_ds) {
                            best_ds = ds;
                            pred_best_ds = (float)p;
                        }
                    }
                    best_us += us_update;
                    best_ds += ds_update[k];
                    bool bc = best_ds > best_us;
                    m[i][c] = bc ? best_ds : best_us;
                    pred_best_ds = pred_best_ds + 0.1f;
                    pred_best_us = pred_best_us + 0.2f;
                    path[i][c] = bc ? pred_best_ds : pred_best_us;
                }
            }
            for (size_t idx = j; idx < std::min(j + 8, L); idx++) {
                float l = m[i][idx - 1] + (float)sm.get(read[j], '-');
                if (l > m[i][idx]) {
                    m[i][idx] = l;
                    path[i][idx] = (float)i + 0.3f;
                }
                if (m[i][idx] <= 0.0f) {
                    m[i][idx] = 0.0f;
                    path[i][idx] = 0.0f;
                }
                if (m[i][idx] >= m[best_row][best_col]) {
                    best_row = i;
                    best_col = idx;
                }
            }
        }
        for (size_t j = max_multiple + 1; j < L; j++) {
            if (!graph.nwp[i]) {
                float l = m[i][j - 1] + (float)sm.get(read[j], '-');
                float u = m[i - 1][j] + (float)sm.get(lnz[i], '-');
                float d = m[i - 1][j - 1] + (float)sm.get(lnz[i], read[j]);
                m[i][j] = std::max(std::max(l, u), d);
                if (m[i][j] < 0.0f) {
                    m[i][j] = 0.0f;
                    path[i][j] = 0.0f;
                } else if (m[i][j] == d)
                    path[i][j] = (float)(i - 1) + 0.1f;
                else if (m[i][j] == u)
                    path[i][j] = (float)(i - 1) + 0.2f;
                else
                    path[i][j] = (float)i + 0.3f;
            } else {
                float u = 0, d = 0;
                size_t u_pred = 0, d_pred = 0;
                bool first = true;
                for (size_t p : graph.preds(i)) {
                    if (first) {
                        u = m[p][j];
                        d = m[p][j - 1];
                        u_pred = p;
                        d_pred = p;
                        first = false;
                    }
                    if (m[p][j] > u) {
                        u = m[p][j];
                        u_pred = p;
                    }
                    if (m[p][j - 1] > d) {
                        d = m[p][j - 1];
                        d_pred = p;
                    }
                }
                u += (float)sm.get(lnz[i], '-');
                d += (float)sm.get(read[j], lnz[i]);
                float l = m[i][j - 1] + (float)sm.get(read[j], '-');
                m[i][j] = std::max(std::max(l, u), d);  // no clamp in this branch (local_poa.rs:126-163)
                if (m[i][j] == d)
                    path[i][j] = (float)d_pred + 0.1f;
                else if (m[i][j] == u)
                    path[i][j] = (float)u_pred + 0.2f;
                else
                    path[i][j] = (float)i + 0.3f;
            }
            if (m[i][j] >= m[best_row][best_col]) {
                best_row = i;
                best_col = j;
            }
        }
    }
    res.cells = (uint64_t)(n - 2) * (L - 1);
    res.score = (int)m[best_row][best_col];
    if (number != 0) {
        res.gaf = gaf_of_local_poa_simd(path, read, name, best_row, best_col, amb_mode, hofp);
        res.has_gaf = true;
    }
    return res;
}

// local_poa.rs:257-293 / gap_local_poa.rs:131-148 — `first` starts false: the running max starts at (0, row 0).
static std::pair<int, size_t> loc_best(const std::vector<std::vector<int>>& m, const std::vector<size_t>& p_arr,
                                       size_t col) {
    int v = 0;
    size_t idx = 0;
    for (size_t p : p_arr) {
        int cur = m[p][col];
        if (cur > v) {
            v = cur;
            idx = p;
        }
    }
    return {v, idx};
}

// local_poa.rs:181-255
PoaResult local_poa_exec(const std::vector<char>& sequence, const std::string& name, size_t number,
                         const LnzGraph& g, const ScoreMatrix& sm, bool amb_mode,
                         const std::vector<std::string>& hofp, std::string& out) {
    (void)out;
    const auto& lnz = g.lnz;
    const size_t n = lnz.size(), L = sequence.size();
    PoaResult res;
    std::vector<std::vector<int>> m(n, std::vector<int>(L, 0));
    PathMat path(n, std::vector<PathCell>(L, 0));
    size_t best_row = 0, best_col = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        for (size_t j = 0; j < L; j++) {
            if (i == 0 || j == 0) {
                path[i][j] = set_path_cell(0, 'O');
            } else {
                int l = m[i][j - 1] + sm.get(sequence[j], '-');
                size_t l_idx = i;
                int d, u;
                size_t d_idx, u_idx;
                if (!g.nwp[i]) {
                    d = m[i - 1][j - 1] + sm.get(sequence[j], lnz[i]);
                    d_idx = i - 1;
                    u = m[i - 1][j] + sm.get('-', lnz[i]);
                    u_idx = i - 1;
                } else {
                    auto bd = loc_best(m, g.preds(i), j - 1);
                    auto bu = loc_best(m, g.preds(i), j);
                    d = bd.first + sm.get(sequence[j], lnz[i]);
                    d_idx = bd.second;
                    u = bu.first + sm.get('-', lnz[i]);
                    u_idx = bu.second;
                }
                if (d < 0 && l < 0 && u < 0) {
                    m[i][j] = 0;
                    path[i][j] = set_path_cell(0, 'O');
                } else {
                    auto bv = get_max_d_u_l(d, u, l);
                    char dir = bv.second;
                    if (dir == 'D' && lnz[i] != sequence[j]) dir = 'd';
                    m[i][j] = bv.first;
                    if (dir == 'D' || dir == 'd')
                        path[i][j] = set_path_cell(d_idx, dir);
                    else if (dir == 'U')
                        path[i][j] = set_path_cell(u_idx, dir);
                    else
                        path[i][j] = set_path_cell(l_idx, dir);
                }
            }
            if (m[i][j] > m[best_row][best_col]) {
                best_row = i;
                best_col = j;
            }
        }
    }
    res.cells = (uint64_t)(n - 2) * (L - 1);
    res.score = m[best_row][best_col];
    if (number != 0) {
        res.gaf = gaf_of_local_poa(path, sequence, name, best_row, best_col, amb_mode, hofp);
        res.has_gaf = true;
    }
    return res;
}

// ------------------------------------------------------------------------------- mode 2
// gap_global_abpoa.rs:371-455
static bool band_ampl_enough_gap(const PathMat& path, const PathMat& path_x, const PathMat& path_y, size_t start_row,
                                 size_t start_col, const Ampl& ampl, size_t sequence_len) {
    size_t i = start_row, j = start_col;
    auto remap = [&](size_t from_row, size_t to_row, size_t col) {
        size_t left = ampl[from_row].first, left_p = ampl[to_row].first;
        return left_p < left ? col + (left - left_p) : col - (left_p - left);
    };
    while (dir_from_bitvec(path[i][j]) != 'O') {
        size_t left = ampl[i].first, right = ampl[i].second;
        if (i == 0 || (j == 0 && left == 0)) return true;
        if ((j == 0 && left != 0) || (j == right - left - 1 && right != sequence_len)) return false;
        PathCell curr = path[i][j];
        size_t pred = pred_from_bitvec(curr);
        switch (dir_from_bitvec(curr)) {
            case 'D':
            case 'd': {
                size_t jp = remap(i, pred, j);
                j = jp - 1;
                i = pred;
                break;
            }
            case 'L':
                if (dir_from_bitvec(path_x[i][j]) == 'X') {
                    while (dir_from_bitvec(path_x[i][j]) == 'X' && j > 0) j -= 1;
                } else {
                    j -= 1;
                }
                break;
            case 'U':
                if (dir_from_bitvec(path_y[i][j]) == 'Y') {
                    while (dir_from_bitvec(path_y[i][j]) == 'Y') {
                        size_t p = pred_from_bitvec(path_y[i][j]);
                        j = remap(i, p, j);
                        i = p;
                        if (i >= path.size() || j >= path[i].size()) throw RefPanic("index out of bounds in band_ampl_enough");
                    }
                } else {
                    size_t p = pred_from_bitvec(path[i][j]);
                    j = remap(i, p, j);
                    i = p;
                }
                break;
            default: return false;
        }
        if (i >= path.size() || j >= path[i].size()) throw RefPanic("index out of bounds in band_ampl_enough");
    }
    return true;
}

// gap_global_abpoa.rs:11-250
PoaResult gap_global_abpoa_exec(const std::vector<char>& sequence, const std::string& name, size_t number,
                                const LnzGraph& g, const ScoreMatrix& sm, int o, int e, size_t bta, bool amb_mode,
                                const std::vector<std::string>& hofp, std::string& out) {
    const auto& lnz = g.lnz;
    const size_t n = lnz.size(), L = sequence.size();
    PoaResult res;
    std::vector<std::vector<int>> m(n), x(n), y(n);
    PathMat path(n), path_x(n), path_y(n);
    std::vector<size_t> r_values = set_r_values(g);
    std::vector<size_t> best_scoring_pos(n, 0);
    Ampl ampl(n, {0, 0});
    static const std::vector<size_t> empty;
    for (size_t i = 0; i + 1 < n; i++) {
        const std::vector<size_t>& p_arr0 = g.nwp[i] ? g.preds(i) : empty;
        auto lr = set_ampl_for_row(i, p_arr0, r_values[i], best_scoring_pos, L, bta, false);
        size_t left = lr.first, right = lr.second;
        ampl[i] = lr;
        if (right < left) throw RefPanic("attempt to subtract with overflow (right - left)");
        res.cells += right - left;
        size_t best_val_pos = 0;
        size_t w = right - left;
        m[i].assign(w, 0);
        x[i].assign(w, 0);
        y[i].assign(w, 0);
        path[i].assign(w, 0);
        path_x[i].assign(w, 0);
        path_y[i].assign(w, 0);
        std::vector<size_t> prev1{i - 1};
        for (size_t j = 0; j < w; j++) {
            if (i == 0 && j == 0) {
                m[i][j] = 0;
                path[i][j] = set_path_cell(0, 'O');
            } else if (i == 0) {
                y[i][j] = o + e * (int)(j + left);
                m[i][j] = y[i][j];
                path[i][j] = set_path_cell(i, 'L');
            } else if (j == 0 && left == 0) {
                size_t best_p = min_pred_or_prev(g, i);
                x[i][j] = o + e * (int)(best_p + 1);
                m[i][j] = x[i][j];
                path[i][j] = set_path_cell(best_p, 'U');
            } else {
                const std::vector<size_t>& p_arr = g.nwp[i] ? g.preds(i) : prev1;
                // get_best_l (gap_global_abpoa.rs:350-368)
                size_t l_pred;
                if (j > 0) {
                    int l_x = x[i][j - 1], l_m = m[i][j - 1] + o;
                    if (l_x > l_m) {
                        x[i][j] = l_x + e;
                        path_x[i][j] = set_path_cell(i, 'X');
                    } else {
                        x[i][j] = l_m + e;
                    }
                    l_pred = i;
                } else {
                    size_t best_p = min_pred_or_prev(g, i);
                    x[i][j] = 2 * o + e * (int)(best_p + 1) + e * (int)(j + left);
                    l_pred = best_p;
                }
                // get_best_u (gap_global_abpoa.rs:296-346)
                size_t u_pred;
                {
                    int u_m = 0, u_y = 0;
                    size_t u_m_idx = 0, u_y_idx = 0;
                    bool first = true;
                    for (size_t p : p_arr) {
                        size_t left_p = ampl[p].first;
                        if (j + left >= ampl[p].first && j + left < ampl[p].second) {
                            size_t j_pos = left_p < left ? j + (left - left_p) : j - (left_p - left);
                            int cum = m[p][j_pos] + o, cuy = y[p][j_pos];
                            if (first) {
                                first = false;
                                u_m = cum;
                                u_y = cuy;
                                u_y_idx = p;
                                u_m_idx = p;
                            }
                            if (cum > u_m) {
                                u_m = cum;
                                u_m_idx = p;
                            }
                            if (cuy > u_y) {
                                u_y = cuy;
                                u_y_idx = p;
                            }
                        }
                    }
                    if (first) {
                        size_t best_p = min_pred_or_prev(g, i);
                        y[i][j] = 2 * o + e * (int)(best_p + 1) + e * (int)(j + left);
                        u_pred = best_p;
                    } else if (u_y > u_m) {
                        y[i][j] = u_y + e;
                        u_pred = u_y_idx;
                        path_y[i][j] = set_path_cell(u_y_idx, 'Y');
                    } else {
                        y[i][j] = u_m + e;
                        u_pred = u_m_idx;
                    }
                }
                int dv = 0;
                size_t d_idx = 0;
                if (get_best_d(p_arr, m, ampl, i, j, dv, d_idx)) {
                    int d = dv + sm.get(lnz[i], sequence[j + left]);
                    int l = x[i][j], u = y[i][j];
                    if (d < l) {
                        if (l < u) {
                            // gap_global_abpoa.rs:153-157: 'u' is not a valid code -> set_path_cell panics.
                            path[i][j] = set_path_cell(u_pred, u_pred == 0 ? 'u' : 'U');
                            m[i][j] = u;
                        } else {
                            path[i][j] = set_path_cell(l_pred, 'L');
                            m[i][j] = l;
                        }
                    } else if (d < u) {
                        path[i][j] = set_path_cell(u_pred, 'U');
                        m[i][j] = u;
                    } else {
                        path[i][j] = set_path_cell(d_idx, lnz[i] == sequence[j + left] ? 'D' : 'd');
                        m[i][j] = d;
                    }
                } else {
                    int l = x[i][j], u = y[i][j];
                    if (l < u) {
                        path[i][j] = set_path_cell(u_pred, 'U');
                        m[i][j] = u;
                    } else {
                        path[i][j] = set_path_cell(l_pred, 'L');
                        m[i][j] = l;
                    }
                }
            }
            if (m[i][j] >= m[i][best_val_pos]) best_val_pos = j;
        }
        best_scoring_pos[i] = best_val_pos + left;
    }
    size_t last_row = n - 2;
    if (m[last_row].empty()) throw RefPanic("attempt to subtract with overflow (empty last row)");
    size_t last_col = m[last_row].size() - 1;
    for (size_t p : g.preds(n - 1)) {
        if (ampl[p].second == ampl[p].first) throw RefPanic("attempt to subtract with overflow (empty row)");
        size_t tmp_last_col = (ampl[p].second - ampl[p].first) - 1;
        if (m[p][tmp_last_col] > m[last_row][last_col]) {
            last_row = p;
            last_col = tmp_last_col;
        }
    }
    int best_value = m[last_row][last_col];
    bool check = band_ampl_enough_gap(path, path_x, path_y, last_row, last_col, ampl, L);
    if (!check) out += "Band length probably too short, maybe try with larger b and f\n";
    res.score = best_value;
    if (number != 0) {
        res.gaf = gaf_of_gap_abpoa(path, path_x, path_y, sequence, name, ampl, last_row, last_col, amb_mode, hofp);
        res.has_gaf = true;
    }
    return res;
}

// ------------------------------------------------------------------------------- mode 3
// gap_local_poa.rs:8-187
PoaResult gap_local_poa_exec(const std::vector<char>& sequence, const std::string& name, size_t number,
                             const LnzGraph& g, const ScoreMatrix& sm, int o, int e, bool amb_mode,
                             const std::vector<std::string>& hofp, std::string& out) {
    (void)out;
    const auto& lnz = g.lnz;
    const size_t n = lnz.size(), L = sequence.size();
    PoaResult res;
    std::vector<std::vector<int>> m(n, std::vector<int>(L, 0)), x(m), y(m);
    PathMat path(n, std::vector<PathCell>(L, 0)), path_x(path), path_y(path);
    size_t best_row = 0, best_col = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        for (size_t j = 0; j < L; j++) {
            if (i == 0 || j == 0) {
                path[i][j] = set_path_cell(0, 'O');
                path_x[i][j] = set_path_cell(0, 'O');
                path_y[i][j] = set_path_cell(0, 'O');
            } else {
                int l_x = x[i][j - 1] + e, l_m = m[i][j - 1] + o + e;
                size_t l_idx = i;
                int l;
                if (l_x > l_m) {
                    path_x[i][j] = set_path_cell(i, 'X');
                    l = l_x;
                } else {
                    path_x[i][j] = set_path_cell(i, 'M');
                    l = l_m;
                }
                x[i][j] = l;
                int d, u;
                size_t d_idx, u_idx;
                if (!g.nwp[i]) {
                    d = m[i - 1][j - 1] + sm.get(sequence[j], lnz[i]);
                    d_idx = i - 1;
                    int u_y = y[i - 1][j] + e, u_m = m[i - 1][j] + o + e;
                    u_idx = i - 1;
                    if (u_y > u_m) {
                        path_y[i][j] = set_path_cell(u_idx, 'Y');
                        u = u_y;
                    } else {
                        path_y[i][j] = set_path_cell(u_idx, 'M');
                        u = u_m;
                    }
                    y[i][j] = u;
                } else {
                    const auto& p_arr = g.preds(i);
                    auto bd = loc_best(m, p_arr, j - 1);
                    d = bd.first;
                    d_idx = bd.second;
                    // get_best_u (gap_local_poa.rs:150-187), `first` starts false
                    int u_m = 0, u_y = 0;
                    size_t u_m_idx = 0, u_y_idx = 0;
                    for (size_t p : p_arr) {
                        int cum = m[p][j] + o, cuy = y[p][j];
                        if (cum > u_m) {
                            u_m = cum;
                            u_m_idx = p;
                        }
                        if (cuy > u_y) {
                            u_y = cuy;
                            u_y_idx = p;
                        }
                    }
                    bool from_m;
                    if (u_m > u_y) {
                        u = u_m;
                        u_idx = u_m_idx;
                        from_m = true;
                    } else {
                        u = u_y;
                        u_idx = u_y_idx;
                        from_m = false;
                    }
                    d += sm.get(sequence[j], lnz[i]);
                    u += e;
                    y[i][j] = u;
                    path_y[i][j] = set_path_cell(u_idx, from_m ? 'M' : 'Y');
                }
                if (d < 0 && l < 0 && u < 0) {
                    m[i][j] = 0;
                    path[i][j] = set_path_cell(0, 'O');
                } else {
                    auto bv = get_max_d_u_l(d, u, l);
                    char dir = bv.second;
                    if (dir == 'D' && lnz[i] != sequence[j]) dir = 'd';
                    m[i][j] = bv.first;
                    if (dir == 'D' || dir == 'd')
                        path[i][j] = set_path_cell(d_idx, dir);
                    else if (dir == 'U')
                        path[i][j] = set_path_cell(u_idx, dir);
                    else
                        path[i][j] = set_path_cell(l_idx, dir);
                }
            }
            if (m[i][j] > m[best_row][best_col]) {
                best_row = i;
                best_col = j;
            }
        }
    }
    res.cells = (uint64_t)(n - 2) * (L - 1);
    res.score = m[best_row][best_col];
    if (number != 0) {
        res.gaf = gaf_of_gap_local_poa(path, path_x, path_y, sequence, name, best_row, best_col, amb_mode, hofp);
        res.has_gaf = true;
    }
    return res;
}

}  // namespace rgo
