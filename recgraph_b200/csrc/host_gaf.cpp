// GAF text from the device's numeric records + run-length step lists.
// The string conventions are the reference's (gaf_output.rs:70-94 to_string; 96-865 per-mode builders;
// pathwise_alignment_output.rs:140-183,471-556; recombination_output.rs:164-234), but nothing here walks a DP
// matrix: the device already did the traceback, the host only regroups runs by segment and prints.
#include <algorithm>
#include <cstring>

#include "host.h"

namespace rg {

namespace {

struct Chunk {  // part of a run that lies inside one segment
    uint32_t op, row, len;
};

// `-s` retries (main.rs:82-101,150-164,198-214,233-249) print segment ids through a handle map built from the handles in
// REVERSE sorted order (utils.rs:144-165 with amb_mode = true) while the rows stay those of the forward linearisation.
thread_local bool t_rev_handles = false;
inline uint64_t sid(const FlatGraph& g, uint32_t row) {
    if (!t_rev_handles) return g.row_seg_id[row];
    const uint32_t seg = g.row_seg[row];
    return seg == UINT32_MAX ? g.row_seg_id[row] : g.seg_ids[g.n_segments - 1 - seg];
}
inline uint32_t run_op(const rg_run& r) { return r.op_count >> 28; }
inline uint32_t run_count(const rg_run& r) { return r.op_count & 0x0fffffffu; }
inline bool consumes_graph(uint32_t op) { return op != RG_OP_L && op != RG_OP_LPAD; }

// Split runs at segment boundaries (rows of one chunk share hofp / nodes_id_pos).
template <typename F>
void for_each_chunk(const FlatGraph& g, const rg_run* runs, uint32_t n_runs, bool ascending, F&& fn) {
    for (uint32_t k = 0; k < n_runs; k++) {
        uint32_t op = run_op(runs[k]), row = runs[k].row, cnt = run_count(runs[k]);
        if (!consumes_graph(op)) {
            fn(Chunk{op, row, cnt});
            continue;
        }
        while (cnt) {
            uint32_t len;
            uint32_t seg = g.row_seg[row];
            if (seg == UINT32_MAX) {
                len = 1;  // rows 0 / n-1
            } else if (!ascending) {
                uint32_t first = g.seg_first_row[seg];
                len = std::min(cnt, row - first + 1);
            } else {
                uint32_t last = (seg + 1 < g.n_segments ? g.seg_first_row[seg + 1] : g.n - 1) - 1;
                len = std::min(cnt, last - row + 1);
            }
            fn(Chunk{op, row, len});
            cnt -= len;
            row = ascending ? row + len : row - len;
        }
    }
}

inline void append_u64(std::string& s, uint64_t v) {
    char buf[24];
    int k = 24;
    do {
        buf[--k] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    s.append(buf + k, 24 - k);
}

// gaf_output.rs:867-874
inline uint64_t node_start(const FlatGraph& g, uint32_t row) {
    if (row == 0) return 0;
    uint32_t seg = g.row_seg[row];
    if (seg == UINT32_MAX) return 0;
    // the reference walks back while the id STRING is equal; adjacent segments never share an id
    return (uint64_t)row - g.seg_first_row[seg] + 1;
}

void gaf_line(std::string& out, const char* name, uint64_t qlen, uint64_t qs, uint64_t qe, char strand,
              const std::vector<uint64_t>& path, uint64_t plen, uint64_t ps, uint64_t pe, uint64_t residues,
              const std::string& comments) {
    out += name;
    out += '\t';
    append_u64(out, qlen);
    out += '\t';
    append_u64(out, qs);
    out += '\t';
    append_u64(out, qe);
    out += '\t';
    out += strand;
    out += "\t>";
    for (size_t k = 0; k < path.size(); k++) {
        if (k) out += '>';
        append_u64(out, path[k]);
    }
    out += '\t';
    append_u64(out, plen);
    out += '\t';
    append_u64(out, ps);
    out += '\t';
    append_u64(out, pe);
    out += '\t';
    append_u64(out, residues);
    out += "\t*\t*\t";
    out += comments;
    out += '\n';
}

// Per-segment CIGAR modes: gaf_of_gap_abpoa (96-253), gaf_of_global_abpoa (254-381), gaf_of_local_poa (383-500),
// gaf_of_gap_local_poa (502-637), gaf_of_local_poa_simd (639-751).
void format_segment_cigar(const FlatGraph& g, const rg_read_result& r, const rg_run* runs, const char* name,
                          uint32_t read_len, bool amb_mode, bool every_diag_is_residue, std::string& out) {
    std::vector<std::string> groups;  // finished per-handle cigars, in traceback order (reference inserts at 0)
    std::vector<std::pair<char, uint64_t>> cur;  // runs of the current handle, traceback order
    std::vector<uint64_t> handles;               // pushed ids, consecutive duplicates removed
    uint64_t count = 0;
    char count_sym = 0;
    bool have_handle = false;
    uint64_t curr_handle = 0;
    char last_dir = ' ';
    uint64_t path_length = 0, residues = 0;
    auto flush_count = [&]() {  // set_cigar_substring (876-892): prepend => we append and reverse later
        if (count) cur.push_back({count_sym, count});
        count = 0;
    };
    auto close_group = [&]() {
        std::string s;
        for (size_t k = cur.size(); k-- > 0;) {
            append_u64(s, cur[k].second);
            s += cur[k].first;
        }
        groups.push_back(std::move(s));
        cur.clear();
    };
    auto on_cell = [&](uint64_t h, char dir) {
        if (!have_handle || h != curr_handle) {
            flush_count();
            close_group();
        }
        have_handle = true;
        curr_handle = h;
        if (dir != last_dir) flush_count();
        last_dir = dir;
    };
    for_each_chunk(g, runs, r.n_runs, false, [&](const Chunk& c) {
        uint64_t h = sid(g, c.row);
        bool row0 = c.row == 0;  // hofp[0] == "-1": only 'L' steps happen there
        uint64_t hkey = row0 ? UINT64_MAX : h;
        switch (c.op) {
            case RG_OP_D:
            case RG_OP_d:
                on_cell(hkey, 'M');
                count_sym = 'M';
                count += c.len;
                path_length += c.len;
                if (c.op == RG_OP_D || every_diag_is_residue) residues += c.len;
                if (handles.empty() || handles.back() != h) handles.push_back(h);
                break;
            case RG_OP_U:
                on_cell(hkey, 'I');
                count_sym = 'I';
                count += c.len;
                path_length += c.len;
                if (handles.empty() || handles.back() != h) handles.push_back(h);
                break;
            case RG_OP_Y:  // inside a `while path_y == 'Y'` chain: no regrouping (gaf_output.rs:186-200)
                count_sym = 'I';
                count += c.len;
                path_length += c.len;
                if (handles.empty() || handles.back() != h) handles.push_back(h);
                break;
            default:  // L
                on_cell(hkey, 'D');
                count_sym = 'D';
                count += c.len;
                break;
        }
    });
    flush_count();
    close_group();
    // cigars[..len-1].join(","): the reference's vector is in forward order with the initial empty group last
    std::string comments;
    for (size_t k = groups.size(); k-- > 1;) {
        comments += groups[k];
        if (k > 1) comments += ',';
    }
    std::reverse(handles.begin(), handles.end());
    gaf_line(out, name, read_len, r.start_col, r.end_col, amb_mode ? '-' : '+', handles, path_length,
             node_start(g, r.start_row), node_start(g, r.end_row), residues, comments);
}

// pathwise_alignment_output.rs:471-556 over a forward-ordered op sequence given as (symbol,count) pieces.
struct CigarBuilder {
    std::string s;
    char sym = 0;
    uint64_t cnt = 0;
    void add(char c, uint64_t n) {
        if (!n) return;
        if (c == sym)
            cnt += n;
        else {
            flush();
            sym = c;
            cnt = n;
        }
    }
    void flush() {
        if (cnt) {
            append_u64(s, cnt);
            s += sym;
        }
        cnt = 0;
        sym = 0;
    }
};
inline char cigar_sym(uint32_t op) {
    switch (op) {
        case RG_OP_D: return 'M';
        case RG_OP_d: return 'X';
        case RG_OP_U:
        case RG_OP_Y:
        case RG_OP_UPAD: return 'I';
        default: return 'D';
    }
}

// Flat-CIGAR modes: gaf_of_global_abpoa_simd (753-865), build_alignment (pathwise_alignment_output.rs:7-184),
// gaf_output_*_no_rec (recombination_output.rs:239-361,633-782), gaf_output_*_rec (12-237,363-631).
struct FlatParts {
    std::string cigar, path_sequence;
    std::vector<uint64_t> handles;  // forward order, de-duplicated
    uint64_t graph_steps = 0, diag_steps = 0;
};
void flat_parts(const FlatGraph& g, const rg_run* fwd, uint32_t n_fwd, const rg_run* rev, uint32_t n_rev,
                FlatParts& fp, uint64_t* fwd_graph_chars) {
    CigarBuilder cb;
    // forward half arrives in traceback order: walk it backwards
    std::vector<Chunk> chunks;
    for_each_chunk(g, fwd, n_fwd, false, [&](const Chunk& c) { chunks.push_back(c); });
    uint64_t fchars = 0;
    for (size_t k = chunks.size(); k-- > 0;) {
        const Chunk& c = chunks[k];
        cb.add(cigar_sym(c.op), c.len);
        if (consumes_graph(c.op)) {
            uint32_t lo = c.row - c.len + 1;
            for (uint32_t rrow = lo; rrow <= c.row; rrow++) fp.path_sequence += CODE_CHARS[g.lnz[rrow]];
            uint64_t h = sid(g, c.row);
            if (fp.handles.empty() || fp.handles.back() != h) fp.handles.push_back(h);
            fp.graph_steps += c.len;
            fchars += c.len;
            if (c.op == RG_OP_D || c.op == RG_OP_d) fp.diag_steps += c.len;
        }
    }
    if (fwd_graph_chars) *fwd_graph_chars = fchars;
    for_each_chunk(g, rev, n_rev, true, [&](const Chunk& c) {
        cb.add(cigar_sym(c.op), c.len);
        if (consumes_graph(c.op)) {
            for (uint32_t rrow = c.row; rrow < c.row + c.len; rrow++) fp.path_sequence += CODE_CHARS[g.lnz[rrow]];
            uint64_t h = g.row_seg_id[c.row];
            if (fp.handles.empty() || fp.handles.back() != h) fp.handles.push_back(h);
            fp.graph_steps += c.len;
        }
    });
    cb.flush();
    fp.cigar = cb.s;
}

// utils.rs:221-254
void path_len_start_end(const FlatGraph& g, uint32_t start, uint32_t end, uint64_t path_len, uint64_t& o_len,
                        uint64_t& o_start, uint64_t& o_end) {
    const auto& ids = g.row_seg_id;
    uint64_t path_start = 0;
    if (start > 0) {
        uint64_t first = ids[start];
        uint32_t counter = start - 1;
        while (counter > 0 && ids[counter] == first) {
            counter--;
            path_start++;
        }
    }
    uint64_t path_end = path_len > 0 ? path_start + path_len - 1 : 0;
    uint64_t end_offset = 0;
    if (end > 0) {
        uint64_t last = ids[end];
        uint32_t counter = end + 1;
        while (counter < g.n - 1 && ids[counter] == last) {
            counter++;
            end_offset++;
        }
    }
    o_len = path_end + end_offset + 1;
    o_start = path_start;
    o_end = path_end;
}
// utils.rs:256-323
void rec_path_len_start_end(const FlatGraph& g, uint32_t fen, uint32_t rsn, uint32_t start, uint32_t end,
                            uint64_t fwd_len, uint64_t rev_len, uint64_t& o_len, uint64_t& o_start, uint64_t& o_end) {
    const auto& ids = g.row_seg_id;
    auto back = [&](uint32_t row) {
        uint64_t c = 0;
        if (row > 0) {
            uint64_t id = ids[row];
            uint32_t counter = row - 1;
            while (counter > 0 && ids[counter] == id) {
                counter--;
                c++;
            }
        }
        return c;
    };
    auto fwd = [&](uint32_t row) {
        uint64_t c = 0;
        if (row > 0) {
            uint64_t id = ids[row];
            uint32_t counter = row + 1;
            while (counter < g.n - 1 && ids[counter] == id) {
                counter++;
                c++;
            }
        }
        return c;
    };
    uint64_t path_start = back(start);
    uint64_t forw_path_end = fwd_len > 0 ? path_start + fwd_len - 1 : 0;
    uint64_t forw_path_len = forw_path_end + fwd(fen) + 1;
    uint64_t rev_path_start = back(rsn);
    uint64_t rev_path_end = rev_len > 0 ? rev_path_start + rev_len - 1 : 0;
    uint64_t path_end = forw_path_len + rev_path_end;
    uint64_t rev_path_len = rev_path_end + fwd(end) + 1;
    o_len = forw_path_len + rev_path_len;
    o_start = path_start;
    o_end = path_end;
}
// pathwise_alignment_recombination.rs:9-22
uint64_t node_offset(const FlatGraph& g, uint32_t row) {
    uint64_t h = g.row_seg_id[row];
    if (h == 0) return 0;
    uint32_t counter = row;
    uint64_t off = 0;
    while (counter > 0 && g.row_seg_id[counter - 1] == h) {
        counter--;
        off++;
    }
    return off;
}

}  // namespace

void format_gaf(const FlatGraph& g, int mode, const rg_read_result& r, const rg_run* all_runs, const char* name,
                uint32_t read_len, int amb_flags, std::string& out) {
    const rg_run* runs = all_runs + r.run_off;
    const bool amb_mode = (amb_flags & RG_AMB_STRAND) != 0;
    struct RevGuard {
        bool old;
        explicit RevGuard(bool on) : old(t_rev_handles) { t_rev_handles = on; }
        ~RevGuard() { t_rev_handles = old; }
    } rev_guard((amb_flags & RG_AMB_HANDLES) != 0 &&
                (mode <= RG_MODE_GAP_LOCAL || mode == RG_MODE_GLOBAL_SCALAR || mode == RG_MODE_LOCAL_SCALAR));
    switch (mode) {
        case RG_MODE_GAP_GLOBAL:
        case RG_MODE_GLOBAL_SCALAR:
            if (r.status & RG_READ_BAND_WARNING) out += "Band length probably too short, maybe try with larger b and f\n";
            format_segment_cigar(g, r, runs, name, read_len, amb_mode, false, out);
            break;
        case RG_MODE_GAP_LOCAL:
        case RG_MODE_LOCAL_SCALAR: format_segment_cigar(g, r, runs, name, read_len, amb_mode, false, out); break;
        case RG_MODE_LOCAL: format_segment_cigar(g, r, runs, name, read_len, amb_mode, true, out); break;
        case RG_MODE_GLOBAL: {
            if (r.status & RG_READ_BAND_NOT_ENOUGH) {
                out += "band not enough for correct output\n";
                out += "\t0\t0\t0\t \t>0\t0\t0\t0\t0\t\t\t\n";  // GAFStruct::new().to_string()
                break;
            }
            FlatParts fp;
            flat_parts(g, runs, r.n_runs, nullptr, 0, fp, nullptr);
            std::string comments = fp.cigar + ", score: " + f32_display(r.score_f32) + "\t" + fp.path_sequence;
            gaf_line(out, name, read_len, r.start_col, r.end_col, amb_mode ? '-' : '+', fp.handles, fp.graph_steps,
                     node_start(g, r.start_row), node_start(g, r.end_row), fp.diag_steps, comments);
            break;
        }
        case RG_MODE_PATHWISE_GLOBAL:
        case RG_MODE_PATHWISE_SEMIGLOBAL:
        case RG_MODE_REC_GLOBAL:
        case RG_MODE_REC_SEMIGLOBAL: {
            const bool rec = (r.status & RG_READ_RECOMBINATION) != 0;
            FlatParts fp;
            uint64_t fwd_chars = 0;
            flat_parts(g, runs, r.n_runs, runs + r.n_runs, rec ? r.n_runs_rev : 0, fp, &fwd_chars);
            uint64_t plen, ps, pe;
            std::string comments;
            uint32_t start = r.start_row == 0 ? 0 : r.start_row + 1;
            if (!rec) {
                path_len_start_end(g, start, r.end_row, fp.graph_steps, plen, ps, pe);
                comments = fp.cigar + ", best path: " + std::to_string(r.best_path) + ", score: " +
                           std::to_string(r.score) + "\t" + fp.path_sequence;
            } else {
                rec_path_len_start_end(g, r.fen, r.rsn, start, r.rev_end_row, fwd_chars, fp.graph_steps - fwd_chars,
                                       plen, ps, pe);
                uint64_t rec_edge = fwd_chars - 1;  // usize wrap as in a release build
                comments = fp.cigar + ", recombination path " + std::to_string(r.best_path) + " " +
                           std::to_string(r.rev_best_path) + ", nodes " + std::to_string(g.row_seg_id[r.fen]) + "[" +
                           std::to_string(node_offset(g, r.fen)) + "] " + std::to_string(g.row_seg_id[r.rsn]) + "[" +
                           std::to_string(node_offset(g, r.rsn)) + "], score: " + f32_display(r.score_f32) +
                           ", displacement: " + std::to_string(r.displacement) + "\t" + fp.path_sequence + "\t" +
                           std::to_string(rec_edge);
            }
            gaf_line(out, name, read_len, 0, read_len - 1, '+', fp.handles, plen, ps, pe, 0, comments);
            break;
        }
        case RG_MODE_PATHWISE_GAP_GLOBAL:
        case RG_MODE_PATHWISE_GAP_SEMIGLOBAL: {
            // pathwise_alignment_gap.rs:563-573 / pathwise_alignment_gap_semi.rs:434-445: the CIGAR line exec println!s
            // (main.rs:277,286 then prints "Best path sequence ..."; the CLI adds that line). Runs arrive in traceback order;
            // mode 6 drops the LAST element of the reversed list, i.e. the first traceback step (…_output.rs:304-305).
            CigarBuilder cb;
            for (uint32_t k = r.n_runs; k-- > 0;) {
                uint32_t cnt = run_count(runs[k]);
                if (mode == RG_MODE_PATHWISE_GAP_GLOBAL && k == 0 && cnt > 0) cnt--;
                if (cnt) cb.add(cigar_sym(run_op(runs[k])), cnt);
            }
            cb.flush();
            out += cb.s;
            if (mode == RG_MODE_PATHWISE_GAP_SEMIGLOBAL) {
                out += "\t(";
                append_u64(out, r.start_row);
                out += ' ';
                append_u64(out, r.rev_end_row);
                out += ')';
            }
            out += '\n';
            break;
        }
        default: break;
    }
}

}  // namespace rg
