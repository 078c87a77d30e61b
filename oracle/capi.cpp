// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). C entry points for ctypes (tests/, bench.py CPU legs).
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "oracle.hpp"

using namespace rgo;

static char* dup(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

static LnzGraph lnz_from_arrays(int n, const char* lnz, const uint8_t* nwp, const uint32_t* pred_off,
                                const uint32_t* pred_idx) {
    LnzGraph g;
    g.lnz.assign(lnz, lnz + n);
    g.nwp.assign(nwp, nwp + n);
    g.pred_hash.assign(n, {});
    for (int i = 0; i < n; i++)
        for (uint32_t k = pred_off[i]; k < pred_off[i + 1]; k++) g.pred_hash[i].push_back(pred_idx[k]);
    return g;
}

extern "C" {

void rgo_free(char* p) { free(p); }

// Run the restated CLI in-process. Returns the exit code; *out / *err are malloc'd (rgo_free).
int rgo_main(int argc, const char** argv, char** out, char** err) {
    std::vector<std::string> args(argv, argv + argc);
    std::string o, e;
    int rc;
    try {
        rc = recgraph_main(args, o, e);
    } catch (const std::exception& ex) {
        e += std::string("oracle exception: ") + ex.what() + "\n";
        rc = 101;
    }
    *out = dup(o);
    *err = dup(e);
    return rc;
}

// Score-only POA on a hand-built LnzGraph (the shape of the reference's inline unit tests, e.g.
// global_abpoa.rs:577-754, gap_global_abpoa.rs:465-756, local_poa.rs:304-377, gap_local_poa.rs:198-277).
// variant: 0 = global_abpoa::exec (scalar), 1 = local_poa::exec, 2 = gap_global_abpoa::exec,
//          3 = gap_local_poa::exec, 10 = global_abpoa::exec_simd, 11 = local_poa::exec_simd.
// keys/vals: score matrix entries, keys as 2 chars per entry. Returns 0 ok, 101 on a reference panic.
int rgo_poa_score(int variant, int n, const char* lnz, const uint8_t* nwp, const uint32_t* pred_off,
                  const uint32_t* pred_idx, int read_len, const char* read, int n_scores, const char* keys,
                  const int* vals, int o, int e, int bta, int* score_out, uint64_t* cells_out) {
    try {
        LnzGraph g = lnz_from_arrays(n, lnz, nwp, pred_off, pred_idx);
        ScoreMatrix sm;
        for (int k = 0; k < n_scores; k++) sm.insert(keys[2 * k], keys[2 * k + 1], vals[k]);
        std::vector<char> seq(read, read + read_len);
        std::vector<std::string> hofp;
        std::string out;
        PoaResult r;
        switch (variant) {
            case 0: r = global_abpoa_exec(seq, "test", 0, g, sm, (size_t)bta, false, hofp, out); break;
            case 1: r = local_poa_exec(seq, "test", 0, g, sm, false, hofp, out); break;
            case 2: r = gap_global_abpoa_exec(seq, "test", 0, g, sm, o, e, (size_t)bta, false, hofp, out); break;
            case 3: r = gap_local_poa_exec(seq, "test", 0, g, sm, o, e, false, hofp, out); break;
            case 10: {
                auto rv = set_r_values(g);
                r = global_abpoa_exec_simd(seq, "test", 0, g, sm, (size_t)bta, false, hofp, rv, out);
                break;
            }
            case 11: r = local_poa_exec_simd(seq, "test", 0, g, sm, false, hofp, out); break;
            default: return 2;
        }
        *score_out = r.score;
        if (cells_out) *cells_out = r.cells;
        return 0;
    } catch (const std::exception&) {
        return 101;
    }
}

// Dump the LnzGraph built from GFA text (graph.rs:31-102) as text:
//   lnz=<chars>\n nwp=<0/1 string>\n pred <i>: a b c\n ... hofp <i>: id
char* rgo_dump_lnz(const char* gfa_text, int amb_mode) {
    try {
        HashGraph hg = parse_gfa_text(gfa_text);
        LnzGraph g = create_graph_struct(hg, amb_mode != 0);
        std::ostringstream os;
        os << "lnz=" << std::string(g.lnz.begin(), g.lnz.end()) << "\n";
        os << "nwp=";
        for (size_t i = 0; i < g.lnz.size(); i++) os << (g.nwp[i] ? '1' : '0');
        os << "\n";
        for (size_t i = 0; i < g.lnz.size(); i++)
            if (!g.pred_hash[i].empty()) {
                os << "pred " << i << ":";
                for (size_t p : g.pred_hash[i]) os << " " << p;
                os << "\n";
            }
        auto hofp = handle_pos_in_lnz(g, hg, amb_mode != 0);
        for (size_t i = 0; i + 1 < g.lnz.size(); i++) os << "hofp " << i << ": " << hofp[i] << "\n";
        auto rv = set_r_values(g);
        os << "r_values=";
        for (size_t i = 0; i < rv.size(); i++) os << (i ? "," : "") << (long)rv[i];
        os << "\n";
        return dup(os.str());
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}

// Dump the PathGraph (pathwise_graph.rs:135-248), optionally its reverse graph (250-282) and dfs/dfe (306-354).
char* rgo_dump_pathgraph(const char* gfa_text, int is_reversed, int reverse_graph) {
    try {
        HashGraph hg = parse_gfa_text(gfa_text);
        PathGraph g = create_path_graph(hg, is_reversed != 0);
        PathGraph rg = create_reverse_path_graph(g);
        const PathGraph& d = reverse_graph ? rg : g;
        std::ostringstream os;
        os << "paths_number=" << d.paths_number << "\n";
        os << "lnz=" << std::string(d.lnz.begin(), d.lnz.end()) << "\n";
        os << "nwp=";
        for (size_t i = 0; i < d.lnz.size(); i++) os << (d.nwp[i] ? '1' : '0');
        os << "\n";
        for (size_t i = 0; i < d.lnz.size(); i++) {
            os << "node " << i << ": id=" << d.nodes_id_pos[i] << " alpha=" << d.alphas[i] << " paths=";
            for (size_t k = 0; k < d.paths_number; k++) os << (d.paths_nodes[i][k] ? '1' : '0');
            os << "\n";
            for (auto& pp : d.pred_hash[i]) {
                os << "pred " << i << " " << pp.first << " ";
                for (size_t k = 0; k < d.paths_number; k++) os << (pp.second[k] ? '1' : '0');
                os << "\n";
            }
        }
        Displacement disp = nodes_displacement_matrix(g, rg);
        os << "dfs=";
        for (size_t i = 0; i < disp.dfs.size(); i++) os << (i ? "," : "") << disp.dfs[i];
        os << "\ndfe=";
        for (size_t i = 0; i < disp.dfe.size(); i++) os << (i ? "," : "") << disp.dfe[i];
        os << "\n";
        return dup(os.str());
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}

// helpers pinned by the reference's own tests
// Modes 4-7 on one read, for the self-consistency checks of the affine pathwise restatement (with o = 0 the affine
// recurrences of modes 6 / 7 reduce to the linear ones of modes 4 / 5). Returns "score best_path cigar-or-line".
char* rgo_pathwise_one(const char* gfa_text, const char* read, int mode, int m, int x, int o, int e) {
    try {
        HashGraph hg = parse_gfa_text(gfa_text);
        PathGraph g = create_path_graph(hg, false);
        ScoreMatrix sm = create_score_matrix_match_mis(m, x);
        std::vector<char> seq{'$'};
        for (const char* c = read; *c; c++) seq.push_back(*c);
        std::ostringstream os;
        if (mode == 6 || mode == 7) {
            std::string line;
            int score = 0;
            size_t bp = mode == 6 ? pathwise_alignment_gap_exec(seq, g, sm, o, e, line, &score)
                                  : pathwise_alignment_gap_semi_exec(seq, g, sm, o, e, line, &score);
            os << score << " " << bp << " " << line;
        } else {
            GAFStruct gaf = mode == 4 ? pathwise_alignment_exec(seq, g, sm) : pathwise_alignment_semiglobal_exec(seq, g, sm);
            os << gaf.comments;
        }
        return dup(os.str());
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}

// api.rs:11-40 (align_global_no_gap) and 76-99 (align_local_no_gap) with their defaults: f32 match/mismatch matrix whose
// gap entries are X (score_matrix.rs:52-66), bases_to_add = (len * 0.1) as usize. Returns what exec_simd println!s
// followed by the GAF record.
char* rgo_api_no_gap(const char* gfa_text, const char* read, const char* name, int local) {
    try {
        HashGraph hg = parse_gfa_text(gfa_text);
        std::vector<char> seq = build_align_string(read);
        LnzGraph lg = create_graph_struct(hg, false);
        ScoreMatrix sm = create_score_matrix_match_mis_f32(2, -4);
        volatile float prod = (float)std::string(read).size() * 0.1f;
        const size_t bta = (size_t)prod;
        std::vector<size_t> r_values = set_r_values(lg);
        std::vector<std::string> hofp = handle_pos_in_lnz(lg, hg, false);
        std::string out;
        PoaResult r = local ? local_poa_exec_simd(seq, name, 1, lg, sm, false, hofp, out)
                            : global_abpoa_exec_simd(seq, name, 1, lg, sm, bta, false, hofp, r_values, out);
        if (r.has_gaf) out += r.gaf.to_string() + "\n";
        return dup(out);
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}

// local_poa::exec (scalar routine, local_poa.rs:181-255; reached only on hosts without AVX2) with the CLI's default i32
// matrix: stdout of exec followed by the GAF record.
char* rgo_local_scalar_gaf(const char* gfa_text, const char* read, const char* name, int m, int x) {
    try {
        HashGraph hg = parse_gfa_text(gfa_text);
        std::vector<char> seq = build_align_string(read);
        LnzGraph lg = create_graph_struct(hg, false);
        ScoreMatrix sm = create_score_matrix_match_mis(m, x);
        std::vector<std::string> hofp = handle_pos_in_lnz(lg, hg, false);
        std::string out;
        PoaResult r = local_poa_exec(seq, name, 1, lg, sm, false, hofp, out);
        if (r.has_gaf) out += r.gaf.to_string() + "\n";
        return dup(out);
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}

char* rgo_rev_and_compl(const char* seq) {
    try {
        std::vector<char> s(seq, seq + strlen(seq));
        auto r = rev_and_compl(s);
        return dup(std::string(r.begin(), r.end()));
    } catch (const std::exception& ex) {
        return dup(std::string("PANIC ") + ex.what());
    }
}
// which: 0 = create_score_matrix_match_mis(m, x), 1 = HOXD70, 2 = HOXD55, 3 = match_mis_f32(m, x)
int rgo_score_lookup(int which, int m, int x, char a, char b, int* present) {
    ScoreMatrix sm = which == 0   ? create_score_matrix_match_mis(m, x)
                     : which == 1 ? create_score_matrix_hoxd("HOXD70.mtx")
                     : which == 2 ? create_score_matrix_hoxd("HOXD55.mtx")
                                  : create_score_matrix_match_mis_f32(m, x);
    *present = sm.has[(int)a][(int)b];
    return sm.v[(int)a][(int)b];
}
int rgo_bases_to_add(float b, float f, int seq_len) { return (int)bases_to_add(b, f, (size_t)seq_len); }
char* rgo_f32_display(float v) { return dup(f32_display(v)); }

}  // extern "C"
