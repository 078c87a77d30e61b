// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp). main.rs:25-329 + args_parser.rs restated, plus a
// C entry layer (ctypes) used by tests/ and bench.py's CPU-baseline legs.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

#include "oracle.hpp"

namespace rgo {

struct Args {  // args_parser.rs:3-147 (same flags, same defaults)
    std::string sequence_path, graph_path, out_file = "standard output";
    int alignment_mode = 0, match_score = 2, mismatch_score = 4, gap_open = 4, gap_extension = 2, base_rec_cost = 4;
    std::string matrix = "none", amb_strand = "false";
    float multi_rec_cost = 0.1f, rec_band_width = 1.0f, extra_f = 0.01f;
    int extra_b = 1;
};

static bool parse_args(const std::vector<std::string>& argv, Args& a, std::string& err) {
    struct Opt {
        char s;
        const char* l;
    };
    static const Opt opts[] = {{'o', "out_file"}, {'m', "aln-mode"}, {'M', "match"}, {'X', "mismatch"},
                               {'t', "matrix"}, {'O', "gap-open"}, {'E', "gap-ext"}, {'r', "multi-rec-cost"},
                               {'R', "base-rec-cost"}, {'B', "rec-band-width"}, {'s', "amb-strand"},
                               {'b', "extra-b"}, {'f', "extra-f"}};
    std::vector<std::string> pos;
    auto set = [&](char s, const std::string& v) -> bool {
        try {
            switch (s) {
                case 'o': a.out_file = v; break;
                case 'm': a.alignment_mode = std::stoi(v); break;
                case 'M': a.match_score = std::stoi(v); break;
                case 'X': a.mismatch_score = std::stoi(v); break;
                case 't': a.matrix = v; break;
                case 'O': a.gap_open = std::stoi(v); break;
                case 'E': a.gap_extension = std::stoi(v); break;
                case 'r': a.multi_rec_cost = std::stof(v); break;
                case 'R': a.base_rec_cost = std::stoi(v); break;
                case 'B': a.rec_band_width = std::stof(v); break;
                case 's':
                    if (v != "true" && v != "false") return false;
                    a.amb_strand = v;
                    break;
                case 'b': a.extra_b = std::stoi(v); break;
                case 'f': a.extra_f = std::stof(v); break;
                default: return false;
            }
        } catch (...) {
            return false;
        }
        return true;
    };
    for (size_t k = 1; k < argv.size(); k++) {
        const std::string& s = argv[k];
        if (s.size() >= 2 && s[0] == '-' && s[1] == '-') {
            std::string name = s.substr(2), val;
            bool has_val = false;
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                val = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_val = true;
            }
            char sc = 0;
            for (auto& o : opts)
                if (name == o.l) sc = o.s;
            if (!sc) {
                err += "error: Found argument '" + s + "' which wasn't expected\n";
                return false;
            }
            if (!has_val) {
                if (k + 1 >= argv.size()) {
                    err += "error: The argument '--" + name + "' requires a value\n";
                    return false;
                }
                val = argv[++k];
            }
            if (!set(sc, val)) {
                err += "error: Invalid value for '--" + name + "'\n";
                return false;
            }
        } else if (s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {
            char sc = s[1];
            bool known = false;
            for (auto& o : opts)
                if (o.s == sc) known = true;
            if (!known) {
                err += "error: Found argument '" + s + "' which wasn't expected\n";
                return false;
            }
            std::string val;
            if (s.size() > 2) {
                val = s.substr(s[2] == '=' ? 3 : 2);
            } else {
                if (k + 1 >= argv.size()) {
                    err += std::string("error: The argument '-") + sc + "' requires a value\n";
                    return false;
                }
                val = argv[++k];
            }
            if (!set(sc, val)) {
                err += std::string("error: Invalid value for '-") + sc + "'\n";
                return false;
            }
        } else {
            pos.push_back(s);
        }
    }
    if (pos.size() != 2) {
        err += "error: The following required arguments were not provided:\n    <SEQUENCE_PATH>\n    <GRAPH_PATH>\n";
        return false;
    }
    a.sequence_path = pos[0];
    a.graph_path = pos[1];
    return true;
}

static std::string read_file(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    if (!f) throw RefPanic("No such file or directory: " + p);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// utils.rs:200-219
static void write_gaf(const Args& a, const std::string& gaf_out, size_t number, std::string& out) {
    if (a.out_file == "standard output") {
        out += gaf_out;
        out += "\n";
    } else {
        struct stat st;
        bool exists = stat(a.out_file.c_str(), &st) == 0;
        FILE* f = fopen(a.out_file.c_str(), (exists && number != 1) ? "ab" : "wb");
        if (!f) throw RefPanic("unable to create file");
        fwrite(gaf_out.data(), 1, gaf_out.size(), f);
        fputc('\n', f);
        fclose(f);
    }
}

int recgraph_main(const std::vector<std::string>& argv, std::string& out, std::string& err) {
    auto t0 = std::chrono::steady_clock::now();
    Args a;
    if (!parse_args(argv, a, err)) return 2;
    try {
        std::vector<std::vector<char>> sequences;
        std::vector<std::string> seq_names;
        get_sequences_text(read_file(a.sequence_path), sequences, seq_names);
        HashGraph hg = parse_gfa_file(a.graph_path);
        LnzGraph graph_struct = create_graph_struct(hg, false);
        // score_matrix.rs:21-34
        ScoreMatrix score_matrix;
        if (a.matrix == "HOXD70.mtx" || a.matrix == "HOXD70" || a.matrix == "HOXD55.mtx" || a.matrix == "HOXD55")
            score_matrix = create_score_matrix_hoxd(a.matrix);
        else if (a.matrix == "none")
            score_matrix = create_score_matrix_match_mis(a.match_score, -a.mismatch_score);
        else
            throw RefPanic("wrong matrix type");
        const int align_mode = a.alignment_mode;
        const bool amb_strand = a.amb_strand == "true";
        const float b = (float)a.extra_b, f = a.extra_f;
        const int g_open = -a.gap_open, g_ext = -a.gap_extension;
        std::vector<std::string> hofp_forward = handle_pos_in_lnz(graph_struct, hg, false);
        std::vector<std::string> hofp_reverse;
        auto ensure_rev = [&]() {
            if (hofp_reverse.empty()) hofp_reverse = handle_pos_in_lnz(graph_struct, hg, true);
        };
        switch (align_mode) {
            case 0: {
                std::vector<size_t> r_values = set_r_values(graph_struct);
                for (size_t i = 0; i < sequences.size(); i++) {
                    const auto& seq = sequences[i];
                    size_t bta = bases_to_add(b, f, seq.size());
                    // x86-64 reference hosts have AVX2: exec_simd (main.rs:58-70)
                    PoaResult al = global_abpoa_exec_simd(seq, seq_names[i], i + 1, graph_struct, score_matrix, bta,
                                                          false, hofp_forward, r_values, out);
                    if (amb_strand && al.score < 0) {
                        ensure_rev();
                        auto rev_seq = rev_and_compl(seq);
                        PoaResult ra = global_abpoa_exec(rev_seq, seq_names[i], i + 1, graph_struct, score_matrix, bta,
                                                         true, hofp_reverse, out);
                        write_gaf(a, (ra.score > al.score ? ra : al).gaf.to_string(), i + 1, out);
                    } else {
                        write_gaf(a, al.gaf.to_string(), i + 1, out);
                    }
                }
                break;
            }
            case 1: {
                for (size_t i = 0; i < sequences.size(); i++) {
                    const auto& seq = sequences[i];
                    PoaResult al = local_poa_exec_simd(seq, seq_names[i], i + 1, graph_struct, score_matrix, false,
                                                       hofp_forward, out);
                    if (amb_strand) {
                        ensure_rev();
                        auto rev_seq = rev_and_compl(seq);
                        PoaResult ra = local_poa_exec_simd(rev_seq, seq_names[i], i + 1, graph_struct, score_matrix,
                                                           true, hofp_reverse, out);
                        write_gaf(a, (al.score < ra.score ? al : ra).gaf.to_string(), i + 1, out);  // main.rs:160-164
                    } else {
                        write_gaf(a, al.gaf.to_string(), i + 1, out);
                    }
                }
                break;
            }
            case 2: {
                for (size_t i = 0; i < sequences.size(); i++) {
                    const auto& seq = sequences[i];
                    size_t bta = bases_to_add(b, f, seq.size());
                    PoaResult al = gap_global_abpoa_exec(seq, seq_names[i], i + 1, graph_struct, score_matrix, g_open,
                                                         g_ext, bta, false, hofp_forward, out);
                    if (amb_strand && al.score < 0) {
                        ensure_rev();
                        auto rev_seq = rev_and_compl(seq);
                        PoaResult ra = gap_global_abpoa_exec(rev_seq, seq_names[i], i + 1, graph_struct, score_matrix,
                                                             g_open, g_ext, bta, true, hofp_reverse, out);
                        write_gaf(a, (ra.score > al.score ? ra : al).gaf.to_string(), i + 1, out);
                    } else {
                        write_gaf(a, al.gaf.to_string(), i + 1, out);
                    }
                }
                break;
            }
            case 3: {
                for (size_t i = 0; i < sequences.size(); i++) {
                    const auto& seq = sequences[i];
                    PoaResult al = gap_local_poa_exec(seq, seq_names[i], i + 1, graph_struct, score_matrix, g_open,
                                                      g_ext, false, hofp_forward, out);
                    if (amb_strand) {
                        ensure_rev();
                        auto rev_seq = rev_and_compl(seq);
                        PoaResult ra = gap_local_poa_exec(rev_seq, seq_names[i], i + 1, graph_struct, score_matrix,
                                                          g_open, g_ext, false, hofp_reverse, out);  // amb_mode=false (main.rs:242)
                        write_gaf(a, (ra.score > al.score ? ra : al).gaf.to_string(), i + 1, out);
                    } else {
                        write_gaf(a, al.gaf.to_string(), i + 1, out);
                    }
                }
                break;
            }
            case 4:
            case 5: {
                PathGraph graph = create_path_graph(hg, false);
                for (size_t i = 0; i < sequences.size(); i++) {
                    GAFStruct gaf = align_mode == 4 ? pathwise_alignment_exec(sequences[i], graph, score_matrix)
                                                    : pathwise_alignment_semiglobal_exec(sequences[i], graph, score_matrix);
                    gaf.query_name = seq_names[i];
                    write_gaf(a, gaf.to_string(), i, out);
                }
                break;
            }
            case 6:
            case 7: {  // main.rs:271-288: exec println!s the CIGAR line, main prints the best path
                PathGraph graph = create_path_graph(hg, false);
                for (size_t i = 0; i < sequences.size(); i++) {
                    std::string line;
                    size_t best_path = align_mode == 6
                                           ? pathwise_alignment_gap_exec(sequences[i], graph, score_matrix, g_open, g_ext, line)
                                           : pathwise_alignment_gap_semi_exec(sequences[i], graph, score_matrix, g_open, g_ext, line);
                    out += line;
                    out += "\n";
                    out += "Best path sequence " + std::to_string(i) + ": " + std::to_string(best_path) + "\n";
                }
                break;
            }
            case 8:
            case 9: {
                PathGraph graph = create_path_graph(hg, false);
                PathGraph rev_graph = create_reverse_path_graph(graph);
                Displacement displ = nodes_displacement_matrix(graph, rev_graph);
                for (size_t i = 0; i < sequences.size(); i++) {
                    GAFStruct gaf = pathwise_alignment_recombination_exec(align_mode, sequences[i], graph, rev_graph,
                                                                          score_matrix, a.base_rec_cost,
                                                                          a.multi_rec_cost, displ, a.rec_band_width);
                    gaf.query_name = seq_names[i];
                    write_gaf(a, gaf.to_string(), i, out);
                }
                break;
            }
            default: throw RefPanic("Alignment mode must be in [0..9]");
        }
    } catch (const RefPanic& e) {
        err += std::string("thread 'main' panicked at '") + e.what() + "'\n";
        return 101;
    } catch (const std::out_of_range& e) {
        err += std::string("thread 'main' panicked at 'index out of bounds: ") + e.what() + "'\n";
        return 101;
    }
    auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
    err += "Done in " + std::to_string(secs) + ".\n";
    return 0;
}

}  // namespace rgo

#ifdef RGO_MAIN
int main(int argc, char** argv) {
    std::vector<std::string> args(argv, argv + argc);
    std::string out, err;
    int rc = rgo::recgraph_main(args, out, err);
    fwrite(out.data(), 1, out.size(), stdout);
    fwrite(err.data(), 1, err.size(), stderr);
    return rc;
}
#endif
