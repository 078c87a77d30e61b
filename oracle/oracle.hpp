// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of AlgoLab/RecGraph's sequence-to-graph DP aligner (the Rust reference under
// /root/reference, which cannot be compiled in this image: no cargo/rustc, crates not vendored).
// Every function cites the reference file:line it follows. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, link or execute this code.
// The product (recgraph_b200/) never includes or links anything from oracle/.
//
// Parity pin: the POA score recurrences are pinned by the reference's own 15 inline score tests
// and the graph builders by its 11 structure tests (tests/test_oracle_golden.py transcribes them).
// Pathwise/recombination DP, all tracebacks and all GAF text are **parity unpinned** by the
// reference (it ships no test or golden output for them); this restatement is the only pin.
//
// Deliberate, documented deviations (the reference's behaviour is nondeterministic there):
//  * std::collections::HashMap iteration orders (graph.rs:120-122 `last_nodes`,
//    pathwise_graph.rs:86-93 `PredHash`) are fixed to ascending key order.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rgo {

// A Rust panic!/unwrap() failure in the reference.
struct RefPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------
// score_matrix.rs: HashMap<(char,char), i32|f32>. Missing key => unwrap() panic.
struct ScoreMatrix {
    int v[128][128];
    bool has[128][128];
    ScoreMatrix() {
        std::memset(v, 0, sizeof v);
        std::memset(has, 0, sizeof has);
    }
    void insert(char a, char b, int s) {
        v[(int)a][(int)b] = s;
        has[(int)a][(int)b] = true;
    }
    void remove(char a, char b) { has[(int)a][(int)b] = false; }
    inline int get(char a, char b) const {
        if ((unsigned char)a > 127 || (unsigned char)b > 127 || !has[(int)a][(int)b])
            throw RefPanic(std::string("score_matrix.get((") + a + "," + b + ")).unwrap() on None");
        return v[(int)a][(int)b];
    }
};
ScoreMatrix create_score_matrix_match_mis(int m, int x);      // score_matrix.rs:35-51
ScoreMatrix create_score_matrix_match_mis_f32(int m, int x);  // score_matrix.rs:52-66 (values are integral)
ScoreMatrix create_score_matrix_hoxd(const std::string& name);  // score_matrix.rs:67-105, tables embedded

// ---------------------------------------------------------------------------------------------
// handlegraph::HashGraph stand-in (un-vendored crate handlegraph ^0.5.0; semantics per SURVEY §8c).
struct Handle {
    uint64_t packed;  // id << 1 | is_reverse  (handlegraph::handle::Handle)
    uint64_t id() const { return packed >> 1; }
    bool is_reverse() const { return packed & 1; }
    Handle flip() const { return Handle{packed ^ 1}; }
    bool operator<(const Handle& o) const { return packed < o.packed; }
    bool operator==(const Handle& o) const { return packed == o.packed; }
};
struct HGNode {
    std::string sequence;
    std::vector<Handle> left_edges, right_edges;  // insertion (GFA L-line) order
};
struct HGPath {
    std::string name;
    std::vector<Handle> nodes;
};
struct HashGraph {
    std::map<uint64_t, HGNode> graph;
    std::vector<HGPath> paths;  // path id = index (P-line order)
    uint64_t max_id = 0;
    Handle append_handle(const std::string& seq);  // ids 1,2,3,... (MutableHandleGraph)
    Handle create_handle(const std::string& seq, uint64_t id);
    void create_edge(Handle left, Handle right);
    size_t create_path_handle(const std::string& name);
    void append_step(size_t path, Handle h);
    std::string sequence(Handle h) const;  // reverse complement for reversed handles
    std::vector<Handle> left_neighbours(Handle h) const;  // handle_edges_iter(h, Direction::Left)
    std::vector<Handle> handles_sorted(bool amb_mode) const;
};
HashGraph parse_gfa_text(const std::string& text);  // gfa ^0.8.0 GFAParser + HashGraph::from_gfa
HashGraph parse_gfa_file(const std::string& path);

// graph.rs:23-27
struct LnzGraph {
    std::vector<char> lnz;
    std::vector<uint8_t> nwp;
    std::vector<std::vector<size_t>> pred_hash;  // indexed by node; empty when no entry
    bool has_pred(size_t i) const { return !pred_hash[i].empty(); }
    const std::vector<size_t>& preds(size_t i) const {
        if (pred_hash[i].empty()) throw RefPanic("pred_hash.get(&i).unwrap() on None");
        return pred_hash[i];
    }
};
LnzGraph create_graph_struct(const HashGraph& g, bool amb_mode);                    // graph.rs:31-102
std::vector<std::string> handle_pos_in_lnz(const LnzGraph& lg, const HashGraph& g, bool amb_mode);  // utils.rs:144-198

typedef std::vector<uint8_t> BitVec;  // one byte per bit; clarity over speed

// pathwise_graph.rs:10-18, 75-125
struct PathGraph {
    std::vector<char> lnz;
    std::vector<uint8_t> nwp;
    std::vector<std::vector<std::pair<size_t, BitVec>>> pred_hash;  // per node: (pred, edge paths), ascending pred
    std::vector<BitVec> paths_nodes;
    std::vector<size_t> alphas;
    size_t paths_number = 0;
    std::vector<uint64_t> nodes_id_pos;
    const std::vector<std::pair<size_t, BitVec>>& get_preds_and_paths(size_t i) const {
        if (pred_hash[i].empty()) throw RefPanic("PredHash.get(&curr_node).unwrap() on None");
        return pred_hash[i];
    }
};
PathGraph create_path_graph(const HashGraph& g, bool is_reversed);      // pathwise_graph.rs:135-248
PathGraph create_reverse_path_graph(const PathGraph& fwd);              // pathwise_graph.rs:250-282
std::vector<long> get_distance_from_start(const PathGraph& rev_graph);  // pathwise_graph.rs:306-329
std::vector<long> get_distance_from_end(const PathGraph& graph);        // pathwise_graph.rs:330-354
// pathwise_graph.rs:284-305 materialises n x n; the oracle keeps the two vectors and evaluates on demand.
struct Displacement {
    std::vector<long> dfs, dfe;
    int at(size_t i, size_t j) const {
        if (i == j) return 0;
        long a = dfs[i] - dfs[j], b = dfe[i] - dfe[j];
        return (int)((a < 0 ? -a : a) + (b < 0 ? -b : b));
    }
};
Displacement nodes_displacement_matrix(const PathGraph& g, const PathGraph& rev);

// sequences.rs:5-82
void get_sequences_text(const std::string& fasta_text, std::vector<std::vector<char>>& seqs,
                        std::vector<std::string>& names);
std::vector<char> build_align_string(const std::string& line);
std::vector<char> rev_and_compl(const std::vector<char>& seq);

// utils.rs
std::pair<size_t, size_t> set_ampl_for_row(size_t i, const std::vector<size_t>& p_arr, size_t r_val,
                                           const std::vector<size_t>& best_scoring_pos, size_t seq_len,
                                           size_t bta, bool simd_version);  // utils.rs:17-98
std::vector<size_t> set_r_values(const LnzGraph& g);                        // utils.rs:103-126
size_t bases_to_add(float b, float f, size_t seq_len);                     // main.rs:57,175

// gaf_output.rs:6-95
struct GAFStruct {
    std::string query_name;
    size_t query_length = 0, query_start = 0, query_end = 0;
    char strand = ' ';
    std::vector<size_t> path{0};
    size_t path_length = 0, path_start = 0, path_end = 0, residue_matches_number = 0;
    std::string alignment_block_length, mapping_quality, comments;
    std::string to_string() const;
};
std::string f32_display(float v);  // Rust `{}` for f32
std::string build_cigar(const std::vector<char>& cigar);  // pathwise_alignment_output.rs:471-556

struct PoaResult {
    int score = 0;
    bool has_gaf = false;
    GAFStruct gaf;
    uint64_t cells = 0;  // in-band cells computed (GCUPS accounting; not in the reference)
};
// `out` receives what the reference println!s to stdout inside exec/gaf builders.
PoaResult global_abpoa_exec_simd(const std::vector<char>& read, const std::string& name, size_t number,
                                 const LnzGraph& g, const ScoreMatrix& sm, size_t bta, bool amb_mode,
                                 const std::vector<std::string>& hofp, const std::vector<size_t>& r_values,
                                 std::string& out);  // global_abpoa.rs:10-257
PoaResult global_abpoa_exec(const std::vector<char>& seq, const std::string& name, size_t number,
                            const LnzGraph& g, const ScoreMatrix& sm, size_t bta, bool amb_mode,
                            const std::vector<std::string>& hofp, std::string& out);  // global_abpoa.rs:260-427
PoaResult local_poa_exec_simd(const std::vector<char>& read, const std::string& name, size_t number,
                              const LnzGraph& g, const ScoreMatrix& sm, bool amb_mode,
                              const std::vector<std::string>& hofp, std::string& out);  // local_poa.rs:10-179
PoaResult local_poa_exec(const std::vector<char>& seq, const std::string& name, size_t number,
                         const LnzGraph& g, const ScoreMatrix& sm, bool amb_mode,
                         const std::vector<std::string>& hofp, std::string& out);  // local_poa.rs:181-255
PoaResult gap_global_abpoa_exec(const std::vector<char>& seq, const std::string& name, size_t number,
                                const LnzGraph& g, const ScoreMatrix& sm, int o, int e, size_t bta,
                                bool amb_mode, const std::vector<std::string>& hofp,
                                std::string& out);  // gap_global_abpoa.rs:11-250
PoaResult gap_local_poa_exec(const std::vector<char>& seq, const std::string& name, size_t number,
                             const LnzGraph& g, const ScoreMatrix& sm, int o, int e, bool amb_mode,
                             const std::vector<std::string>& hofp, std::string& out);  // gap_local_poa.rs:8-129

// pathwise family
GAFStruct pathwise_alignment_exec(const std::vector<char>& seq, const PathGraph& g,
                                  const ScoreMatrix& sm);  // pathwise_alignment.rs:5-340
GAFStruct pathwise_alignment_semiglobal_exec(const std::vector<char>& seq, const PathGraph& g,
                                             const ScoreMatrix& sm);  // pathwise_alignment_semiglobal.rs:6-242
// Experimental affine pathwise modes 6 / 7: return the best path, `line` = the CIGAR line exec println!s
// (pathwise_alignment_gap.rs:4-574, pathwise_alignment_gap_semi.rs:5-473, pathwise_alignment_output.rs:186-451).
size_t pathwise_alignment_gap_exec(const std::vector<char>& seq, const PathGraph& g, const ScoreMatrix& sm, int o, int e,
                                   std::string& line, int* best_score = nullptr);
size_t pathwise_alignment_gap_semi_exec(const std::vector<char>& seq, const PathGraph& g, const ScoreMatrix& sm, int o, int e,
                                        std::string& line, int* best_score = nullptr);
GAFStruct pathwise_alignment_recombination_exec(int aln_mode, const std::vector<char>& seq, const PathGraph& g,
                                                const PathGraph& rev_g, const ScoreMatrix& sm, int base_rec_cost,
                                                float multi_rec_cost, const Displacement& displ,
                                                float rbw);  // pathwise_alignment_recombination.rs:23-127

// main.rs:25-329 restated; argv[0] ignored. Returns the process exit code; stdout text is appended to `out`,
// stderr text to `err`.
int recgraph_main(const std::vector<std::string>& args, std::string& out, std::string& err);

}  // namespace rgo
