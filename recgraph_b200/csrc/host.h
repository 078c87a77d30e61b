// Host side of recgraph_b200: GFA/FASTA ingestion, graph flattening into device-ready arrays, GAF text.
// Mirrors the reference's host modules (graph.rs, pathwise_graph.rs, sequences.rs, score_matrix.rs,
// gaf_output.rs string work) but is laid out for the device: everything is flat, index-based and CSR.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/recgraph_b200.h"

namespace rg {

enum : uint8_t { CODE_A = 0, CODE_C = 1, CODE_G = 2, CODE_T = 3, CODE_N = 4, CODE_GAP = 5, CODE_START = 6, CODE_END = 7 };
static const char CODE_CHARS[9] = "ACGTN-$F";

// prev_slot special values
enum : uint8_t { PREV_ALWAYS = 0xFE, PREV_NONE = 0xFF };

struct GfaGraph {  // parsed GFA1 (graph.rs:11-17: GFAParser + HashGraph::from_gfa)
    std::vector<uint64_t> seg_id;                  // ascending (graph.rs:32-33 sorts handles)
    std::vector<std::string> seg_seq;              // by segment index
    std::vector<std::vector<uint32_t>> seg_preds;  // left neighbours by segment index, L-line order, de-duplicated
    std::vector<std::vector<uint32_t>> paths;      // steps as segment indices, P-line order (path id = index)
    std::vector<std::string> path_names;
};

struct FlatGraph {
    // ---- LnzGraph (graph.rs:23-102) as flat arrays; n rows incl. '$' (row 0) and 'F' (row n-1)
    uint32_t n = 0;
    std::vector<uint8_t> lnz;        // codes
    std::vector<uint8_t> nwp;        // 0/1
    std::vector<uint32_t> pred_off;  // n+1
    std::vector<uint32_t> pred_idx;  // list order preserved (tie-breaking depends on it)
    std::vector<uint32_t> min_pred;  // "best_p": i-1 for non-nwp rows, min over preds otherwise (global_abpoa.rs:42)
    std::vector<uint8_t> min_pred_slot;  // position of min_pred inside the pred list
    std::vector<uint8_t> prev_slot;      // slot whose pred == i-1, PREV_NONE, or PREV_ALWAYS for non-nwp rows
    std::vector<int32_t> r_values;       // utils.rs:103-126 (-1 == usize::MAX)
    std::vector<uint8_t> is_pred_row;    // row is somebody's predecessor through pred_idx
    std::vector<uint64_t> row_seg_id;    // hofp as integers: segment id of each row (0 for rows 0 and n-1)
    std::vector<uint32_t> row_seg;       // dense segment index per row (UINT32_MAX for rows 0, n-1)
    std::vector<uint64_t> seg_ids;       // id of each dense segment (row order); `-s` retries print them in reverse
                                         // order (utils.rs:144-165 with amb_mode: handles sorted, reversed)
    std::vector<uint32_t> seg_first_row; // per dense segment
    uint32_t n_segments = 0;
    uint32_t max_indeg = 1;
    uint32_t max_lookback = 1;  // max over rows of (i - min pred(i)); ring depth for predecessor rows
    bool has_paths = false;

    // ---- PathGraph (pathwise_graph.rs:10-248); predecessors come from the paths, not from L lines
    uint32_t P = 0, PW = 0;               // paths, 32-bit words per bitset
    std::vector<uint8_t> pw_nwp;          // nwp of the PathGraph (segment starts on >= 1 path, and F)
    std::vector<uint32_t> pw_pred_off, pw_pred_idx;  // per row, ascending pred (fixed order)
    std::vector<uint32_t> pw_edge_bits;   // PW words per pred entry: paths using that edge
    std::vector<uint32_t> node_bits;      // PW words per row: paths through the row
    std::vector<uint32_t> alphas;         // lowest path id through the row (P+1 when none)
    // reverse graph (pathwise_graph.rs:250-282)
    std::vector<uint8_t> rv_nwp;
    std::vector<uint32_t> rv_pred_off, rv_pred_idx, rv_edge_bits;
    std::vector<int32_t> dfs, dfe;        // pathwise_graph.rs:306-354
    uint32_t pw_max_lookback = 1, rv_max_lookback = 1;
};

// Parse GFA1 text. Returns false and fills err on malformed input (the reference unwrap()-panics).
bool parse_gfa(const char* text, size_t len, GfaGraph& g, std::string& err);
// Flatten. Returns an rg_status.
int flatten_graph(const GfaGraph& g, FlatGraph& f, std::string& err);
int flat_from_lnz(uint32_t n, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                  const uint32_t* pred_idx, const uint64_t* seg_id, FlatGraph& f, std::string& err);

int flat_from_path_graph(uint32_t n, uint32_t P, const uint8_t* lnz_codes, const uint8_t* nwp, const uint32_t* pred_off,
                         const uint32_t* pred_idx, const uint32_t* edge_bits, const uint32_t* node_bits, const uint32_t* alphas,
                         const uint64_t* seg_id, FlatGraph& f, std::string& err);

// score_matrix.rs builders on the 6x6 code table.
int make_score_matrix(int kind, int32_t match, int32_t mismatch, rg_scoring* s);
// (b + f * L) as usize in f32 arithmetic, saturating (main.rs:57,175). L includes the '$'.
uint32_t bases_to_add(float b, float f, uint32_t L);

// sequences.rs:5-45
bool parse_fasta(const char* text, size_t len, std::vector<std::string>& names, std::vector<uint8_t>& codes,
                 std::vector<uint64_t>& off, std::string& err, int* status);

std::string f32_display(float v);  // Rust `{}` for f32

// GAF text for one read from its numeric record + runs (gaf_output.rs / pathwise_alignment_output.rs /
// recombination_output.rs string work). Appends everything the reference prints to stdout for the read.
void format_gaf(const FlatGraph& g, int mode, const rg_read_result& r, const rg_run* runs, const char* name,
                uint32_t read_len, int amb_flags, std::string& out);  // amb_flags: RG_AMB_STRAND | RG_AMB_HANDLES

}  // namespace rg
