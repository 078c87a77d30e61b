// Device-side data layout shared by the kernels and the C-ABI implementation.
#pragma once
#include <cstdint>

#include "../../include/recgraph_b200.h"

namespace rg {

// rowflags bits
enum : uint8_t { RF_NWP = 1, RF_IS_PRED = 2, RF_F_PRED = 4, RF_SINGLE_PREV = 8 /* nwp row whose only predecessor is i-1 */ };

// prev_slot values: slot of the predecessor list that holds row i-1
enum : unsigned { PREV_ALWAYS_SLOT = 0xFE /* row inside a segment */, PREV_NONE_SLOT = 0xFF };

// Everything the forward pass needs to know about a row, packed for one 128-bit load.
struct RowInfo {
    int32_t r_value;
    uint32_t min_pred;
    uint32_t pred_off;
    uint8_t lnz, flags, min_pred_slot, npred;
};

// Graph arrays resident in HBM (one copy per device; read-only during alignment).
struct DevGraph {
    uint32_t n;
    const uint8_t* lnz;        // n codes
    const uint8_t* rowflags;   // n
    const uint32_t* pred_off;  // n+1
    const uint32_t* pred_idx;
    const uint32_t* min_pred;       // n  (best_p)
    const uint8_t* min_pred_slot;   // n
    const uint8_t* prev_slot;       // n
    const int32_t* r_values;        // n
    const RowInfo* rowinfo;         // n
    uint32_t ring;                  // power of two > max look-back: depth of the predecessor-row ring
    const uint32_t* nwp_ord;        // n: ordinal of the row among segment starts that gather predecessors (RF_NWP && !RF_SINGLE_PREV)
    uint32_t n_gather;              // number of such rows
};

struct DevScoring {
    int32_t sc[6][8];  // [graph/first key][read/second key], padded rows
    int32_t o, e;
    float b, f;
    int32_t fixed_bta;
    int32_t R;    // base recombination cost
    float r;      // displacement multiplier
    float rbw;    // recombination band width
};

// Per-row record kept for every row of the read in flight: trace base (offset of column 0 of the row inside the
// slot's trace region, i.e. row offset - left), band [left,right) and the row's best scoring column.
struct RowMeta {
    int32_t base;
    uint32_t left, right, bsp;
};

// Work-space of one slot (= one warp): everything a read in flight needs. Reused read after read.
struct PoaWorkspace {
    RowMeta* rowmeta;    // slots * n
    int32_t* ring_m;     // slots * ring * wstride
    int32_t* ring_y;     // slots * ring * wstride   (affine modes)
    uint8_t* trace;      // slots * trace_cap * trace_bytes
    rg_run* runs;        // slots * run_cap
    uint64_t trace_cap;  // cells per slot
    uint32_t run_cap;    // runs per slot
    uint32_t wstride;    // ints per ring row (>= max L, multiple of 32)
    uint32_t slots;
    uint32_t use16;      // mode 2: allow the packed 16-bit fast path (RG_NO_S16 disables it for A/B tests)
    uint64_t side_off;   // blocked mode-2 kernel: byte offset of the predecessor-slot planes inside a slot's trace region
};

struct PoaBatch {
    const uint8_t* reads;      // concatenated codes
    const uint64_t* read_off;  // n_reads+1
    int32_t n_reads;
    const int32_t* order;      // processing order (longest first) or nullptr
    rg_read_result* results;   // n_reads
    rg_run* out_runs;          // global run output
    uint64_t out_run_cap;
    unsigned long long* counters;  // [0] next read, [1] runs used
};

// ---------------------------------------------------------------------------------------------------------------
// Pathwise family (modes 4/5/8/9). Rows are partitioned into GROUPS, one per incoming edge:
// members = paths(row) & paths(edge), leader = alphas[pred] if a member, else alphas[row] if a member, else the
// lowest member (pathwise_alignment_semiglobal.rs:91-152). Built once per graph on the host.
struct PwGroup {
    uint32_t pred;    // predecessor row (forward graph) / successor row (reverse graph)
    uint32_t leader;  // path id
    uint32_t lead_is_alpha_of_pred;  // leader == alphas[pred]: its row is available in compact form
    uint32_t pad;
};
struct DevPathGraph {
    uint32_t n, P, PW;
    const uint8_t* lnz;
    const uint32_t* alphas;      // n
    const uint32_t* node_bits;   // n * PW
    const uint32_t* grp_off;     // n + 1
    const PwGroup* grp;          // groups of all rows
    const uint32_t* grp_mask;    // PW words per group: members
    const uint8_t* nwp;          // PathGraph nwp of this direction
    const uint64_t* seg;         // nodes_id_pos: segment id of each row (0 for rows 0 and n-1)
    const int32_t* dfs;          // distance from start / from end (pathwise_graph.rs:306-354)
    const int32_t* dfe;
    uint32_t ring;               // rows kept (pow2 > look-back)
    uint32_t max_groups;         // max groups per row
    // ---- score-transport kernel (pathwise_tr.cu): per-row record, table ring depth, highest non-member path per row
    const struct PwtRow* rows;   // n
    const int32_t* nonmem_hi;    // n: highest path id that does NOT go through the row (-1: every path does)
    uint32_t TR;                 // tables alive at once (power of two)
    uint32_t n_groups;           // groups of all rows (size of the move trace in rows)
};
// Row record of the score-transport kernel. A row is a TRANSPORT row when it has exactly one group (then the group holds
// every path of the row): all its paths apply the leader's move, so their scores relative to each other are copied from
// the source cell and only the leader's DP plus a per-column ORIGIN index is computed. Any other row (several incoming
// edges) MATERIALISES the scores of all its paths into a new table.
struct PwtRow {
    uint32_t pred;     // predecessor row of the single group (transport rows)
    uint32_t g0;       // first group of the row (== grp_off[row])
    uint32_t tid;      // id of the table the row's frame refers to (materialising rows: the table they create)
    uint8_t leader;    // leader path of the single group (P <= 128)
    uint8_t nmh;       // highest path id that does NOT go through the row, 255: every path does
    uint8_t lnz;
    uint8_t kind;      // PWT_* bits
};
enum : uint8_t {
    PWT_T = 1,          // transport row
    PWT_RING = 2,       // some later row other than the next one reads this row's frame: keep it in the global ring
    PWT_MXREBUILD = 4,  // modes 8/9: (table, path set) differs from the previous row's: rebuild the per-origin maxima
    PWT_FPRED = 8       // predecessor of the end row (mode 4 results)
};
struct PwWorkspace {
    int32_t* S;         // slots * ring * Lp * Pp     (absolute scores, [row][col][path])
    int32_t* lead;      // slots * ring * Lp          (score of the row's alpha path, compact)
    uint32_t* trace;    // slots * n * Lp * PW * 2    (2-bit own-argmax code per path-cell, as two bit planes)
    rg_run* runs;       // slots * run_cap
    uint32_t Lp, Pp;    // padded columns / paths
    uint32_t run_cap;
    uint32_t slots;
};
// extra buffers of modes 8/9: the reverse pass and the per-(row, column) arg-max tables of best_alignment
struct PwRecWorkspace {
    int32_t* S;        // slots * ring_rev * Lp * Pp
    int32_t* lead;     // slots * ring_rev * Lp
    uint32_t* trace;   // slots * n * Lp * PW * 2
    int2* fm;          // slots * n * Lp : forward  {max over all path slots, path | member << 31}
    int2* rw;          // slots * n * Lp : reverse, column index L-1-j
    int32_t* lastcol;  // slots * n * Pp : forward scores of the last column
};
// Work-space of the score-transport kernel, per slot (= CTA = read in flight).
struct PwtWorkspace {
    int32_t* tables;    // slots * TRmax * Pp * LT      absolute scores of materialised rows, [table][path][column]
    int32_t* ring_lead; // slots * ringmax * LP         frames of rows kept for later segment starts
    uint16_t* ring_org; // slots * ringmax * LP
    uint4* ring_meta;   // slots * ringmax              {leader path of the frame, table id, -, -}
    uint8_t* mv_f;      // slots * groups_f * LP/4      leader moves, 2 bit per (group, column)
    uint8_t* mv_r;      // slots * groups_r * LP/4      (modes 8/9)
    uint8_t* own;       // slots * n * LP/4             own arg-max codes of the replayed path (shared by both directions)
    uint32_t* own_pred; // slots * n                    predecessor row of the replayed path per row
    int2* cb_f;         // slots * n * LP               modes 8/9: per (row, column) {max over all slots, path | member << 31}
    int2* cb_r;
    int32_t* lastcol;   // slots * n * Pp
    int32_t* colmax;    // slots * 2 * LP           modes 8/9: column maxima of cb_f / cb_r
    rg_run* runs;       // slots * run_cap
    uint32_t LP, LT, Pp, CPT;   // columns (= 256 * CPT), table row stride (LP + 32), padded paths, columns per thread
    uint32_t TRmax, ringmax;
    uint32_t run_cap, slots;
    int32_t* stage;     // wide instance only: slots * 2 * LP staging rows of the materialising rows (else nullptr: shared memory)
    uint32_t NT;        // threads per CTA of the instance that runs (256 / 384)
    uint32_t diag;      // RG_PW_DIAG: per-read phase timings (kilo-cycles) overwrite result fields — profiling only
};
int pathwise_tr_cpt(uint32_t Lmax);   // columns per thread for reads of up to Lmax columns; 0 = too long
int pathwise_tr_blocks_per_sm(const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwtWorkspace& ws, bool rec, int* nb);
int launch_pathwise_tr(int mode, const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwtWorkspace& ws,
                       const PoaBatch& b, int blocks, void* stream);
// the same kernel with 384 threads per CTA (pathwise_tr_wide.cu): reads of up to 12 287 bases
int pathwise_tr_cpt_wide(uint32_t Lmax);
int pathwise_tr_blocks_per_sm_wide(const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwtWorkspace& ws, bool rec, int* nb);
int launch_pathwise_tr_wide(int mode, const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwtWorkspace& ws,
                            const PoaBatch& b, int blocks, void* stream);
// Modes 6 / 7 (pathwise_gap.cu): the reference's three delta-encoded tensors per read in flight.
struct PwGapWorkspace {
    int32_t* T;      // slots * 3 * n * Lp * Pp   (dpm, x, y)
    rg_run* runs;    // slots * run_cap
    uint32_t Lp, Pp, run_cap, slots;
};
int launch_pathwise_gap(int mode, const DevPathGraph& g, const DevScoring& s, const PwGapWorkspace& ws, const PoaBatch& b, int blocks,
                        void* stream);
constexpr int REC_SURV = 2048;  // forward nodes of one column staged for the pair expansion
int pathwise_blocks_per_sm(const DevPathGraph& g, const DevPathGraph& rg_, const PwWorkspace& ws, bool rec, int* nb);
int launch_pathwise(int mode, const DevPathGraph& g, const DevPathGraph& rg_, const DevScoring& s, const PwWorkspace& ws,
                    const PwRecWorkspace& rw, const PoaBatch& b, int blocks, void* stream);

// launchers (poa_kernels.cu)
int poa_launch_config(int mode, int trace_bytes, uint32_t Lmax, int* ws_cols, int* blocks_per_sm);
int launch_poa(int mode, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
               int trace_bytes, int blocks, int ws_cols, void* stream);
// register-blocked mode-2 kernel (poa_gap_blk.cu)
int gap_blk_cols(uint32_t Lmax);
int gap_blk_blocks_per_sm(int C, int trace_bytes, int* nb);
int gap_blk_warps_per_block();
int launch_gap_global_blk(int C, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
                          int trace_bytes, int blocks, void* stream);
// modes 0 / 1 / 3 (poa_lin.cu)
int poa_lin_trace_bytes(uint32_t max_indeg);
int launch_poa_lin(int mode, int C, const DevGraph& g, const DevScoring& s, const PoaWorkspace& ws, const PoaBatch& b,
                   int trace_bytes, int blocks, void* stream);
int launch_int_peak(double* iadd, double* imnmx, double* viaddmnmx, void* stream);

}  // namespace rg
