/* recgraph_b200 — C ABI of the B200-native RecGraph aligner (drop-in boundary).
 *
 * The reference (AlgoLab/RecGraph, Rust) exposes no FFI; its two public surfaces are the CLI
 * (/root/reference/src/main.rs:25-329) and the library API (/root/reference/src/api.rs:11-164).
 * Every entry point below names the reference interface it replaces. A Rust `api.rs` / `main.rs`
 * shim binds these with `extern "C"`; this repo's own host side (C++ CLI
 * `recgraph`, Python ctypes harness) calls exactly the same symbols.
 *
 * Conventions: return 0 on success, a negative rg_status otherwise; never throws or aborts across
 * the boundary; all pointers are plain host pointers owned by the caller unless stated; buffers
 * returned inside rg_batch_result are owned by the ctx and stay valid until the next align/fetch call.
 * One ctx = one CUDA device + one stream; calls on one ctx are not re-entrant, different ctxs may
 * be driven from different threads. There is NO CPU fallback: without a CUDA device rg_init fails.
 */
#ifndef RECGRAPH_B200_H
#define RECGRAPH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_ctx rg_ctx;

typedef enum rg_status {
    RG_OK = 0,
    RG_ERR_INVALID = -1,     /* bad argument / call order */
    RG_ERR_CUDA = -2,        /* CUDA runtime error (rg_last_error has the text) */
    RG_ERR_NO_DEVICE = -3,   /* no usable CUDA device: the product never falls back to the CPU */
    RG_ERR_IO = -4,          /* file could not be read / malformed GFA or FASTA (reference: unwrap() panic) */
    RG_ERR_BAD_CHAR = -5,    /* character outside A,C,G,T,N in read or graph (reference: score lookup panic) */
    RG_ERR_UNSUPPORTED = -6, /* input outside the device path's documented domain */
    RG_ERR_REF_PANIC = -7,   /* the reference would panic on this input (message in rg_last_error) */
    RG_ERR_NOMEM = -8
} rg_status;

/* Alphabet codes used on the boundary: A=0 C=1 G=2 T=3 N=4 gap('-')=5. */
enum { RG_A = 0, RG_C = 1, RG_G = 2, RG_T = 3, RG_N = 4, RG_GAP = 5 };

/* Scoring and banding parameters. Replaces score_matrix.rs:21-105 (HashMap<(char,char),i32|f32>),
 * args_parser.rs:148-202 getters, and the Option<> arguments of api.rs:11-128.
 * score[a][b] is the reference's score_matrix.get(&(a,b)) for codes a,b (entry [5][5] unused).
 * o,e are the NEGATIVE gap open / extension as the reference passes them to exec (main.rs:172). */
typedef struct rg_scoring {
    int32_t score[6][6];
    int32_t gap_open;      /* o  (<= 0 for CLI use; any value accepted)               */
    int32_t gap_ext;       /* e                                                        */
    int32_t base_rec_cost; /* -R, pathwise_alignment_recombination.rs:29               */
    float multi_rec_cost;  /* -r                                                       */
    float rec_band_width;  /* -B                                                       */
    float extra_b;         /* -b as f32 (args_parser.rs:180)                           */
    float extra_f;         /* -f                                                       */
    int32_t fixed_bta;     /* >= 0: use this band half-width for every read instead of (b + f*L) — the
                              `bases_to_add` argument of api.rs:16,51 / the inline unit tests; -1 = off */
} rg_scoring;

/* Alignment modes == the reference's `-m` values (main.rs:48-318). Modes 0/1 run the AVX2 semantics
 * (global_abpoa::exec_simd / local_poa::exec_simd) as every x86-64 reference build does; the scalar
 * `exec` variants are RG_MODE_GLOBAL_SCALAR (global_abpoa::exec, global_abpoa.rs:260-427: the `-s` retry of mode 0,
 * main.rs:89-97) and RG_MODE_LOCAL_SCALAR (local_poa::exec, reached only on hosts without AVX2).
 * Device status: every mode runs on the GPU (6 / 7, experimental in the reference, keep its n x L x P tensors per read in
 * flight); inputs outside a kernel's documented domain return RG_ERR_UNSUPPORTED — there is no CPU fallback. The domain:
 * characters A,C,G,T,N; in-degree <= 31 (modes 0/1/3) / 64 (mode 2); <= 128 paths; reads <= 1023 bases in modes 0/1/3 and
 * <= 12 287 bases through the fast pathwise kernel. */
enum {
    RG_MODE_GLOBAL = 0,
    RG_MODE_LOCAL = 1,
    RG_MODE_GAP_GLOBAL = 2,
    RG_MODE_GAP_LOCAL = 3,
    RG_MODE_PATHWISE_GLOBAL = 4,
    RG_MODE_PATHWISE_SEMIGLOBAL = 5,
    RG_MODE_PATHWISE_GAP_GLOBAL = 6,
    RG_MODE_PATHWISE_GAP_SEMIGLOBAL = 7,
    RG_MODE_REC_GLOBAL = 8,
    RG_MODE_REC_SEMIGLOBAL = 9,
    RG_MODE_GLOBAL_SCALAR = 10,
    RG_MODE_LOCAL_SCALAR = 11
};

/* Per-read status bits. */
enum {
    RG_READ_OK = 0,
    RG_READ_BAND_WARNING = 1,   /* reference prints "Band length probably too short, ..." (gap_global_abpoa.rs:226) */
    RG_READ_BAND_NOT_ENOUGH = 2,/* reference prints "band not enough for correct output" + empty GAF (gaf_output.rs:862) */
    RG_READ_REF_PANIC = 4,      /* reference panics on this read (e.g. gap_global_abpoa.rs:153-154 'u' code) */
    RG_READ_TRACE_OVERFLOW = 8, /* device trace/run buffers too small even after retry */
    RG_READ_RECOMBINATION = 16  /* modes 8/9: forward and reverse best paths differ */
};

/* One alignment step run, in TRACEBACK order (last alignment column first). Lossless run-length form of
 * the reference's per-cell traceback loops (gaf_output.rs:96-865, pathwise_alignment_output.rs:7-184,
 * recombination_output.rs:12-782): `count` consecutive steps with the same op starting at lnz row `row`;
 * the row decreases by 1 per step for graph-consuming ops and stays fixed for RG_OP_L. */
typedef struct rg_run {
    uint32_t row;
    uint32_t op_count; /* op in bits 31..28, count in bits 27..0 */
} rg_run;
enum {
    RG_OP_D = 0,  /* diagonal, lnz[row] == read[col]  ('D') */
    RG_OP_d = 1,  /* diagonal, mismatch               ('d') */
    RG_OP_U = 2,  /* graph char vs gap                ('U'): an outer-loop step of the reference traceback */
    RG_OP_L = 3,  /* read char vs gap                 ('L') */
    RG_OP_Y = 4,  /* 'U' step taken INSIDE a `while path_y == 'Y'` chain (gaf_output.rs:186-200): no handle/dir regrouping */
    RG_OP_UPAD = 5, /* global-mode padding 'U' to the graph start/end (pathwise_alignment_output.rs:116-138) */
    RG_OP_LPAD = 6  /* padding 'L' to column 0 / L-1 (pathwise_alignment_output.rs:111-114) */
};

/* Numeric alignment record of one read: everything GAFStruct (gaf_output.rs:6-20) and the comment field need. */
typedef struct rg_read_result {
    int32_t status;        /* RG_READ_* bits */
    int32_t score;         /* best score (modes 0-7); modes 8/9: baseline (no recombination) score */
    float score_f32;       /* modes 0/1: f32 score as printed; modes 8/9: curr_best_score */
    int32_t displacement;  /* modes 8/9 */
    uint32_t end_row, end_col;     /* traceback start cell (last_row,last_col / ending node, L-1) */
    uint32_t start_row, start_col; /* where the traceback stopped (query_start = start_col) */
    uint32_t best_path, rev_best_path;       /* modes 4-9 */
    uint32_t fen, rsn, rec_col;              /* modes 8/9: forward ending node, reverse starting node, column */
    uint32_t rev_end_row;                    /* modes 8/9: rev_ending_node */
    uint64_t cells;        /* DP cells computed for this read (GCUPS accounting) */
    uint64_t run_off;      /* first run of this read in rg_batch_result.runs */
    uint32_t n_runs;       /* forward part (traceback order) */
    uint32_t n_runs_rev;   /* modes 8/9 with recombination: runs of the reverse half, in FORWARD order, after the first n_runs */
} rg_read_result;

typedef struct rg_batch_result {
    int32_t n_reads;
    rg_read_result* reads; /* n_reads records, input order */
    rg_run* runs;          /* all runs, indexed by run_off */
    uint64_t n_runs_total;
    double kernel_ms;      /* device time of the DP+traceback kernels of this call (CUDA events) */
    uint64_t gpu_launches; /* kernels launched by this call */
} rg_batch_result;

/* ---- lifetime ------------------------------------------------------------------------------------- */
int rg_init(int device, rg_ctx** out);
void rg_destroy(rg_ctx* ctx);
const char* rg_strerror(int status);
const char* rg_last_error(const rg_ctx* ctx);

/* ---- graph ---------------------------------------------------------------------------------------- */
/* Replaces graph::read_graph (graph.rs:11-17) + pathwise_graph::read_graph_w_path (pathwise_graph.rs:127-133)
 * + utils::create_handle_pos_in_lnz (utils.rs:144-165): parses GFA1 (H/S/L/P; integer segment names) and
 * flattens it into device-resident arrays (topologically ordered rows, CSR predecessors, per-edge path
 * bitsets, reverse graph, dfs/dfe). */
int rg_load_gfa_file(rg_ctx* ctx, const char* path);
int rg_load_gfa_text(rg_ctx* ctx, const char* text, size_t len);
/* Replaces handing a prebuilt LnzGraph {lnz, nwp, pred_hash} (graph.rs:23-27) to the exec functions — the
 * form the reference's inline unit tests use (e.g. gap_global_abpoa.rs:465-498). lnz_codes[0] and
 * lnz_codes[n-1] are the '$' and 'F' rows (their codes are ignored). pred_off has n+1 entries; rows without an
 * entry in pred_hash have an empty range. seg_id may be NULL (GAF paths then use 0). Arrays are copied. */
int rg_set_lnz_graph(rg_ctx* ctx, uint32_t n, const uint8_t* lnz_codes, const uint8_t* nwp,
                     const uint32_t* pred_off, const uint32_t* pred_idx, const uint64_t* seg_id);
/* Replaces handing a prebuilt PathGraph {lnz, nwp, pred_hash, paths_nodes, alphas, paths_number, nodes_id_pos}
 * (pathwise_graph.rs:10-18) to the pathwise exec functions — the form the reference's inline tests build
 * (pathwise_graph.rs:364-544). PredHash as CSR: pred_off has n+1 entries, every (node, predecessor) entry k carries the
 * bitset of the paths using that edge in edge_path_bits[k * PW .. (k+1) * PW), PW = (n_paths + 31) / 32, bit p of word
 * p / 32. node_path_bits: n * PW words (paths_nodes); alphas: n entries; seg_id: nodes_id_pos (0 for rows 0 and n-1).
 * The reverse graph and the distance vectors behind nodes_displacement_matrix (pathwise_graph.rs:250-354) are derived
 * inside. Arrays are copied. After this call modes 4-9 use the given graph; the POA modes see the same predecessor lists. */
int rg_set_path_graph(rg_ctx* ctx, uint32_t n, uint32_t n_paths, const uint8_t* lnz_codes, const uint8_t* nwp,
                      const uint32_t* pred_off, const uint32_t* pred_idx, const uint32_t* edge_path_bits,
                      const uint32_t* node_path_bits, const uint32_t* alphas, const uint64_t* seg_id);
/* Graph facts for callers: n = lnz.len(), P = paths_number. */
int rg_graph_info(const rg_ctx* ctx, uint32_t* n, uint32_t* n_segments, uint32_t* n_paths);

/* ---- scoring -------------------------------------------------------------------------------------- */
/* Replaces score_matrix::create_score_matrix_match_mis (score_matrix.rs:35-51; gap entries = 2*x),
 * create_score_matrix_match_mis_f32 (52-66; gap entries = x) and create_score_matrix_from_matrix_file
 * (67-105; HOXD55/HOXD70 embedded, gap entries = -200). kind: 0 = match/mismatch i32, 1 = match/mismatch f32
 * (api.rs default), 2 = "HOXD55", 3 = "HOXD70". Fills s->score only. */
int rg_make_score_matrix(int kind, int32_t match, int32_t mismatch, rg_scoring* s);
/* CLI defaults of args_parser.rs:3-147: M=2 X=4 O=4 E=2 R=4 r=0.1 B=1.0 b=1 f=0.01. */
void rg_default_scoring(rg_scoring* s);
int rg_set_scoring(rg_ctx* ctx, const rg_scoring* s);

/* ---- alignment (the hot path) --------------------------------------------------------------------- */
/* Replaces the per-read loops of main.rs:56-312 over global_abpoa::exec_simd (global_abpoa.rs:10-257),
 * local_poa::exec_simd (local_poa.rs:10-179), gap_global_abpoa::exec (gap_global_abpoa.rs:11-250),
 * gap_local_poa::exec (gap_local_poa.rs:8-129), pathwise_alignment::exec (pathwise_alignment.rs:5-340),
 * pathwise_alignment_semiglobal::exec (pathwise_alignment_semiglobal.rs:6-242) and
 * pathwise_alignment_recombination::exec (pathwise_alignment_recombination.rs:23-127), including their
 * tracebacks. read_codes: concatenated codes (A=0..N=4) WITHOUT the '$' sentinel; read_off: n_reads+1 offsets.
 * Host buffers in, host records out; H2D/D2H copies happen inside the call. */
int rg_align_batch(rg_ctx* ctx, int mode, int32_t n_reads, const uint8_t* read_codes, const uint64_t* read_off,
                   rg_batch_result* out);
/* The same computation in three steps, so that callers (bench.py) can time the kernels with the reads already
 * resident in HBM: stage the reads on the device, run the DP + traceback kernels leaving the records in HBM,
 * fetch the records. rg_align_batch == rg_upload_reads + rg_align_staged + rg_fetch_results. */
int rg_upload_reads(rg_ctx* ctx, int32_t n_reads, const uint8_t* read_codes, const uint64_t* read_off);
int rg_align_staged(rg_ctx* ctx, int mode);
int rg_fetch_results(rg_ctx* ctx, rg_batch_result* out);
/* Device time (ms, CUDA events on the ctx stream) and kernel launches of the last rg_align_staged call. */
int rg_last_kernel_stats(const rg_ctx* ctx, double* kernel_ms, uint64_t* launches, uint64_t* cells);

/* ---- output --------------------------------------------------------------------------------------- */
/* Replaces GAFStruct::to_string (gaf_output.rs:70-94) and the six gaf_of_* / build_alignment / gaf_output_*
 * builders' string work. Writes what the reference prints to stdout for this read — warning lines included —
 * into buf (NUL-terminated, '\n'-terminated lines). Returns the length needed (excluding NUL); call again with a
 * larger buffer if the return value >= cap.
 * amb_mode: 0 for a forward alignment; for the `-s` reverse-strand retries (main.rs:82-101,150-164,198-214,233-249)
 * 1 = strand '-' and segment ids through the reversed handle map (utils.rs:144-165, amb_mode = true), or an OR of
 * RG_AMB_HANDLES / RG_AMB_STRAND (mode 3 passes amb_mode = false with the reversed map, main.rs:242). */
enum { RG_AMB_HANDLES = 2, RG_AMB_STRAND = 4 };
int64_t rg_format_gaf(rg_ctx* ctx, int mode, const rg_batch_result* res, int32_t read_index, const char* read_name,
                      uint32_t read_len, int amb_mode, char* buf, size_t cap);

/* The same for every read of a batch in input order, into one malloc'd buffer (free with rg_free): the per-read println! /
 * write_gaf loop of main.rs:56-312 as one call, so that emitting GAF is not one FFI round trip per read. names may be NULL
 * ("read<first_index + i>" is used, the naming of this repo's synthetic FASTA — first_index lets a shard of a larger read
 * set keep its global numbering); read_off are the offsets given to rg_align_batch. */
int rg_format_gaf_all(rg_ctx* ctx, int mode, const rg_batch_result* res, const char* const* names, int64_t first_index,
                      const uint64_t* read_off, int amb_mode, char** out_text, size_t* out_len);

/* FASTA reader with the reference's normalisation (sequences.rs:5-45: upper-case, '-' -> 'N'); encodes to codes.
 * Buffers are malloc'd; free with rg_free_reads. */
typedef struct rg_reads {
    int32_t n_reads;
    uint8_t* codes;
    uint64_t* off;  /* n_reads + 1 */
    char** names;
} rg_reads;
int rg_read_fasta_file(const char* path, rg_reads* out, char* errbuf, size_t errcap);
int rg_read_fasta_text(const char* text, size_t len, rg_reads* out, char* errbuf, size_t errcap);
void rg_free_reads(rg_reads* r);

/* Whole-CLI entry (main.rs:25-329): argv as the reference binary takes it. stdout/stderr text is returned in
 * malloc'd buffers (free with rg_free). Used by the `recgraph` executable and the parity tests. */
int rg_cli_main(int argc, const char** argv, char** out_text, char** err_text);
void rg_free(void* p);

/* Host-side diagnostics, no device needed: the flattened LnzGraph (graph.rs:31-123 + utils.rs:103-165: lnz, nwp, predecessor
 * lists in list order, segment id per row, r-values) and the PathGraph or its reverse (pathwise_graph.rs:135-354: per-row
 * alpha and path sets, per-edge path sets, distance-from-start / from-end) of a GFA text as a malloc'd text dump
 * (free with rg_free; "ERROR ..." on malformed input). Used by the CPU tests of the host builders. */
char* rg_debug_dump_lnz(const char* gfa_text, size_t len);
char* rg_debug_dump_pathgraph(const char* gfa_text, size_t len, int reverse_graph);

/* INT32 ALU microbenchmark used as the roofline denominator: giga-ops/s of IADD3, VIMNMX,
 * VIADDMNMX measured on the ctx's device. */
int rg_int_peak(rg_ctx* ctx, double* iadd_gops, double* imnmx_gops, double* viaddmnmx_gops);

#ifdef __cplusplus
}
#endif
#endif /* RECGRAPH_B200_H */
