"""ctypes binding of include/recgraph_b200.h (the C-ABI drop-in boundary).

The shared library is built in-tree by `__graft_entry__.build()` / `make -C recgraph_b200/csrc`.
There is no fallback of any kind: if the library or a CUDA device is missing the calls fail loudly.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RG_LIB_DIR: an alternative in-tree build directory (kernel tuning experiments: `make OUT=../_build_x EXTRA=-D...`)
LIB_PATH = os.path.join(os.environ.get("RG_LIB_DIR") or os.path.join(HERE, "_build"), "librecgraph_b200.so")

c_i32, c_u32, c_u64, c_f32 = ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float


class Scoring(ctypes.Structure):  # rg_scoring
    _fields_ = [("score", (c_i32 * 6) * 6), ("gap_open", c_i32), ("gap_ext", c_i32), ("base_rec_cost", c_i32),
                ("multi_rec_cost", c_f32), ("rec_band_width", c_f32), ("extra_b", c_f32), ("extra_f", c_f32),
                ("fixed_bta", c_i32)]


class Run(ctypes.Structure):  # rg_run
    _fields_ = [("row", c_u32), ("op_count", c_u32)]


class ReadResult(ctypes.Structure):  # rg_read_result
    _fields_ = [("status", c_i32), ("score", c_i32), ("score_f32", c_f32), ("displacement", c_i32),
                ("end_row", c_u32), ("end_col", c_u32), ("start_row", c_u32), ("start_col", c_u32),
                ("best_path", c_u32), ("rev_best_path", c_u32), ("fen", c_u32), ("rsn", c_u32), ("rec_col", c_u32),
                ("rev_end_row", c_u32), ("cells", c_u64), ("run_off", c_u64), ("n_runs", c_u32), ("n_runs_rev", c_u32)]


class BatchResult(ctypes.Structure):  # rg_batch_result
    _fields_ = [("n_reads", c_i32), ("reads", ctypes.POINTER(ReadResult)), ("runs", ctypes.POINTER(Run)),
                ("n_runs_total", c_u64), ("kernel_ms", ctypes.c_double), ("gpu_launches", c_u64)]


class Reads(ctypes.Structure):  # rg_reads
    _fields_ = [("n_reads", c_i32), ("codes", ctypes.POINTER(ctypes.c_uint8)), ("off", ctypes.POINTER(c_u64)),
                ("names", ctypes.POINTER(ctypes.c_char_p))]


# every symbol include/recgraph_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["rg_init", "rg_destroy", "rg_strerror", "rg_last_error", "rg_load_gfa_file", "rg_load_gfa_text",
           "rg_set_lnz_graph", "rg_set_path_graph", "rg_graph_info", "rg_make_score_matrix", "rg_default_scoring", "rg_set_scoring",
           "rg_align_batch", "rg_upload_reads", "rg_align_staged", "rg_fetch_results", "rg_last_kernel_stats",
           "rg_format_gaf", "rg_format_gaf_all", "rg_read_fasta_file", "rg_read_fasta_text", "rg_free_reads", "rg_cli_main", "rg_free",
           "rg_int_peak", "rg_debug_dump_lnz", "rg_debug_dump_pathgraph"]

_lib = None


def load():
    """Load librecgraph_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C recgraph_b200/csrc` (recgraph_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.rg_init.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.rg_destroy.argtypes = [vp]
    lib.rg_destroy.restype = None
    lib.rg_strerror.argtypes = [ctypes.c_int]
    lib.rg_strerror.restype = ctypes.c_char_p
    lib.rg_last_error.argtypes = [vp]
    lib.rg_last_error.restype = ctypes.c_char_p
    lib.rg_load_gfa_file.argtypes = [vp, ctypes.c_char_p]
    lib.rg_load_gfa_text.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    lib.rg_set_lnz_graph.argtypes = [vp, c_u32, vp, vp, vp, vp, vp]
    lib.rg_set_path_graph.argtypes = [vp, c_u32, c_u32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rg_graph_info.argtypes = [vp, ctypes.POINTER(c_u32), ctypes.POINTER(c_u32), ctypes.POINTER(c_u32)]
    lib.rg_make_score_matrix.argtypes = [ctypes.c_int, c_i32, c_i32, ctypes.POINTER(Scoring)]
    lib.rg_default_scoring.argtypes = [ctypes.POINTER(Scoring)]
    lib.rg_default_scoring.restype = None
    lib.rg_set_scoring.argtypes = [vp, ctypes.POINTER(Scoring)]
    lib.rg_align_batch.argtypes = [vp, ctypes.c_int, c_i32, vp, vp, ctypes.POINTER(BatchResult)]
    lib.rg_upload_reads.argtypes = [vp, c_i32, vp, vp]
    lib.rg_align_staged.argtypes = [vp, ctypes.c_int]
    lib.rg_fetch_results.argtypes = [vp, ctypes.POINTER(BatchResult)]
    lib.rg_debug_dump_lnz.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    lib.rg_debug_dump_lnz.restype = vp
    lib.rg_debug_dump_pathgraph.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    lib.rg_debug_dump_pathgraph.restype = vp
    lib.rg_last_kernel_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_u64),
                                         ctypes.POINTER(c_u64)]
    lib.rg_format_gaf.argtypes = [vp, ctypes.c_int, ctypes.POINTER(BatchResult), c_i32, ctypes.c_char_p, c_u32,
                                  ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    lib.rg_format_gaf.restype = ctypes.c_int64
    lib.rg_format_gaf_all.argtypes = [vp, ctypes.c_int, ctypes.POINTER(BatchResult), vp, ctypes.c_int64, vp, ctypes.c_int, ctypes.POINTER(vp),
                                      ctypes.POINTER(ctypes.c_size_t)]
    lib.rg_read_fasta_file.argtypes = [ctypes.c_char_p, ctypes.POINTER(Reads), ctypes.c_char_p, ctypes.c_size_t]
    lib.rg_read_fasta_text.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(Reads), ctypes.c_char_p,
                                       ctypes.c_size_t]
    lib.rg_free_reads.argtypes = [ctypes.POINTER(Reads)]
    lib.rg_free_reads.restype = None
    lib.rg_cli_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(vp),
                                ctypes.POINTER(vp)]
    lib.rg_free.argtypes = [vp]
    lib.rg_free.restype = None
    lib.rg_int_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib
