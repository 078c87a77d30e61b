"""recgraph_b200 — B200-native RecGraph aligner (sequence-to-graph DP on sm_100a behind a C ABI).

Python is only the harness around the C ABI (include/recgraph_b200.h): `Aligner` wraps one context,
`run_cli` calls the `recgraph` command line in-process. Nothing here computes alignments on the CPU.
"""
from .api import (Aligner, GAFStruct, RecGraphError, align_global_gap, align_global_no_gap,  # noqa: F401
                  align_local_gap, align_local_no_gap, create_score_matrix_f32, create_score_matrix_i32,
                  encode_read, run_cli)
