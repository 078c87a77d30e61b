"""Host-side mirror of the reference's interface over the C ABI (see include/recgraph_b200.h).

`run_cli(argv)` == the reference binary's `main` (main.rs:25-329); `Aligner` == one device context with the
graph resident in HBM, exposing the batch form of the exec functions that main.rs loops over.
"""
import ctypes

import numpy as np

from . import _lib

CODES = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}
_LUT = np.full(256, 255, dtype=np.uint8)
for _c, _v in CODES.items():
    _LUT[ord(_c)] = _v
    _LUT[ord(_c.lower())] = _v
_LUT[ord("-")] = 4  # sequences.rs:17-19: '-' -> 'N'


class RecGraphError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"recgraph_b200 status {status}: {msg}")
        self.status = status


def encode_read(seq: str) -> np.ndarray:
    codes = _LUT[np.frombuffer(seq.encode(), dtype=np.uint8)]
    if (codes == 255).any():
        raise RecGraphError(-5, "read character outside A,C,G,T,N")
    return codes


def run_cli(argv):
    """`recgraph <argv...>` in-process. Returns (exit_code, stdout_text, stderr_text)."""
    lib = _lib.load()
    args = [b"recgraph"] + [str(a).encode() for a in argv]
    arr = (ctypes.c_char_p * len(args))(*args)
    out, err = ctypes.c_void_p(), ctypes.c_void_p()
    rc = lib.rg_cli_main(len(args), arr, ctypes.byref(out), ctypes.byref(err))
    o = ctypes.string_at(out).decode() if out else ""
    e = ctypes.string_at(err).decode() if err else ""
    lib.rg_free(out)
    lib.rg_free(err)
    return rc, o, e


class Aligner:
    """One rg_ctx: one CUDA device, one stream, one resident graph."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.rg_init(device, ctypes.byref(self.ctx))
        if rc != 0:
            raise RecGraphError(rc, self.lib.rg_strerror(rc).decode())
        self.scoring = _lib.Scoring()
        self.lib.rg_default_scoring(ctypes.byref(self.scoring))
        self._keep = None

    def close(self):
        if self.ctx:
            self.lib.rg_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RecGraphError(rc, self.lib.rg_last_error(self.ctx).decode() or self.lib.rg_strerror(rc).decode())

    # ---- graph
    def load_gfa(self, path):
        self._check(self.lib.rg_load_gfa_file(self.ctx, str(path).encode()))

    def load_gfa_text(self, text):
        b = text.encode()
        self._check(self.lib.rg_load_gfa_text(self.ctx, b, len(b)))

    def set_lnz_graph(self, lnz, nwp_idx, pred_hash, seg_id=None):
        """Hand-built LnzGraph as in the reference's inline tests: lnz chars incl. '$' and 'F'."""
        n = len(lnz)
        codes = np.array([CODES.get(c, 0) for c in lnz], dtype=np.uint8)
        nwp = np.zeros(n, dtype=np.uint8)
        nwp[list(nwp_idx)] = 1
        off = [0]
        idx = []
        for i in range(n):
            idx += list(pred_hash.get(i, []))
            off.append(len(idx))
        off = np.array(off, dtype=np.uint32)
        idx = np.array(idx if idx else [0], dtype=np.uint32)
        sid = None if seg_id is None else np.array(seg_id, dtype=np.uint64)
        self._check(self.lib.rg_set_lnz_graph(self.ctx, n, codes.ctypes.data, nwp.ctypes.data, off.ctypes.data,
                                              idx.ctypes.data, None if sid is None else sid.ctypes.data))

    def set_path_graph(self, lnz, nwp, preds, paths_nodes, alphas, n_paths, nodes_id_pos=None):
        """Prebuilt PathGraph (pathwise_graph.rs:10-18): lnz chars incl. '$' / 'F', nwp flags, preds = {node: {pred: set(paths)}},
        paths_nodes = per-row sets of path ids, alphas per row."""
        n = len(lnz)
        PW = (n_paths + 31) // 32
        codes = np.array([CODES.get(c, 0) for c in lnz], dtype=np.uint8)
        nwp_a = np.array([1 if x else 0 for x in nwp], dtype=np.uint8)
        off, idx, ebits = [0], [], []
        for i in range(n):
            for p, ps in preds.get(i, {}).items():
                idx.append(p)
                w = [0] * PW
                for q in ps:
                    w[q // 32] |= 1 << (q % 32)
                ebits += w
            off.append(len(idx))
        nbits = []
        for i in range(n):
            w = [0] * PW
            for q in paths_nodes[i]:
                w[q // 32] |= 1 << (q % 32)
            nbits += w
        off = np.array(off, dtype=np.uint32)
        idx = np.array(idx if idx else [0], dtype=np.uint32)
        ebits = np.array(ebits if ebits else [0], dtype=np.uint32)
        nbits = np.array(nbits, dtype=np.uint32)
        al = np.array(alphas, dtype=np.uint32)
        sid = None if nodes_id_pos is None else np.array(nodes_id_pos, dtype=np.uint64)
        self._keep_graph = (codes, nwp_a, off, idx, ebits, nbits, al, sid)
        self._check(self.lib.rg_set_path_graph(self.ctx, n, n_paths, codes.ctypes.data, nwp_a.ctypes.data, off.ctypes.data,
                                               idx.ctypes.data, ebits.ctypes.data, nbits.ctypes.data, al.ctypes.data,
                                               None if sid is None else sid.ctypes.data))

    def graph_info(self):
        n, s, p = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.rg_graph_info(self.ctx, ctypes.byref(n), ctypes.byref(s), ctypes.byref(p)))
        return n.value, s.value, p.value

    # ---- scoring
    def set_scoring(self, match=2, mismatch=4, gap_open=4, gap_ext=2, matrix="none", base_rec_cost=4,
                    multi_rec_cost=0.1, rec_band_width=1.0, extra_b=1, extra_f=0.01, fixed_bta=-1, kind=None,
                    table=None):
        """CLI-style parameters (penalties positive, as on the reference command line)."""
        s = self.scoring
        if table is not None:
            for i in range(6):
                for j in range(6):
                    s.score[i][j] = int(table[i][j])
        else:
            if kind is None:
                kind = {"none": 0, "HOXD55": 2, "HOXD55.mtx": 2, "HOXD70": 3, "HOXD70.mtx": 3}[matrix]
            self._check(self.lib.rg_make_score_matrix(kind, match, -mismatch, ctypes.byref(s)))
        s.gap_open, s.gap_ext = -gap_open, -gap_ext
        s.base_rec_cost, s.multi_rec_cost, s.rec_band_width = base_rec_cost, multi_rec_cost, rec_band_width
        s.extra_b, s.extra_f, s.fixed_bta = float(extra_b), extra_f, fixed_bta
        self._check(self.lib.rg_set_scoring(self.ctx, ctypes.byref(s)))

    # ---- alignment
    @staticmethod
    def pack_reads(reads):
        """list of str / uint8 code arrays -> (codes, offsets) numpy arrays."""
        arrs = [encode_read(r) if isinstance(r, str) else np.asarray(r, dtype=np.uint8) for r in reads]
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        if arrs:
            off[1:] = np.cumsum([len(a) for a in arrs])
        codes = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
        return np.ascontiguousarray(codes), off

    def align_packed(self, mode, codes, off):
        """rg_align_batch: host buffers in, records out (H2D + kernels + D2H)."""
        res = _lib.BatchResult()
        self._keep = (codes, off)
        self._check(self.lib.rg_align_batch(self.ctx, mode, len(off) - 1, codes.ctypes.data, off.ctypes.data,
                                            ctypes.byref(res)))
        return res

    def upload(self, codes, off):
        self._keep = (codes, off)
        self._check(self.lib.rg_upload_reads(self.ctx, len(off) - 1, codes.ctypes.data, off.ctypes.data))

    def align_staged(self, mode):
        self._check(self.lib.rg_align_staged(self.ctx, mode))

    def fetch(self):
        res = _lib.BatchResult()
        self._check(self.lib.rg_fetch_results(self.ctx, ctypes.byref(res)))
        return res

    def kernel_stats(self):
        ms, launches, cells = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64()
        self.lib.rg_last_kernel_stats(self.ctx, ctypes.byref(ms), ctypes.byref(launches), ctypes.byref(cells))
        return ms.value, launches.value, cells.value

    def format_gaf(self, mode, res, index, name, read_len, amb_mode=False):
        buf = ctypes.create_string_buffer(1 << 16)
        need = self.lib.rg_format_gaf(self.ctx, mode, ctypes.byref(res), index, name.encode(), read_len,
                                      int(amb_mode), buf, len(buf))
        if need < 0:
            raise RecGraphError(need, "rg_format_gaf failed")
        if need >= len(buf):
            buf = ctypes.create_string_buffer(need + 1)
            self.lib.rg_format_gaf(self.ctx, mode, ctypes.byref(res), index, name.encode(), read_len, int(amb_mode),
                                   buf, len(buf))
        return buf.value.decode()

    def format_gaf_all(self, mode, res, off, amb_mode=False, first_index=0):
        """GAF text of the whole batch in input order (read names "read<first_index + i>")."""
        out, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self.lib.rg_format_gaf_all(self.ctx, mode, ctypes.byref(res), None, first_index, off.ctypes.data, int(amb_mode),
                                               ctypes.byref(out), ctypes.byref(n)))
        text = ctypes.string_at(out, n.value).decode()
        self.lib.rg_free(out)
        return text

    def align(self, mode, reads, names=None):
        """Align a list of reads; returns (records, gaf_text)."""
        codes, off = self.pack_reads(reads)
        res = self.align_packed(mode, codes, off)
        names = names or [f"read{i}" for i in range(len(reads))]
        text = "".join(self.format_gaf(mode, res, i, names[i], int(off[i + 1] - off[i])) for i in range(len(reads)))
        recs = [res.reads[i] for i in range(res.n_reads)]
        return recs, text

    def int_peak(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.rg_int_peak(self.ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"iadd3_gops": a.value, "vimnmx_gops": b.value, "viaddmnmx_gops": c.value}


# ---------------------------------------------------------------------------------------------------------------
# api.rs:11-164 — library entry points (only the POA modes are exposed there). Same names, argument meaning and
# defaults as the reference: `bases_to_add` is a FRACTION of the read length (default 0.1, api.rs:21,55), the
# default f32 matrix has gap-vs-char = X (score_matrix.rs:52-66, not 2X), default o = -10, e = -6 (api.rs:65-66).
# `graph` is an Aligner with a graph loaded (the HashGraph argument of the reference); scores are passed as the
# 6x6 table over A,C,G,T,N,'-' (see create_score_matrix_i32 / _f32 below).
class GAFStruct:
    def __init__(self, line):
        f = line.rstrip("\n").split("\t")
        self.line = line.rstrip("\n")
        self.query_name, self.query_length, self.query_start, self.query_end = f[0], int(f[1]), int(f[2]), int(f[3])
        self.strand = f[4]
        self.path = [int(x) for x in f[5].split(">") if x]
        self.path_length, self.path_start, self.path_end = int(f[6]), int(f[7]), int(f[8])
        self.residue_matches_number = int(f[9])
        self.alignment_block_length, self.mapping_quality = f[10], f[11]
        self.comments = "\t".join(f[12:])

    def to_string(self):
        return self.line


def create_score_matrix_i32(match_score=None, mismatch_score=None, matrix_type=None):
    """api.rs:131-149. mismatch_score is the (negative) score itself, as in the reference."""
    s = _lib.Scoring()
    lib = _lib.load()
    kind = {None: 0, "HOXD55.mtx": 2, "HOXD55": 2, "HOXD70.mtx": 3, "HOXD70": 3}[matrix_type]
    lib.rg_make_score_matrix(kind, match_score or 0, mismatch_score or 0, ctypes.byref(s))
    return [[s.score[i][j] for j in range(6)] for i in range(6)]


def create_score_matrix_f32(match_score=None, mismatch_score=None, matrix_type=None):
    """api.rs:153-164: same values as the i32 builder, as f32."""
    return create_score_matrix_i32(match_score, mismatch_score, matrix_type)


def _default_f32_table():
    s = _lib.Scoring()
    _lib.load().rg_make_score_matrix(1, 2, -4, ctypes.byref(s))  # create_score_matrix_match_mis_f32(2, -4)
    return [[s.score[i][j] for j in range(6)] for i in range(6)]


def _api_align(graph, mode, read, sequence_name, table, bta, o=-10, e=-6):
    name = sequence_name[0] if sequence_name else "no_name"
    graph.set_scoring(table=table, gap_open=-o, gap_ext=-e, fixed_bta=bta if bta is not None else -1)
    _recs, text = graph.align(mode, [read.upper().replace("-", "N")], names=[name])
    lines = text.splitlines()
    return GAFStruct(lines[-1]) if lines else None


def _bases_to_add(read, frac):
    import numpy as _np
    return int(_np.float32(len(read)) * _np.float32(0.1 if frac is None else frac))  # api.rs:21: f32 product, truncated


def align_global_no_gap(read, graph, sequence_name=None, score_matrix=None, bases_to_add=None):
    """api.rs:11-40 -> global_abpoa::exec_simd."""
    return _api_align(graph, 0, read, sequence_name, score_matrix or _default_f32_table(), _bases_to_add(read, bases_to_add))


def align_global_gap(read, graph, sequence_name=None, score_matrix=None, bases_to_add=None, o=None, e=None):
    """api.rs:43-72 -> gap_global_abpoa::exec."""
    table = score_matrix or create_score_matrix_i32(2, -4)
    return _api_align(graph, 2, read, sequence_name, table, _bases_to_add(read, bases_to_add),
                      -10 if o is None else o, -6 if e is None else e)


def align_local_no_gap(read, graph, sequence_name=None, score_matrix=None):
    """api.rs:76-99 -> local_poa::exec_simd."""
    return _api_align(graph, 1, read, sequence_name, score_matrix or _default_f32_table(), None)


def align_local_gap(read, graph, sequence_name=None, score_matrix=None, o=None, e=None):
    """api.rs:102-128 -> gap_local_poa::exec."""
    table = score_matrix or create_score_matrix_i32(2, -4)
    return _api_align(graph, 3, read, sequence_name, table, None, -10 if o is None else o, -6 if e is None else e)
