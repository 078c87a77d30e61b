"""ctypes access to oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Nothing under recgraph_b200/ may import this module.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_lib = None
_built = False


def build():
    """Always ask make: it checks liboracle.so AND the recgraph_oracle executable against the sources."""
    global _built
    if _built:
        return
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-j8"], stdout=subprocess.DEVNULL)
    _built = True


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(LIB)
    lib.rgo_free.argtypes = [ctypes.c_void_p]
    lib.rgo_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p),
                             ctypes.POINTER(ctypes.c_void_p)]
    lib.rgo_main.restype = ctypes.c_int
    lib.rgo_poa_score.restype = ctypes.c_int
    for f in ("rgo_dump_lnz", "rgo_dump_pathgraph", "rgo_rev_and_compl", "rgo_f32_display", "rgo_pathwise_one"):
        getattr(lib, f).restype = ctypes.c_void_p
    lib.rgo_dump_lnz.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.rgo_pathwise_one.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 5
    lib.rgo_dump_pathgraph.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    lib.rgo_rev_and_compl.argtypes = [ctypes.c_char_p]
    lib.rgo_f32_display.argtypes = [ctypes.c_float]
    lib.rgo_score_lookup.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char, ctypes.c_char,
                                     ctypes.POINTER(ctypes.c_int)]
    lib.rgo_score_lookup.restype = ctypes.c_int
    lib.rgo_bases_to_add.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
    lib.rgo_bases_to_add.restype = ctypes.c_int
    _lib = lib
    return lib


def _take(lib, p):
    s = ctypes.string_at(p).decode()
    lib.rgo_free(p)
    return s


def run_cli(args):
    """Run the restated `recgraph` CLI in-process. Returns (exit_code, stdout, stderr)."""
    lib = load()
    argv = [b"recgraph"] + [str(a).encode() for a in args]
    arr = (ctypes.c_char_p * len(argv))(*argv)
    out = ctypes.c_void_p()
    err = ctypes.c_void_p()
    rc = lib.rgo_main(len(argv), arr, ctypes.byref(out), ctypes.byref(err))
    return rc, _take(lib, out), _take(lib, err)


def poa_score(variant, lnz, nwp_idx, pred_hash, read, scores, o=0, e=0, bta=0):
    """Score of a POA variant on a hand-built LnzGraph (mirrors the reference's inline unit tests)."""
    lib = load()
    n = len(lnz)
    nwp = (ctypes.c_uint8 * n)(*[1 if i in nwp_idx else 0 for i in range(n)])
    off = [0]
    idx = []
    for i in range(n):
        idx += pred_hash.get(i, [])
        off.append(len(idx))
    pred_off = (ctypes.c_uint32 * (n + 1))(*off)
    pred_idx = (ctypes.c_uint32 * max(1, len(idx)))(*idx)
    keys = "".join(a + b for (a, b) in scores.keys()).encode()
    vals = (ctypes.c_int * len(scores))(*scores.values())
    score = ctypes.c_int()
    cells = ctypes.c_uint64()
    rc = lib.rgo_poa_score(variant, n, "".join(lnz).encode(), nwp, pred_off, pred_idx, len(read),
                           "".join(read).encode(), len(scores), keys, vals, o, e, bta, ctypes.byref(score),
                           ctypes.byref(cells))
    return rc, score.value, cells.value


def dump_lnz(gfa_text, amb_mode=False):
    lib = load()
    return _take(lib, lib.rgo_dump_lnz(gfa_text.encode(), int(amb_mode)))


def dump_pathgraph(gfa_text, is_reversed=False, reverse_graph=False):
    lib = load()
    return _take(lib, lib.rgo_dump_pathgraph(gfa_text.encode(), int(is_reversed), int(reverse_graph)))


def pathwise_one(gfa_text, read, mode, m=2, x=-4, o=-4, e=-2):
    lib = load()
    return _take(lib, lib.rgo_pathwise_one(gfa_text.encode(), read.encode(), mode, m, x, o, e))


def rev_and_compl(s):
    lib = load()
    return _take(lib, lib.rgo_rev_and_compl(s.encode()))


def f32_display(v):
    lib = load()
    return _take(lib, lib.rgo_f32_display(v))
