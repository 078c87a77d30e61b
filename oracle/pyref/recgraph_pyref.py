"""SECOND, INDEPENDENT restatement of the reference's pathwise / recombination path — TEST INFRASTRUCTURE ONLY.

Written in plain Python from the Rust sources alone (never from oracle/*.cpp), so that the C++ oracle — the checker of
the CUDA path — is itself checked by a second reading of the reference for the parts the reference's own tests do not pin
(tests/test_pyref_vs_oracle.py diffs the two on the example and on hundreds of random small graphs):

  pathwise_graph.rs:135-354                      create_path_graph, create_reverse_path_graph, displacement
  pathwise_alignment.rs:5-340                    mode 4 exec
  pathwise_alignment_semiglobal.rs:6-277         mode 5 exec, best_ending_node
  pathwise_alignment_output.rs:7-184,471-556     build_alignment, build_cigar
  pathwise_alignment_recombination.rs:9-897      exec, rev_align, align, absolute_scores, best_alignment, ending_node
  recombination_output.rs:12-782                 the four gaf_output_* builders
  utils.rs:221-323, gaf_output.rs:70-94          get_path_len_start_end, get_rec_path_len_start_end, GAFStruct::to_string
  score_matrix.rs:21-51,67-105 (match / mismatch and the two matrix files), sequences.rs:5-45, main.rs:253-312
and, for the headline mode 2 whose traceback the reference's unit tests do not pin either:
  graph.rs:31-123, utils.rs:17-72,103-165      create_graph_struct, set_ampl_for_row, set_r_values, handle map
  gap_global_abpoa.rs:11-455                     exec, get_best_d / u / l, band_ampl_enough
  bitfield_path.rs:3-44                          the 32-bit trace cell with its 16-bit predecessor
  gaf_output.rs:96-253,867-892                   gaf_of_gap_abpoa, node_start, set_cigar_substring

and for the other POA modes on an AVX2 machine (what the reference's CLI runs):
  global_abpoa.rs:10-257, gaf_output.rs:753-865  mode 0: exec_simd lane by lane, f32 path values decoded through their text
  local_poa.rs:10-179, gaf_output.rs:639-751     mode 1: exec_simd, gaf_of_local_poa_simd
  gap_local_poa.rs:8-187, gaf_output.rs:502-637  mode 3: exec (get_best_d / get_best_u with their never-set `first`)
  utils.rs:74-99,129-140                         set_left_right_x64, get_max_d_u_l
and the -s true flows of modes 0-3 (main.rs:47-252):
  global_abpoa.rs:260-566, gaf_output.rs:254-382 the scalar exec of mode 0's retry, gaf_of_global_abpoa
  sequences.rs:64-82, utils.rs:144-165           rev_and_compl, the reversed handle map, strand '-', selection rules
and the experimental affine pathwise modes 6 / 7:
  pathwise_alignment_gap.rs:4-574, pathwise_alignment_gap_semi.rs:5-473   exec (delta-encoded dpm / x / y), best_ending_node
  pathwise_alignment_output.rs:186-451           build_alignment_gap, build_alignment_semiglobal_gap; main.rs:271-288

`rev_align` is `align` mirrored in i and j (checked mechanically: sed 's/i + 1/i - 1/; s/j + 1/j - 1/' on lines 129-435
diffs clean against 436-745 apart from the border cases), so one cell routine parameterised by direction serves both.
HashMap iteration orders of the reference (predecessors of a node) are fixed to ascending predecessor index.
f32 arithmetic uses numpy.float32 so that every operation rounds as in Rust. Integer overflow follows a RELEASE build (the
reference is built with `cargo build --release`): it wraps; a RuntimeError is raised only where the wrapped value is certain to
panic next (a `vec![..; huge]` capacity overflow, an index out of bounds) — bounds checks stay on in release builds.
"""
import numpy as np

F32 = np.float32
_STRAND = ["+"]   # strand field of the POA GAF builders: '-' while a builder runs with amb_mode = true (-s retries)


# ------------------------------------------------------------------------------------------------ inputs
def read_gfa(text):
    """GFA1 S / L / P lines with integer segment names (gfa crate, GFA<usize, ()>)."""
    segs, paths = {}, []
    for ln in text.splitlines():
        f = ln.split("\t")
        if f[0] == "S":
            segs[int(f[1])] = f[2]
        elif f[0] == "P":
            paths.append([int(s[:-1]) for s in f[2].split(",")])
    return segs, paths


def read_fasta(text):
    """sequences.rs:5-45"""
    names, seqs, cur = [], [], []
    for line in text.splitlines():
        if not line.startswith(">") and line != "":
            cur += ["N" if c == "-" else c.upper() for c in line]
        elif line.startswith(">"):
            names.append(line[1:])
            if cur:
                seqs.append(["$"] + cur)
            cur = []
    if cur:
        seqs.append(["$"] + cur)
    assert len(seqs) == len(names), "wrong fasta file format"
    return seqs, names


def score_matrix_match_mis(m, x):
    """score_matrix.rs:35-51"""
    sm = {}
    for i in "ACGTN-":
        for j in "ACGTN-":
            if i == j:
                sm[(i, j)] = m
            elif i == "-" or j == "-":
                sm[(i, j)] = x * 2
            else:
                sm[(i, j)] = x
    sm[("N", "N")] = x
    del sm[("-", "-")]
    return sm


# the two matrix files shipped with the reference (HOXD70.mtx, HOXD55.mtx), transcribed; note (G, T) != (T, G) in HOXD70
_MTX = {
    "HOXD70": """    A       C       G       T       N
A   91      -114    -31     -123    0
C   -114    100     -125    -31     0
G   -31     -125    100     -114    0
T   -123    -31     -144    91      0
N   0       0       0       0       0
""",
    "HOXD55": """    A       C       G       T       N
A   91      -90     -25     -100    0
C   -90     100     -100    -25     0
G   -25     -100    100     -90     0
T   -100    -25     -90     91      0
N   0       0       0       0       0
""",
}


def score_matrix_from_matrix_file(name):
    """score_matrix.rs:67-105: key (row character, column character)"""
    matrix = [[e for e in line.split(" ") if e != ""] for line in _MTX[name].splitlines()]
    matrix[0].insert(0, "X")
    sm = {}
    for i in range(1, len(matrix)):
        for j in range(1, len(matrix[0])):
            sm[(matrix[i][0][0], matrix[0][j][0])] = int(matrix[i][j])
    for ch in "ACGTN":
        sm[(ch, "-")] = -200
        sm[("-", ch)] = -200
    sm.pop(("-", "-"), None)
    return sm


def _scores(match, mismatch, matrix):
    """score_matrix.rs:21-34"""
    if matrix in ("HOXD70", "HOXD70.mtx"):
        return score_matrix_from_matrix_file("HOXD70")
    if matrix in ("HOXD55", "HOXD55.mtx"):
        return score_matrix_from_matrix_file("HOXD55")
    assert matrix in (None, "none"), "wrong matrix type"
    return score_matrix_match_mis(match, -mismatch)


# ------------------------------------------------------------------------------------------------ graph
class PathGraph:
    pass


def create_path_graph(segs, paths):
    """pathwise_graph.rs:135-248 (is_reversed = false)"""
    g = PathGraph()
    lnz = ["$"]
    ids = [0]
    pos = {}
    for sid in sorted(segs):
        start = len(lnz)
        for ch in segs[sid]:
            lnz.append(ch)
            ids.append(sid)
        pos[sid] = (start, len(lnz) - 1)
    lnz.append("F")
    ids.append(0)
    n, P = len(lnz), len(paths)
    nwp = [False] * n
    pred = {}  # node -> {pred -> set(paths)}
    alphas = [P + 1] * n
    pn = [set() for _ in range(n)]
    pn[0] = set(range(P))
    alphas[0] = 0
    alphas[n - 1] = 0
    for pid, path in enumerate(paths):
        for k, sid in enumerate(path):
            hs, he = pos[sid]
            for idx in range(hs, he + 1):
                pn[idx].add(pid)
                if alphas[idx] == P + 1:
                    alphas[idx] = pid
            nwp[hs] = True
            if k == 0:
                pred.setdefault(hs, {}).setdefault(0, set()).add(pid)
            else:
                pe = pos[path[k - 1]][1]
                pred.setdefault(hs, {}).setdefault(pe, set()).add(pid)
                if k == len(path) - 1:
                    pred.setdefault(n - 1, {}).setdefault(he, set()).add(pid)
    nwp[n - 1] = True
    pn[n - 1] = set(range(P))
    g.lnz, g.nwp, g.pred, g.pn, g.alphas, g.P, g.ids = lnz, nwp, pred, pn, alphas, P, ids
    return g


def create_reverse_path_graph(fg):
    """pathwise_graph.rs:250-282"""
    g = PathGraph()
    n = len(fg.lnz)
    nwp = [False] * n
    pred = {}
    for node, ps in fg.pred.items():
        for p, paths in ps.items():
            nwp[p] = True
            for q in paths:
                pred.setdefault(p, {}).setdefault(node, set()).add(q)
    g.lnz, g.nwp, g.pred, g.pn, g.alphas, g.P, g.ids = fg.lnz, nwp, pred, fg.pn, fg.alphas, fg.P, fg.ids
    return g


def preds_and_paths(g, node):
    """PredHash::get_preds_and_paths; the reference unwrap()s a missing entry."""
    return sorted(g.pred[node].items())


def distance_from_start(rg):
    """pathwise_graph.rs:306-329 (called with the REVERSE graph)"""
    n = len(rg.lnz)
    r = [-1] * n
    r[0] = 0
    for p, _ in preds_and_paths(rg, 0):
        r[p] = 1
    for i in range(1, n - 1):
        if r[i] == -1 or r[i] > r[i - 1] + 1:
            r[i] = r[i - 1] + 1
        if rg.nwp[i]:
            for p, _ in preds_and_paths(rg, i):
                if r[p] == -1 or r[p] > r[i] + 1:
                    r[p] = r[i] + 1
    return r


def distance_from_end(g):
    """pathwise_graph.rs:330-354"""
    n = len(g.lnz)
    r = [-1] * n
    r[n - 1] = 0
    for p, _ in preds_and_paths(g, n - 1):
        r[p] = 1
    for i in range(n - 2, 0, -1):
        if r[i] == -1 or r[i] > r[i + 1] + 1:
            r[i] = r[i + 1] + 1
        if g.nwp[i]:
            for p, _ in preds_and_paths(g, i):
                if r[p] == -1 or r[p] > r[i] + 1:
                    r[p] = r[i] + 1
    return r


class Displacement:
    """pathwise_graph.rs:284-305 without materialising the n x n matrix"""

    def __init__(self, g, rg):
        self.dfe = distance_from_end(g)
        self.dfs = distance_from_start(rg)

    def at(self, i, j):
        if i == j:
            return 0
        return abs(self.dfs[i] - self.dfs[j]) + abs(self.dfe[i] - self.dfe[j])


# ------------------------------------------------------------------------------------------------ the delta-encoded DP
def _dp(seq, g, sm, rev, free_border, border_is_dp):
    """pathwise_alignment.rs:19-304 / pathwise_alignment_semiglobal.rs:19-225 (rev = False) and
    pathwise_alignment_recombination.rs:146-431 (rev = True). free_border: the border column is all zeros."""
    lnz, nwp, pn, alphas, P = g.lnz, g.nwp, g.pn, g.alphas, g.P
    n, L = len(lnz), len(seq)
    dpm = [[[0] * P for _ in range(L)] for _ in range(n)]
    di = 1 if rev else -1

    def border(i, j):  # (_, 0) forward / j == last_char_pos reverse
        ip = i + di
        if not nwp[i]:
            common = pn[i] & pn[ip]
            if alphas[ip] in common:
                for path in sorted(common):
                    if path == alphas[i]:
                        dpm[i][j][path] = dpm[ip][j][path] + sm[(lnz[i], "-")]
                    else:
                        dpm[i][j][path] = dpm[ip][j][path]
            else:
                dpm[i][j][alphas[i]] = dpm[ip][j][alphas[i]] + dpm[ip][j][alphas[ip]] + sm[(lnz[i], "-")]
                for path in sorted(common):
                    if path != alphas[i]:
                        dpm[i][j][path] = dpm[ip][j][path] - dpm[ip][j][alphas[i]]
        else:
            alphas_deltas = {}
            for p, p_paths in preds_and_paths(g, i):
                common = pn[i] & p_paths
                paths = sorted(common)
                if alphas[p] in common:
                    alphas_deltas[alphas[p]] = paths
                    dpm[i][j][alphas[p]] = dpm[p][j][alphas[p]] + sm[(lnz[i], "-")]
                    for path in paths:
                        if path != alphas[p]:
                            dpm[i][j][path] = dpm[p][j][path]
                else:
                    ta = alphas[i] if alphas[i] in common else paths[0]
                    alphas_deltas[ta] = paths
                    dpm[i][j][ta] = dpm[p][j][alphas[p]] + dpm[p][j][ta] + sm[(lnz[i], "-")]
                    for path in paths:
                        if path != ta:
                            dpm[i][j][path] = dpm[p][j][path] - dpm[p][j][ta]
            for a in sorted(alphas_deltas):
                if a != alphas[i]:
                    dpm[i][j][a] -= dpm[i][j][alphas[i]]
                    for path in alphas_deltas[a]:
                        if path != a:
                            dpm[i][j][path] += dpm[i][j][a]

    def general(i, j):
        ip, jp = i + di, j + di
        sub, gi_, gj_ = sm[(lnz[i], seq[j])], sm[(lnz[i], "-")], sm[(seq[j], "-")]
        if not nwp[i]:
            common = pn[i] & pn[ip]
            ai, ap = alphas[i], alphas[ip]
            if ap in common:
                u = dpm[ip][j][ap] + gi_
                d = dpm[ip][jp][ap] + sub
                l = dpm[i][jp][ai] + gj_
                best = max(d, u, l)
                dpm[i][j][ai] = best
                for path in common:
                    if path != ai:
                        if best == d:
                            dpm[i][j][path] = dpm[ip][jp][path]
                        elif best == u:
                            dpm[i][j][path] = dpm[ip][j][path]
                        else:
                            dpm[i][j][path] = dpm[i][jp][path]
            else:
                u = dpm[ip][j][ap] + dpm[ip][j][ai] + gi_
                d = dpm[ip][jp][ap] + dpm[ip][jp][ai] + sub
                l = dpm[i][jp][ai] + gj_
                best = max(d, u, l)
                dpm[i][j][ai] = best
                for path in common:
                    if path != ai:
                        if best == d:
                            dpm[i][j][path] = dpm[ip][jp][path] - dpm[ip][jp][ai]
                        elif best == u:
                            dpm[i][j][path] = dpm[ip][j][path] - dpm[ip][j][ai]
                        else:
                            dpm[i][j][path] = dpm[i][jp][path]
        else:
            alphas_deltas = {}
            ai = alphas[i]
            for p, p_paths in preds_and_paths(g, i):
                common = pn[i] & p_paths
                paths = sorted(common)
                ap = alphas[p]
                if ap in common:
                    alphas_deltas[ap] = paths
                    u = dpm[p][j][ap] + gi_
                    d = dpm[p][jp][ap] + sub
                    if ai == ap:
                        l = dpm[i][jp][ap] + gj_
                    else:
                        l = dpm[i][jp][ap] + dpm[i][jp][ai] + gj_
                    best = max(d, u, l)
                    dpm[i][j][ap] = best
                    for path in paths:
                        if path != ap:
                            if best == d:
                                dpm[i][j][path] = dpm[p][jp][path]
                            elif best == u:
                                dpm[i][j][path] = dpm[p][j][path]
                            elif ap == ai:
                                dpm[i][j][path] = dpm[i][jp][path]
                            else:
                                dpm[i][j][path] = dpm[i][jp][path] - dpm[i][jp][ap]
                else:
                    ta = ai if ai in common else paths[0]
                    alphas_deltas[ta] = paths
                    u = dpm[p][j][ap] + dpm[p][j][ta] + gi_
                    d = dpm[p][jp][ap] + dpm[p][jp][ta] + sub
                    if ai == ta:
                        l = dpm[i][jp][ta] + gj_
                    else:
                        l = dpm[i][jp][ta] + dpm[i][jp][ai] + gj_
                    best = max(d, u, l)
                    dpm[i][j][ta] = best
                    for path in paths:
                        if path != ta:
                            if best == d:
                                dpm[i][j][path] = dpm[p][jp][path] - dpm[p][jp][ta]
                            elif best == u:
                                dpm[i][j][path] = dpm[p][j][path] - dpm[p][j][ta]
                            elif ta == ai:
                                dpm[i][j][path] = dpm[i][jp][path]
                            else:
                                dpm[i][j][path] = dpm[i][jp][path] - dpm[i][jp][ta]
            for a in sorted(alphas_deltas):
                if a != ai:
                    dpm[i][j][a] -= dpm[i][j][ai]
                    for path in alphas_deltas[a]:
                        if path != a:
                            dpm[i][j][path] += dpm[i][j][a]

    if not rev:
        for i in range(0, n - 1):
            for j in range(0, L):
                if i == 0 and j == 0:
                    pass
                elif j == 0:
                    if not free_border:
                        border(i, j)
                elif i == 0:
                    a0 = alphas[0]
                    dpm[i][j][a0] = dpm[i][j - 1][a0] + sm[(seq[j], "-")]
                    for k in range(a0 + 1, P):
                        dpm[i][j][k] = dpm[i][j - 1][k]
                else:
                    general(i, j)
    else:
        last_i, last_j = n - 1, L - 1
        for i in range(last_i, 0, -1):
            for j in range(last_j, 0, -1):
                if i == last_i and j == last_j:
                    pass
                elif i == last_i:
                    a = alphas[i]
                    dpm[i][j][a] = dpm[i][j + 1][a] + sm[(seq[j], "-")]
                    for k in range(a + 1, P):
                        dpm[i][j][k] = dpm[i][j + 1][k]
                elif j == last_j:
                    if not free_border:
                        border(i, j)
                else:
                    general(i, j)
    return dpm


# ------------------------------------------------------------------------------------------------ output helpers
def build_cigar(cigar):
    """pathwise_alignment_output.rs:471-556"""
    out = ""
    d = u = l = mm = 0
    for ch in cigar:
        if ch == "D":
            if u:
                out += f"{u}I"
                u = 0
            if l:
                out += f"{l}D"
                l = 0
            if mm:
                out += f"{mm}X"
                mm = 0
            d += 1
        elif ch == "U":
            if d:
                out += f"{d}M"
                d = 0
            if l:
                out += f"{l}D"
                l = 0
            if mm:
                out += f"{mm}X"
                mm = 0
            u += 1
        elif ch == "d":
            if d:
                out += f"{d}M"
                d = 0
            if l:
                out += f"{l}D"
                l = 0
            if u:
                out += f"{u}I"
                u = 0
            mm += 1
        else:
            if d:
                out += f"{d}M"
                d = 0
            if u:
                out += f"{u}I"
                u = 0
            if mm:
                out += f"{mm}X"
                mm = 0
            l += 1
    if d:
        out += f"{d}M"
    if u:
        out += f"{u}I"
    if l:
        out += f"{l}D"
    if mm:
        out += f"{mm}X"
    return out


def dedup(v):
    out = []
    for x in v:
        if not out or out[-1] != x:
            out.append(x)
    return out


def get_path_len_start_end(ids, start, end, path_len):
    """utils.rs:221-254"""
    path_start = 0
    if start > 0:
        first = ids[start]
        counter = start - 1
        while counter > 0 and ids[counter] == first:
            counter -= 1
            path_start += 1
    path_end = path_start + path_len - 1 if path_len > 0 else 0
    end_offset = 0
    if end > 0:
        last = ids[end]
        counter = end + 1
        while counter < len(ids) - 1 and ids[counter] == last:
            counter += 1
            end_offset += 1
    return path_end + end_offset + 1, path_start, path_end


def get_rec_path_len_start_end(ids, fen, rsn, start, end, forw_len, rev_len):
    """utils.rs:256-323"""
    def back(row):
        c = 0
        if row > 0:
            first = ids[row]
            counter = row - 1
            while counter > 0 and ids[counter] == first:
                counter -= 1
                c += 1
        return c

    def fwd(row):
        c = 0
        if row > 0:
            last = ids[row]
            counter = row + 1
            while counter < len(ids) - 1 and ids[counter] == last:
                counter += 1
                c += 1
        return c

    path_start = back(start)
    forw_path_end = path_start + forw_len - 1 if forw_len > 0 else 0
    forw_path_len = forw_path_end + fwd(fen) + 1
    rev_path_start = back(rsn)
    rev_path_end = rev_path_start + rev_len - 1 if rev_len > 0 else 0
    path_end = forw_path_len + rev_path_end
    rev_path_len = rev_path_end + fwd(end) + 1
    return forw_path_len + rev_path_len, path_start, path_end


def get_node_offset(ids, node):
    """pathwise_alignment_recombination.rs:9-22"""
    h = ids[node]
    if h == 0:
        return 0
    counter, off = node, 0
    while ids[counter - 1] == h:
        counter -= 1
        off += 1
    return off


def gaf_string(name, qlen, qs, qe, strand, path, plen, ps, pe, residues, abl, mq, comments):
    """gaf_output.rs:70-94"""
    return "\t".join([name, str(qlen), str(qs), str(qe), strand, ">" + ">".join(str(x) for x in path), str(plen), str(ps),
                      str(pe), str(residues), abl, mq, comments])


def f32_display(v):
    """Rust `{}` of an f32: shortest decimal that round-trips, never scientific, integers without a fraction."""
    s = np.format_float_positional(F32(v), unique=True, trim="-")
    return s


# ------------------------------------------------------------------------------------------------ modes 4 / 5
def build_alignment(dpm, g, seq, sm, best_path, ending_node, global_align, name):
    """pathwise_alignment_output.rs:7-184"""
    lnz, alphas, nwp, ids = g.lnz, g.alphas, g.nwp, g.ids
    cigar, hia, pseq = [], [], []
    path_length = 0
    i, j = ending_node, len(dpm[ending_node]) - 1

    def absv(ii, jj):
        if alphas[ii] == best_path:
            return dpm[ii][jj][best_path]
        return dpm[ii][jj][best_path] + dpm[ii][jj][alphas[ii]]

    score = absv(i, j)
    while i > 0 and j > 0:
        predecessor = None
        if not nwp[i]:
            d = absv(i - 1, j - 1) + sm[(lnz[i], seq[j])]
            u = absv(i - 1, j) + sm[(lnz[i], "-")]
            l = absv(i, j - 1) + sm[("-", seq[j])]
        else:
            d = u = l = 0
            for pred, paths in preds_and_paths(g, i):
                if best_path in paths:
                    predecessor = pred
                    d = absv(pred, j - 1) + sm[(lnz[i], seq[j])]
                    u = absv(pred, j) + sm[(lnz[i], "-")]
                    l = absv(i, j - 1) + sm[("-", seq[j])]
        mx = max(d, u, l)
        if mx == d:
            cigar.append("d" if lnz[i] != seq[j] else "D")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i - 1 if predecessor is None else predecessor
            j -= 1
            path_length += 1
        elif mx == u:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i - 1 if predecessor is None else predecessor
            path_length += 1
        else:
            cigar.append("L")
            j -= 1
    while j > 0:
        cigar.append("L")
        j -= 1
    if global_align:
        while i > 0:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            path_length += 1
            if not nwp[i]:
                i = i - 1
            else:
                p = 0
                for pred, paths in preds_and_paths(g, i):
                    if best_path in paths:
                        p = pred
                i = p
    cigar.reverse()
    pseq.reverse()
    L = len(dpm[0])
    hia = dedup(hia)
    hia.reverse()
    plen, ps, pe = get_path_len_start_end(ids, i if i == 0 else i + 1, ending_node, path_length)
    comments = f"{build_cigar(cigar)}, best path: {best_path}, score: {score}\t{''.join(pseq)}"
    return gaf_string(name, L - 1, 0, L - 2, "+", hia, plen, ps, pe, 0, "*", "*", comments)


def mode4(seq, g, sm, name):
    """pathwise_alignment.rs:5-340"""
    dpm = _dp(seq, g, sm, False, False, True)
    n, L, P = len(g.lnz), len(seq), g.P
    ending = [0] * P
    results = [0] * P
    for pred, paths in preds_and_paths(g, n - 1):
        for path in sorted(paths):
            if path == g.alphas[pred]:
                results[path] = dpm[pred][L - 1][path]
            else:
                results[path] = dpm[pred][L - 1][path] + dpm[pred][L - 1][g.alphas[pred]]
            ending[path] = pred
    best = max((s, p) for p, s in enumerate(results))[1]
    return build_alignment(dpm, g, seq, sm, best, ending[best], True, name)


def mode5(seq, g, sm, name):
    """pathwise_alignment_semiglobal.rs:6-277"""
    dpm = _dp(seq, g, sm, False, True, True)
    n, L, P = len(g.lnz), len(seq), g.P
    mx, ending_node, chosen = None, 0, 0
    for i in range(1, n - 1):
        paths = g.pn[i]
        ab = list(dpm[i][L - 1])
        for path in sorted(paths):
            if path != g.alphas[i]:
                ab[path] = ab[path] + ab[g.alphas[i]]
        bp = None
        for path, score in enumerate(ab):
            if path in paths and (bp is None or bp[0] < score):
                bp = (score, path)
        if mx is None or bp[0] > mx:
            mx, ending_node, chosen = bp[0], i, bp[1]
    return build_alignment(dpm, g, seq, sm, chosen, ending_node, False, name)


# ------------------------------------------------------------------------------------------------ modes 8 / 9
def absolute_scores(dpm, g):
    """pathwise_alignment_recombination.rs:747-757 (the last row is left as it is)"""
    for i in range(len(dpm) - 1):
        a = g.alphas[i]
        for j in range(len(dpm[i])):
            row = dpm[i][j]
            for path in g.pn[i]:
                if path != a:
                    row[path] += row[a]


def get_rev_sequence(seq):
    return list(seq[1:]) + ["F"]


def best_alignment(m, w, displ, brc, mrc, mode, g, rbw):
    """pathwise_alignment_recombination.rs:759-873"""
    n, L, P = len(m), len(m[0]), g.P
    pn, ids = g.pn, g.ids
    mx, best_path = None, None
    if mode == 8:
        for pred, paths in preds_and_paths(g, n - 1):
            for path in sorted(paths):
                if mx is None or mx < m[pred][L - 1][path]:
                    mx, best_path = m[pred][L - 1][path], path
    else:
        for i in range(n - 1):
            for path in range(P):
                if path in pn[i]:
                    if mx is None or mx < m[i][L - 1][path]:
                        mx, best_path = m[i][L - 1][path], path
    curr = F32(mx)
    fbp = rbp = best_path
    onedge = False
    oob = max(int(F32(F32(L) * F32(F32(1.0) - F32(rbw))) / F32(2.0)), 1)
    fen = rsn = col = 0
    rec_pen = 0
    brc_f, mrc_f = F32(brc), F32(mrc)
    # The triple loop of the reference, with the innermost loop (over rev_i, ascending) evaluated as numpy float32 vectors:
    # every candidate is computed with the same three separately rounded f32 operations, and the sequential acceptance rule
    # is applied by jumping to the next index that satisfies it (each acceptance raises curr or sets onedge).
    nid = np.array(ids)
    inner = np.arange(1, n - 1)
    edge_ri = nid[1:n - 1] != nid[0:n - 2]
    dfs, dfe = np.array(displ.dfs), np.array(displ.dfe)
    for j in range(oob, L - oob):
        fp_of = [max((s, p) for p, s in enumerate(m[i][j]))[1] for i in range(n)]
        rp_of = [max((s, p) for p, s in enumerate(w[i][j]))[1] for i in range(n)]
        rp_arr = np.array(rp_of[1:n - 1])
        wv = np.array([w[ri][j][rp_of[ri]] for ri in range(1, n - 1)], dtype=np.int64)
        rmemb = np.array([rp_of[ri] in pn[ri] for ri in range(1, n - 1)])
        for i in range(1, n - 1):
            fp = fp_of[i]
            if fp not in pn[i]:
                continue
            ok = (nid[1:n - 1] != nid[i]) & (rp_arr != fp) & rmemb
            if not ok.any():
                continue
            dd = np.abs(dfs[i] - dfs[1:n - 1]) + np.abs(dfe[i] - dfe[1:n - 1])
            dd = np.where(inner == i, 0, dd)
            penalty = (brc_f + (mrc_f * dd.astype(F32)).astype(F32)).astype(F32)
            new = ((m[i][j][fp] + wv).astype(F32) - penalty).astype(F32)
            iedge = (i + 1 == n or ids[i] != ids[i + 1])
            edge = edge_ri & iedge
            pos = 0
            while True:
                cond = ok & ((new > curr) | ((new == curr) & (not onedge) & edge))
                cond[:pos] = False
                hit = np.flatnonzero(cond)
                if hit.size == 0:
                    break
                k = int(hit[0])
                onedge = bool(edge[k])
                curr = new[k]
                fen, rsn, fbp, rbp, col, rec_pen = i, k + 1, fp, int(rp_arr[k]), j, int(dd[k])
                pos = k + 1
    return fen, rsn, fbp, rbp, col, (curr, rec_pen)


def ending_node_of(dpm, best_path, g):
    """pathwise_alignment_recombination.rs:885-897"""
    best, node = None, 0
    L = len(dpm[0])
    for i in range(1, len(dpm) - 1):
        if best_path in g.pn[i]:
            if best is None or dpm[i][L - 1][best_path] > best:
                best, node = dpm[i][L - 1][best_path], i
    return node


def _walk_fwd(dpm, g, seq, sm, best_path, i, j, pad_global):
    """forward-matrix walk shared by the four builders (absolute scores)"""
    lnz, nwp, ids = g.lnz, g.nwp, g.ids
    cigar, hia, pseq = [], [], []
    plen = 0
    while i > 0 and j > 0:
        predecessor = None
        if not nwp[i]:
            d = dpm[i - 1][j - 1][best_path] + sm[(lnz[i], seq[j])]
            u = dpm[i - 1][j][best_path] + sm[(lnz[i], "-")]
            l = dpm[i][j - 1][best_path] + sm[("-", seq[j])]
        else:
            d = u = l = 0
            for pred, paths in preds_and_paths(g, i):
                if best_path in paths:
                    predecessor = pred
                    d = dpm[pred][j - 1][best_path] + sm[(lnz[i], seq[j])]
                    u = dpm[pred][j][best_path] + sm[(lnz[i], "-")]
                    l = dpm[i][j - 1][best_path] + sm[("-", seq[j])]
        mx = max(d, u, l)
        if mx == d:
            cigar.append("d" if lnz[i] != seq[j] else "D")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i - 1 if predecessor is None else predecessor
            j -= 1
            plen += 1
        elif mx == u:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i - 1 if predecessor is None else predecessor
            plen += 1
        else:
            cigar.append("L")
            j -= 1
    while j > 0:
        cigar.append("L")
        j -= 1
    if pad_global:
        while i > 0:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            predecessor = None
            if nwp[i]:
                for pred, paths in preds_and_paths(g, i):
                    if best_path in paths:
                        predecessor = pred
            i = i - 1 if predecessor is None else predecessor
            plen += 1
    return cigar, hia, pseq, plen, i


def gaf_no_rec(dpm, g, seq, sm, best_path, ending_node, glob, name):
    """recombination_output.rs:239-361 (semiglobal) / 633-782 (global)"""
    L = len(dpm[0])
    if glob:
        i = 0
        for node, paths in preds_and_paths(g, len(dpm) - 1):
            if best_path in paths:
                i = node
        ending_node = i
    i, j = ending_node, L - 1
    score = dpm[i][j][best_path]
    cigar, hia, pseq, plen, i = _walk_fwd(dpm, g, seq, sm, best_path, i, j, glob)
    cigar.reverse()
    pseq.reverse()
    hia = dedup(hia)
    hia.reverse()
    pl, ps, pe = get_path_len_start_end(g.ids, i if i == 0 else i + 1, ending_node, plen)
    comments = f"{build_cigar(cigar)}, best path: {best_path}, score: {score}\t{''.join(pseq)}"
    return gaf_string(name, L - 1, 0, L - 2, "+", hia, pl, ps, pe, 0, "*", "*", comments)


def gaf_rec(dpm, rdpm, g, rg, seq, sm, bp, rbp, fen, rsn, rec_col, best_score, glob, name):
    """recombination_output.rs:12-237 (semiglobal) / 363-631 (global)"""
    lnz, ids = g.lnz, g.ids
    n, L = len(dpm), len(dpm[0])
    cigar, hia, pseq = [], [], []
    rev_plen = 0
    i, j = rsn, rec_col
    rev_end = i
    r_seq = get_rev_sequence(seq)
    while i > 0 and i < n - 1 and j < L - 1:
        predecessor = None
        if not rg.nwp[i]:
            d = rdpm[i + 1][j + 1][rbp] + sm[(lnz[i], r_seq[j])]
            u = rdpm[i + 1][j][rbp] + sm[(lnz[i], "-")]
            l = rdpm[i][j + 1][rbp] + sm[("-", r_seq[j])]
        else:
            d = u = l = 0
            for pred, paths in preds_and_paths(rg, i):
                if rbp in paths:
                    predecessor = pred
                    d = rdpm[pred][j + 1][rbp] + sm[(lnz[i], r_seq[j])]
                    u = rdpm[pred][j][rbp] + sm[(lnz[i], "-")]
                    l = rdpm[i][j + 1][rbp] + sm[("-", r_seq[j])]
        mx = max(d, u, l)
        rev_end = i
        if mx == d:
            cigar.append("d" if lnz[i] != r_seq[j] else "D")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i + 1 if predecessor is None else predecessor
            j += 1
            rev_plen += 1
        elif mx == u:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            i = i + 1 if predecessor is None else predecessor
            rev_plen += 1
        else:
            cigar.append("L")
            j += 1
    while j < L - 1:
        cigar.append("L")
        j += 1
    if glob:
        while i < n - 1:
            cigar.append("U")
            hia.append(ids[i])
            pseq.append(lnz[i])
            predecessor = None
            if rg.nwp[i]:
                for pred, paths in preds_and_paths(rg, i):
                    if rbp in paths:
                        predecessor = pred
            i = i + 1 if predecessor is None else predecessor
            rev_plen += 1
    tc, th, tp, plen, i = _walk_fwd(dpm, g, seq, sm, bp, fen, rec_col, glob)
    rec_edge = (len(tp) - 1) % (1 << 64)   # usize arithmetic of a release build
    tc.reverse()
    tc += cigar
    th.reverse()
    th += hia
    th = dedup(th)
    tp.reverse()
    tp += pseq
    start = i if i == 0 else i + 1
    pl, ps, pe = get_rec_path_len_start_end(ids, fen, rsn, start, rev_end, plen, rev_plen)
    if bp == rbp:
        recomb = f"No recombination, best path: {bp}"
    else:
        recomb = (f"recombination path {bp} {rbp}, nodes {ids[fen]}[{get_node_offset(ids, fen)}] {ids[rsn]}[{get_node_offset(ids, rsn)}], "
                  f"score: {f32_display(best_score[0])}, displacement: {best_score[1]}\t{''.join(tp)}\t{rec_edge}")
    comments = f"{build_cigar(tc)}, {recomb}"
    return gaf_string(name, L - 1, 0, L - 2, "+", th, pl, ps, pe, 0, "*", "*", comments)


def mode89(mode, seq, g, rg, sm, brc, mrc, displ, rbw, name):
    """pathwise_alignment_recombination.rs:23-127"""
    free = mode == 9
    m = _dp(seq, g, sm, False, free, True)
    absolute_scores(m, g)
    w = _dp(get_rev_sequence(seq), rg, sm, True, free, True)
    absolute_scores(w, rg)
    fen, rsn, fbp, rbp, col, score = best_alignment(m, w, displ, brc, mrc, mode, g, rbw)
    if fbp == rbp:
        if mode == 8:
            return gaf_no_rec(m, g, seq, sm, fbp, None, True, name)
        return gaf_no_rec(m, g, seq, sm, fbp, ending_node_of(m, fbp, g), False, name)
    return gaf_rec(m, w, g, rg, seq, sm, fbp, rbp, fen, rsn, col, score, mode == 8, name)


# ------------------------------------------------------------------------------------------------ mode 2 (POA, affine, banded)
def read_gfa_links(text):
    """left neighbours of every segment in L-line order (handlegraph: handle_edges_iter(h, Left))"""
    left = {}
    for ln in text.splitlines():
        f = ln.split("\t")
        if f[0] == "L":
            a, b = int(f[1]), int(f[3])
            if a not in left.setdefault(b, []):
                left[b].append(a)
    return left


def create_graph_struct(segs, left):
    """graph.rs:31-123 (amb_mode = false)"""
    lnz = ["$"]
    pos = {}
    for sid in sorted(segs):
        start = len(lnz)
        lnz += list(segs[sid])
        pos[sid] = (start, len(lnz) - 1)
    last_nodes = {sid: pos[sid][1] for sid in segs}
    nwp = [False] * (len(lnz) + 1)
    pred = {}
    for sid in sorted(segs):
        hs = pos[sid][0]
        ln = left.get(sid, [])
        if not ln:
            nwp[hs] = True
            pred.setdefault(hs, []).append(0)
        for p in ln:
            last_nodes.pop(p, None)
            nwp[hs] = True
            pred.setdefault(hs, []).append(pos[p][1])
    lnz.append("F")
    nwp[len(lnz) - 1] = True
    for idx in sorted(last_nodes.values()):
        pred.setdefault(len(lnz) - 1, []).append(idx)
    hofp = {0: "-1"}
    cur = 0
    ids = sorted(segs)
    for i in range(1, len(nwp) - 1):
        if nwp[i]:
            cur += 1
        hofp[i] = str(ids[cur - 1])
    return lnz, nwp, pred, hofp


def set_r_values(nwp, pred, n):
    """utils.rs:103-126"""
    r = [-1] * n
    r[n - 1] = 0
    for p in pred[n - 1]:
        r[p] = 0
    for i in range(n - 2, 0, -1):
        if r[i] == -1 or r[i] > r[i + 1] + 1:
            r[i] = r[i + 1] + 1
        if nwp[i]:
            for p in pred[i]:
                if r[p] == -1 or r[p] > r[i] + 1:
                    r[p] = r[i] + 1
    return [x if x >= 0 else (1 << 64) - 1 for x in r]   # `as usize`


def _as_i32(v):
    v &= 0xffffffff
    return v - (1 << 32) if v >= (1 << 31) else v


def set_ampl_for_row(i, p_arr, r_val, bsp, seq_len, bta):
    """utils.rs:17-72 (simd_version = false)"""
    if i == 0:
        ms = me = 0
    elif not p_arr:
        ms = me = bsp[i - 1] + 1
    else:
        pl = min(bsp[p] for p in p_arr)
        pr = max(bsp[p] for p in p_arr)
        ms, me = pl + 1, pr + 1
    tmp = min(_as_i32(ms), _as_i32(_as_i32(seq_len) - _as_i32(r_val)) - _as_i32(bta))
    band_start = 0 if tmp < 0 else max(0, tmp)
    if seq_len > r_val:
        band_end = min(seq_len, max(me, seq_len - r_val) + bta)
    else:
        band_end = min(seq_len, me + bta)
    return band_start, band_end


class _Row(list):
    """a Rust Vec: an index below 0 (a wrapped usize) is out of bounds and panics; a Python list would wrap it around"""
    def __getitem__(self, k):
        if isinstance(k, int) and k < 0:
            raise IndexError("index out of bounds")
        return list.__getitem__(self, k)

    def __setitem__(self, k, v):
        if isinstance(k, int) and k < 0:
            raise IndexError("index out of bounds")
        list.__setitem__(self, k, v)


def cell(pred, d):
    """bitfield_path.rs:39-44: the predecessor is truncated to 16 bits"""
    return (pred & 0xffff, d)


def mode2_exec(seq, name, lnz, nwp, pred, sm, o, e, bta, hofp):
    """gap_global_abpoa.rs:11-250; returns (stdout text of the call, score)"""
    n, L = len(lnz), len(seq)
    m, x, y = [[] for _ in range(n)], [[] for _ in range(n)], [[] for _ in range(n)]
    path, path_x, path_y = [[] for _ in range(n)], [[] for _ in range(n)], [[] for _ in range(n)]
    r_values = set_r_values(nwp, pred, n)
    bsp = [0] * n
    ampl = [(0, 0)] * n

    def jpos(p, i, j):
        lp, li = ampl[p][0], ampl[i][0]
        return j + (li - lp) if lp < li else j - (lp - li)

    for i in range(n - 1):
        p_arr = pred[i] if nwp[i] else []
        left, right = set_ampl_for_row(i, p_arr, r_values[i], bsp, L, bta)
        ampl[i] = (left, right)
        W = right - left
        if W < 0:
            raise RuntimeError("vec![0; right - left] with a wrapped length: capacity overflow")
        # W == 0: a row without cells is legal (gap_global_abpoa.rs:59-67); whoever indexes it later panics
        m[i], x[i], y[i] = _Row([0] * W), _Row([0] * W), _Row([0] * W)
        path[i], path_x[i], path_y[i] = _Row([(0, "O")] * W), _Row([(0, "O")] * W), _Row([(0, "O")] * W)
        best = 0
        for j in range(W):
            if i == 0 and j == 0:
                m[i][j] = 0
                path[i][j] = cell(0, "O")
            elif i == 0:
                y[i][j] = o + e * (j + left)
                m[i][j] = y[i][j]
                path[i][j] = cell(i, "L")
            elif j == 0 and left == 0:
                best_p = i - 1 if not nwp[i] else min(pred[i])
                x[i][j] = o + e * (best_p + 1)
                m[i][j] = x[i][j]
                path[i][j] = cell(best_p, "U")
            else:
                pa = pred[i] if nwp[i] else [i - 1]
                best_p = i - 1 if not nwp[i] else min(pred[i])
                # l
                if j > 0:
                    l_x, l_m = x[i][j - 1], m[i][j - 1] + o
                    if l_x > l_m:
                        x[i][j] = l_x + e
                        path_x[i][j] = cell(i, "X")
                    else:
                        x[i][j] = l_m + e
                    l_pred = i
                else:
                    x[i][j] = 2 * o + e * (best_p + 1) + e * (j + left)
                    l_pred = best_p
                # u
                first = True
                u_m = u_y = um_i = uy_i = 0
                for p in pa:
                    if ampl[p][0] <= j + left < ampl[p][1]:
                        jp = jpos(p, i, j)
                        cm, cy = m[p][jp] + o, y[p][jp]
                        if first:
                            first = False
                            u_m, u_y, um_i, uy_i = cm, cy, p, p
                        if cm > u_m:
                            u_m, um_i = cm, p
                        if cy > u_y:
                            u_y, uy_i = cy, p
                if first:
                    y[i][j] = 2 * o + e * (best_p + 1) + e * (j + left)
                    u_pred = best_p
                elif u_y > u_m:
                    y[i][j] = u_y + e
                    u_pred = uy_i
                    path_y[i][j] = cell(uy_i, "Y")
                else:
                    y[i][j] = u_m + e
                    u_pred = um_i
                # d
                first = True
                d = d_idx = 0
                for p in pa:
                    if ampl[p][0] < j + left <= ampl[p][1]:
                        cd = m[p][jpos(p, i, j) - 1]
                        if first:
                            d, d_idx, first = cd, p, False
                        if cd > d:
                            d, d_idx = cd, p
                l, u = x[i][j], y[i][j]
                if not first:
                    d += sm[(lnz[i], seq[j + left])]
                    if d < l:
                        if l < u:
                            if u_pred == 0:
                                raise RuntimeError("set_path_cell(u_pred, 'u'): impossible direction char (reference panic)")
                            path[i][j] = cell(u_pred, "U")
                            m[i][j] = u
                        else:
                            path[i][j] = cell(l_pred, "L")
                            m[i][j] = l
                    elif d < u:
                        path[i][j] = cell(u_pred, "U")
                        m[i][j] = u
                    else:
                        path[i][j] = cell(d_idx, "D" if lnz[i] == seq[j + left] else "d")
                        m[i][j] = d
                elif l < u:
                    path[i][j] = cell(u_pred, "U")
                    m[i][j] = u
                else:
                    path[i][j] = cell(l_pred, "L")
                    m[i][j] = l
            if m[i][j] >= m[i][best]:
                best = j
        bsp[i] = best + left
    last_row = n - 2
    last_col = len(m[last_row]) - 1
    for p in pred[n - 1]:
        t = (ampl[p][1] - ampl[p][0]) - 1
        if m[p][t] > m[last_row][last_col]:
            last_row, last_col = p, t
    best_value = m[last_row][last_col]
    out = ""
    if not band_ampl_enough(path, path_x, path_y, last_row, last_col, ampl, L):
        out += "Band length probably too short, maybe try with larger b and f\n"
    out += gaf_of_gap_abpoa(path, path_x, path_y, seq, name, ampl, last_row, last_col, hofp) + "\n"
    return out, best_value


def band_ampl_enough(path, path_x, path_y, i, j, ampl, L):
    """gap_global_abpoa.rs:371-455"""
    def jp(p, row, col):
        lp, lr = ampl[p][0], ampl[row][0]
        return col + (lr - lp) if lp < lr else col - (lp - lr)

    while path[i][j][1] != "O":
        left, right = ampl[i]
        if i == 0 or (j == 0 and left == 0):
            return True
        if (j == 0 and left != 0) or (j == right - left - 1 and right != L):
            return False
        pred, d = path[i][j]
        if d in ("D", "d"):
            j = jp(pred, i, j) - 1
            i = pred
        elif d == "L":
            if path_x[i][j][1] == "X":
                while path_x[i][j][1] == "X" and j > 0:
                    j -= 1
            else:
                j -= 1
        elif d == "U":
            if path_y[i][j][1] == "Y":
                while path_y[i][j][1] == "Y":
                    p = path_y[i][j][0]
                    j = jp(p, i, j)
                    i = p
            else:
                j = jp(pred, i, j)
                i = pred
        else:
            return False
    return True


def set_cigar_substring(cm, ci, cd, cs):
    """gaf_output.rs:876-892"""
    assert cm * ci + ci * cd + cm * cd == 0, "wrong format in cigar string"
    if cm > 0:
        return f"{cm}M{cs}"
    if ci > 0:
        return f"{ci}I{cs}"
    if cd > 0:
        return f"{cd}D{cs}"
    return cs


def node_start(hofp, row):
    """gaf_output.rs:867-874"""
    h = hofp[row]
    i = row
    while hofp[i] == h and i > 0:
        i -= 1
    return row - i


def gaf_of_gap_abpoa(path, path_x, path_y, seq, name, ampl, last_row, last_col, hofp):
    """gaf_output.rs:96-253 (amb_mode = false)"""
    col, row = last_col, last_row
    hia, cigars = [], []
    cigar = ""
    cm = ci = cd = 0
    curr_handle, last_dir = "", " "
    path_length = residues = 0
    while path[row][col][1] != "O":
        pred, d = path[row][col]
        if hofp[row] != curr_handle:
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cigars.insert(0, cigar)
            cigar = ""
            cm = ci = cd = 0
        curr_handle = hofp[row]
        if d.upper() != last_dir.upper():
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cm = ci = cd = 0
        last_dir = d
        p_left = ampl[pred][0]
        if ampl[row][0] < p_left:
            j_pos = col - (p_left - ampl[row][0])
        else:
            j_pos = col + (ampl[row][0] - p_left)
        if d == "D":
            hia.append(hofp[row])
            row, col = pred, j_pos - 1
            cm += 1
            path_length += 1
            residues += 1
        elif d == "d":
            hia.append(hofp[row])
            row, col = pred, j_pos - 1
            cm += 1
            path_length += 1
        elif d == "L":
            if path_x[row][col][1] == "X":
                while path_x[row][col][1] == "X":
                    cd += 1
                    col -= 1
            else:
                cd += 1
                col -= 1
        elif d == "U":
            if path_y[row][col][1] == "Y":
                while path_y[row][col][1] == "Y":
                    lr = ampl[row][0]
                    p = path_y[row][col][0]
                    lp = ampl[p][0]
                    jp = col + (lr - lp) if lp < lr else col - (lp - lr)
                    hia.append(hofp[row])
                    ci += 1
                    path_length += 1
                    col, row = jp, p
            else:
                hia.append(hofp[row])
                ci += 1
                path_length += 1
                row, col = pred, j_pos
        else:
            raise RuntimeError("impossible value in poa path")
    cigar = set_cigar_substring(cm, ci, cd, cigar)
    cigars.insert(0, cigar)
    hia = dedup(hia)
    hia.reverse()
    comments = ",".join(cigars[:-1])
    return gaf_string(name, len(seq) - 1, col, last_col + ampl[last_row][0], _STRAND[0], [int(h) for h in hia], path_length,
                      node_start(hofp, row), node_start(hofp, last_row), residues, "*", "*", comments)


def run_mode2(fasta_text, gfa_text, match=2, mismatch=4, gap_open=4, gap_ext=2, extra_b=1, extra_f=0.01, max_reads=None, matrix=None):
    """main.rs:171-214 without -s; returns stdout"""
    seqs, names = read_fasta(fasta_text)
    segs, _paths = read_gfa(gfa_text)
    lnz, nwp, pred, hofp = create_graph_struct(segs, read_gfa_links(gfa_text))
    sm = _scores(match, mismatch, matrix)
    out = ""
    for k, seq in enumerate(seqs):
        if max_reads is not None and k >= max_reads:
            break
        v = F32(F32(extra_b) + F32(F32(extra_f) * F32(len(seq))))
        bta = 0 if not (v > 0) else int(v)
        text, _score = mode2_exec(seq, names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext, bta, hofp)
        out += text
    return out


# ------------------------------------------------------------------------------------------------ modes 0 / 1 / 3
def set_left_right_x64(left, right, seq_len):
    """utils.rs:74-99. usize arithmetic of a RELEASE build (`cargo build --release`, no overflow checks: README, Cargo.toml):
    `new_right - 1` with new_right = 0 wraps, and only its residue modulo 8 is used — Python's `%` of a negative number gives
    the same residue, 2^64 being a multiple of 8."""
    nr, nl = right, left
    while (nr - nl) % 8 != 0:
        if (nr - nl) % 2 == 0 and nr < seq_len:
            nr += 1
        elif nl > 0:
            nl -= 1
        else:
            break
    if nl == 0:
        while (nr - 1) % 8 != 0 and nr < seq_len:
            nr += 1
    if nr == seq_len:
        while (nr - nl) % 8 != 0 and nl > 1:
            nl -= 1
    return nl, nr


def f32_scores(sm):
    """score_matrix.rs:10-17"""
    return {k: F32(v) for k, v in sm.items()}


def _split_path_value(val):
    """`val.to_string().split('.')`, integer part -> predecessor, fraction digits -> direction (gaf_output.rs:668-672)"""
    parts = f32_display(val).split(".")
    if len(parts) < 2:
        raise RuntimeError("index out of bounds: the len is 1 but the index is 1")
    return int(parts[0]), int(parts[1])


def mode0_exec_simd(read, name, lnz, nwp, pred, sm, bta, hofp, r_values):
    """global_abpoa.rs:10-257 (exec_simd, amb_mode = false), lane by lane; returns (stdout text, score)"""
    n, L = len(lnz), len(read)
    min_score = F32(F32(2.0) * F32(L)) * sm[(read[1], "-")]
    m = np.full((n, L), min_score, dtype=F32)
    path = np.full((n, L), -1.0, dtype=F32)
    bsp = [0] * n
    m[0][0] = 0.0
    path[0][0] = 0.0
    for i in range(1, n - 1):
        if not nwp[i]:
            m[i][0] = m[i - 1][0] + sm[(lnz[i], "-")]
            path[i][0] = F32(i - 1) + F32(0.2)
        else:
            best_p = min(pred[i])
            m[i][0] = m[best_p][0] + sm[(lnz[i], "-")]
            path[i][0] = F32(best_p) + F32(0.2)
    left, right = set_left_right_x64(*set_ampl_for_row(0, [], r_values[0], bsp, L, bta), L)
    for j in range(1, right):
        m[0][j] = m[0][j - 1] + sm[(read[j], "-")]
        path[0][j] = F32(0.3)
    for i in range(1, n - 1):
        p_arr = pred[i] if nwp[i] else []
        left, right = set_left_right_x64(*set_ampl_for_row(i, p_arr, r_values[i], bsp, L, bta), L)
        best_col = left
        start = 1 if left == 0 else left
        if right == L:
            if right < start:
                raise RuntimeError("attempt to subtract with overflow")
            end = ((right - start) // 8) * 8 + start
        else:
            end = right
        us_update = sm[(lnz[i], "-")]
        for j in range(start, end, 8):
            if j + 7 >= L:
                raise RuntimeError("index out of bounds (read[j + 7])")
            res, pth = [], []
            for k in range(8):      # the eight lanes; every input is a finished row
                idx = j + k
                ds_update = sm[(lnz[i], read[idx])]
                if not nwp[i]:
                    us = m[i - 1][idx] + us_update
                    ds = m[i - 1][idx - 1] + ds_update
                    take_d = ds > us
                    res.append(ds if take_d else us)
                    pth.append(F32(i - 1) + (F32(0.1) if take_d else F32(0.2)))
                else:
                    preds = pred[i]
                    best_us, best_ds = m[preds[0]][idx], m[preds[0]][idx - 1]
                    pbu = pbd = F32(preds[0])
                    for q in preds[1:]:
                        if m[q][idx] > best_us:
                            best_us, pbu = m[q][idx], F32(q)
                        if m[q][idx - 1] > best_ds:
                            best_ds, pbd = m[q][idx - 1], F32(q)
                    best_us = best_us + us_update
                    best_ds = best_ds + ds_update
                    take_d = best_ds > best_us
                    res.append(best_ds if take_d else best_us)
                    pth.append(pbd + F32(0.1) if take_d else pbu + F32(0.2))
            for k in range(8):
                m[i][j + k] = res[k]
                path[i][j + k] = pth[k]
            for idx in range(j, j + 8):
                l = m[i][idx - 1] + sm[(read[j], "-")]     # read[j], not read[idx] (global_abpoa.rs:157)
                if l > m[i][idx]:
                    m[i][idx] = l
                    path[i][idx] = F32(i) + F32(0.3)
                if m[i][idx] >= m[i][best_col]:
                    best_col = idx
        if end < right:
            for j in range(end, right):
                if not nwp[i]:
                    l = m[i][j - 1] + sm[(read[j], "-")]
                    u = m[i - 1][j] + sm[(lnz[i], "-")]
                    d = m[i - 1][j - 1] + sm[(lnz[i], read[j])]
                    m[i][j] = max(l, u, d)
                    if m[i][j] == d:
                        path[i][j] = F32(i - 1) + F32(0.1)
                    elif m[i][j] == u:
                        path[i][j] = F32(i - 1) + F32(0.2)
                    else:
                        path[i][j] = F32(i) + F32(0.3)
                else:
                    u = d = F32(0)
                    u_pred = d_pred = 0
                    first = True
                    for q in pred[i]:
                        if first:
                            u, d, u_pred, d_pred, first = m[q][j], m[q][j - 1], q, q, False
                        if m[q][j] > u:
                            u, u_pred = m[q][j], q
                        if m[q][j - 1] > d:
                            d, d_pred = m[q][j - 1], q
                    u = u + sm[(lnz[i], "-")]
                    d = d + sm[(read[j], lnz[i])]
                    l = m[i][j - 1] + sm[(read[j], "-")]
                    m[i][j] = max(l, u, d)
                    if m[i][j] == d:
                        path[i][j] = F32(d_pred) + F32(0.1)
                    elif m[i][j] == u:
                        path[i][j] = F32(u_pred) + F32(0.2)
                    else:
                        path[i][j] = F32(i) + F32(0.3)
                if m[i][j] >= m[i][best_col]:
                    best_col = j
        bsp[i] = best_col
    best_result, last_row, first = F32(0), 0, True
    for q in pred[n - 1]:
        if first:
            best_result, last_row, first = m[q][L - 1], q, False
        if m[q][L - 1] > best_result:
            best_result, last_row = m[q][L - 1], q
    return gaf_of_global_abpoa_simd(path, read, name, last_row, L - 1, hofp, lnz, best_result), int(best_result)


def gaf_of_global_abpoa_simd(path, seq, name, last_row, last_col, hofp, lnz, best_score):
    """gaf_output.rs:753-865 (amb_mode = false); returns the stdout text of the read"""
    col, row = last_col, last_row
    hia, cigar, path_sequence = [], [], []
    path_length = residues = 0
    out_ok = True
    while path[row][col] != 0.0:
        val = path[row][col]
        if val == F32(-1.0):
            out_ok = False
            break
        prd, d = _split_path_value(val)
        if d == 1:
            hia.append(hofp[row])
            path_sequence.append(lnz[row])
            row = prd
            col -= 1
            cigar.append("D" if lnz[row] == seq[col] else "d")   # compared AFTER the move (gaf_output.rs:793)
            path_length += 1
            residues += 1
        elif d == 3:
            col -= 1
            cigar.append("L")
        elif d == 2:
            hia.append(hofp[row])
            path_sequence.append(lnz[row])
            row = prd
            cigar.append("U")
            path_length += 1
        else:
            raise RuntimeError("impossible value in poa path")
    if not out_ok:
        return "band not enough for correct output\n" + gaf_string("", 0, 0, 0, " ", [0], 0, 0, 0, 0, "", "", "") + "\n"
    cigar.reverse()
    cigar_out = build_cigar(cigar)
    path_sequence.reverse()
    hia = dedup(hia)
    hia.reverse()
    comments = f"{cigar_out}, score: {f32_display(best_score)}\t{''.join(path_sequence)}"
    return gaf_string(name, len(seq) - 1, col, last_col, _STRAND[0], [int(x) for x in hia], path_length, node_start(hofp, row),
                      node_start(hofp, last_row), residues, "*", "*", comments) + "\n"


def mode1_exec_simd(read, name, lnz, nwp, pred, sm, hofp):
    """local_poa.rs:10-179 (exec_simd, amb_mode = false), lane by lane; returns (stdout text, score as i32)"""
    n, L = len(lnz), len(read)
    m = np.zeros((n, L), dtype=F32)
    path = np.zeros((n, L), dtype=F32)
    if L % 8 != 0:
        max_multiple = (L // 8) * 8
    else:
        if L < 8:
            raise RuntimeError("attempt to subtract with overflow")
        max_multiple = L - 8
    best_row = best_col = 0
    for i in range(1, n - 1):
        us_update = sm[(lnz[i], "-")]
        for j in range(1, max_multiple + 1, 8):
            if j + 7 >= L:
                raise RuntimeError("index out of bounds (read[j + 7])")
            res, pth = [], []
            for k in range(8):
                idx = j + k
                ds_update = sm[(lnz[i], read[idx])]
                if not nwp[i]:
                    us = m[i - 1][idx] + us_update
                    ds = m[i - 1][idx - 1] + ds_update
                    take_d = ds > us
                    res.append(ds if take_d else us)
                    pth.append(F32(i - 1) + (F32(0.1) if take_d else F32(0.2)))
                else:
                    preds = pred[i]
                    best_us, best_ds = m[preds[0]][idx], m[preds[0]][idx - 1]
                    pbu = pbd = F32(preds[0])
                    for q in preds[1:]:
                        if m[q][idx] > best_us:
                            best_us, pbu = m[q][idx], F32(q)
                        if m[q][idx - 1] > best_ds:
                            best_ds, pbd = m[q][idx - 1], F32(q)
                    best_us = best_us + us_update
                    best_ds = best_ds + ds_update
                    take_d = best_ds > best_us
                    res.append(best_ds if take_d else best_us)
                    pth.append(pbd + F32(0.1) if take_d else pbu + F32(0.2))
            for k in range(8):
                m[i][j + k] = res[k]
                path[i][j + k] = pth[k]
            for idx in range(j, min(j + 8, L)):
                l = m[i][idx - 1] + sm[(read[j], "-")]
                if l > m[i][idx]:
                    m[i][idx] = l
                    path[i][idx] = F32(i) + F32(0.3)
                if m[i][idx] <= 0.0:
                    m[i][idx] = 0.0
                    path[i][idx] = 0.0
                if m[i][idx] >= m[best_row][best_col]:
                    best_row, best_col = i, idx
        for j in range(max_multiple + 1, L):
            if not nwp[i]:
                l = m[i][j - 1] + sm[(read[j], "-")]
                u = m[i - 1][j] + sm[(lnz[i], "-")]
                d = m[i - 1][j - 1] + sm[(lnz[i], read[j])]
                m[i][j] = max(l, u, d)
                if m[i][j] < 0.0:
                    m[i][j] = 0.0
                    path[i][j] = 0.0
                elif m[i][j] == d:
                    path[i][j] = F32(i - 1) + F32(0.1)
                elif m[i][j] == u:
                    path[i][j] = F32(i - 1) + F32(0.2)
                else:
                    path[i][j] = F32(i) + F32(0.3)
            else:
                u = d = F32(0)
                u_pred = d_pred = 0
                first = True
                for q in pred[i]:
                    if first:
                        u, d, u_pred, d_pred, first = m[q][j], m[q][j - 1], q, q, False
                    if m[q][j] > u:
                        u, u_pred = m[q][j], q
                    if m[q][j - 1] > d:
                        d, d_pred = m[q][j - 1], q
                u = u + sm[(lnz[i], "-")]
                d = d + sm[(read[j], lnz[i])]
                l = m[i][j - 1] + sm[(read[j], "-")]
                m[i][j] = max(l, u, d)        # no clamp at 0 on this branch (local_poa.rs:149-160)
                if m[i][j] == d:
                    path[i][j] = F32(d_pred) + F32(0.1)
                elif m[i][j] == u:
                    path[i][j] = F32(u_pred) + F32(0.2)
                else:
                    path[i][j] = F32(i) + F32(0.3)
            if m[i][j] >= m[best_row][best_col]:
                best_row, best_col = i, j
    return gaf_of_local_poa_simd(path, read, name, best_row, best_col, hofp), int(m[best_row][best_col])


def _segment_cigar_walk(seq, name, last_row, last_col, hofp, step):
    """The shared frame of gaf_of_local_poa_simd (gaf_output.rs:639-751) and gaf_of_gap_local_poa (:502-637): one CIGAR
    per segment, flushed when the handle or the kind of move changes. `step(row, col)` returns None at the origin cell or
    (kind, new_row, new_col, d_count, i_count, m_count, handle_pushes, path_len, residues)."""
    col, row = last_col, last_row
    hia, cigars = [], []
    cigar = ""
    cm = ci = cd = 0
    curr_handle, last_dir = "", None
    path_length = residues = 0
    while True:
        st = step(row, col)
        if st is None:
            break
        kind = st[0]
        if hofp[row] != curr_handle:
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cigars.insert(0, cigar)
            cigar = ""
            cm = ci = cd = 0
        curr_handle = hofp[row]
        if kind != last_dir:
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cm = ci = cd = 0
        last_dir = kind
        _k, nrow, ncol, dd, di, dm, pushes, plen, res = st
        hia += pushes
        row, col = nrow, ncol
        cd += dd
        ci += di
        cm += dm
        path_length += plen
        residues += res
    cigar = set_cigar_substring(cm, ci, cd, cigar)
    cigars.insert(0, cigar)
    hia = dedup(hia)
    hia.reverse()
    comments = ",".join(cigars[:len(cigars) - 1])
    return gaf_string(name, len(seq) - 1, col, last_col, _STRAND[0], [int(x) for x in hia], path_length, node_start(hofp, row),
                      node_start(hofp, last_row), residues, "*", "*", comments) + "\n"


def gaf_of_local_poa_simd(path, seq, name, last_row, last_col, hofp):
    """gaf_output.rs:639-751 (amb_mode = false)"""
    def step(row, col):
        val = path[row][col]
        if val == 0.0:
            return None
        prd, d = _split_path_value(val)
        if d == 1:
            return (1, prd, col - 1, 0, 0, 1, [hofp[row]], 1, 1)
        if d == 3:
            return (3, row, col - 1, 1, 0, 0, [], 0, 0)
        if d == 2:
            return (2, prd, col, 0, 1, 0, [hofp[row]], 1, 0)
        raise RuntimeError("impossible value in poa path")
    return _segment_cigar_walk(seq, name, last_row, last_col, hofp, step)


def mode3_exec(seq, name, lnz, nwp, pred, sm, o, e, hofp):
    """gap_local_poa.rs:8-187 (amb_mode = false); returns (stdout text, score)"""
    n, L = len(lnz), len(seq)
    m = [[0] * L for _ in range(n)]
    x = [[0] * L for _ in range(n)]
    y = [[0] * L for _ in range(n)]
    O = cell(0, "O")
    path = [[O] * L for _ in range(n)]     # rows the loops never reach keep bitvec![0; 32] = (0, 'O')
    path_x = [[O] * L for _ in range(n)]
    path_y = [[O] * L for _ in range(n)]
    best_row = best_col = 0
    for i in range(n - 1):
        for j in range(L):
            if i == 0 or j == 0:
                path[i][j] = path_x[i][j] = path_y[i][j] = O
            else:
                l_x = x[i][j - 1] + e
                l_m = m[i][j - 1] + o + e
                if l_x > l_m:
                    path_x[i][j] = cell(i, "X")
                    l = l_x
                else:
                    path_x[i][j] = cell(i, "M")
                    l = l_m
                x[i][j] = l
                if not nwp[i]:
                    d = m[i - 1][j - 1] + sm[(seq[j], lnz[i])]
                    d_idx = u_idx = i - 1
                    u_y = y[i - 1][j] + e
                    u_m = m[i - 1][j] + o + e
                    if u_y > u_m:
                        path_y[i][j] = cell(u_idx, "Y")
                        u = u_y
                    else:
                        path_y[i][j] = cell(u_idx, "M")
                        u = u_m
                    y[i][j] = u
                else:
                    # get_best_d / get_best_u start from 0 / predecessor 0: their `first` flag is initialised to false
                    # (gap_local_poa.rs:134,163), so the first predecessor never seeds the maximum
                    d, d_idx = 0, 0
                    for q in pred[i]:
                        if m[q][j - 1] > d:
                            d, d_idx = m[q][j - 1], q
                    um, uy, um_idx, uy_idx = 0, 0, 0, 0
                    for q in pred[i]:
                        if m[q][j] + o > um:
                            um, um_idx = m[q][j] + o, q
                        if y[q][j] > uy:
                            uy, uy_idx = y[q][j], q
                    if um > uy:
                        u, u_idx, from_m = um, um_idx, True
                    else:
                        u, u_idx, from_m = uy, uy_idx, False
                    d += sm[(seq[j], lnz[i])]
                    u += e
                    y[i][j] = u
                    path_y[i][j] = cell(u_idx, "M" if from_m else "Y")
                if d < 0 and l < 0 and u < 0:
                    m[i][j] = 0
                    path[i][j] = O
                else:
                    # utils::get_max_d_u_l (utils.rs:129-140)
                    if d < u:
                        best_val, dr = (l, "L") if u < l else (u, "U")
                    else:
                        best_val, dr = (l, "L") if d < l else (d, "D")
                    if dr == "D" and lnz[i] != seq[j]:
                        dr = "d"
                    m[i][j] = best_val
                    path[i][j] = cell(d_idx if dr in "Dd" else (u_idx if dr == "U" else i), dr)
            if m[i][j] > m[best_row][best_col]:
                best_row, best_col = i, j
    return gaf_of_gap_local_poa(path, path_x, path_y, seq, name, best_row, best_col, hofp), m[best_row][best_col]


def gaf_of_gap_local_poa(path, path_x, path_y, seq, name, last_row, last_col, hofp):
    """gaf_output.rs:502-637 (amb_mode = false). The X / Y chains are walked inside ONE step of the outer loop."""
    def step(row, col):
        prd, d = path[row][col]
        if d == "O":
            return None
        if d == "D":
            return ("D", prd, col - 1, 0, 0, 1, [hofp[row]], 1, 1)
        if d == "d":
            return ("D", prd, col - 1, 0, 0, 1, [hofp[row]], 1, 0)      # compared upper-cased (:539)
        if d == "L":
            if path_x[row][col][1] == "X":
                cnt = 0
                while path_x[row][col][1] == "X":
                    cnt += 1
                    col -= 1
                return ("L", row, col, cnt, 0, 0, [], 0, 0)
            return ("L", row, col - 1, 1, 0, 0, [], 0, 0)
        if d == "U":
            if path_y[row][col][1] == "Y":
                pushes, cnt = [], 0
                while path_y[row][col][1] == "Y":
                    q = path_y[row][col][0]
                    pushes.append(hofp[row])
                    row = q
                    cnt += 1
                return ("U", row, col, 0, cnt, 0, pushes, cnt, 0)
            return ("U", prd, col, 0, 1, 0, [hofp[row]], 1, 0)
        raise RuntimeError("impossible value in poa path")
    return _segment_cigar_walk(seq, name, last_row, last_col, hofp, step)


def _at(row, idx):
    """checked indexing (a Rust Vec panics where a Python list would wrap around)"""
    if idx < 0 or idx >= len(row):
        raise RuntimeError("index out of bounds")
    return row[idx]


def mode0_exec_scalar(seq, name, lnz, nwp, pred, sm, bta, hofp):
    """global_abpoa.rs:260-427 (the scalar exec: what the -s retry of mode 0 runs); returns (stdout text of the call, score)"""
    n, L = len(lnz), len(seq)
    r_values = set_r_values(nwp, pred, n)
    bsp = [0] * n
    m, path = [[] for _ in range(n)], [[] for _ in range(n)]
    ampl = [(0, 0)] * n

    def jpos(p, i, j):
        lp, li = ampl[p][0], ampl[i][0]
        return j + (li - lp) if lp < li else j - (lp - li)

    def min_pred(i):
        return i - 1 if not nwp[i] else min(pred[i])

    for i in range(n - 1):
        p_arr0 = pred[i] if nwp[i] else []
        left, right = set_ampl_for_row(i, p_arr0, r_values[i], bsp, L, bta)
        ampl[i] = (left, right)
        if right < left:
            raise RuntimeError("attempt to subtract with overflow")
        w = right - left
        m[i] = [0] * w
        path[i] = [cell(0, "O")] * w
        best_val_pos = 0
        for j in range(w):
            if i == 0 and j == 0:
                m[i][j] = 0
                path[i][j] = cell(0, "O")
            elif i == 0:
                m[i][j] = m[i][j - 1] + sm[("-", seq[j + left])]
                path[i][j] = cell(i, "L")
            elif j == 0 and left == 0:
                best_p = min_pred(i)
                m[i][j] = _at(m[best_p], j) + sm[("-", lnz[i])]
                path[i][j] = cell(best_p, "U")
            else:
                p_arr = pred[i] if nwp[i] else [i - 1]
                if j > 0:      # get_best_l
                    l, l_pred = m[i][j - 1] + sm[(seq[j + left], "-")], i
                else:
                    l, l_pred = sm[(seq[j + left], "-")] * (i + left + j), min_pred(i)
                # get_best_u (global_abpoa.rs:529-566)
                u = u_idx = None
                for q in p_arr:
                    if ampl[q][0] <= j + left < ampl[q][1]:
                        cu = _at(m[q], jpos(q, i, j))
                        if u is None or cu > u:
                            u, u_idx = cu, q
                if u is None:
                    u, u_pred = sm[(lnz[i], "-")] * (i + left + j), min_pred(i)
                else:
                    u, u_pred = u + sm[(lnz[i], "-")], u_idx
                # get_best_d (:487-526)
                d = d_idx = None
                for q in p_arr:
                    if ampl[q][0] < j + left <= ampl[q][1]:
                        cd = _at(m[q], jpos(q, i, j) - 1)
                        if d is None or cd > d:
                            d, d_idx = cd, q
                if d is None:
                    d, d_pred = sm[(lnz[i], "-")] * (i + left), min_pred(i)
                else:
                    d, d_pred = d + sm[(lnz[i], seq[j + left])], d_idx
                if d < u:
                    best_val, dr = (l, "L") if u < l else (u, "U")
                else:
                    best_val, dr = (l, "L") if d < l else (d, "D")
                if dr == "D" and seq[j + left] != lnz[i]:
                    dr = "d"
                m[i][j] = best_val
                path[i][j] = cell(d_pred if dr in "Dd" else (u_pred if dr == "U" else l_pred), dr)
            if m[i][j] >= m[i][best_val_pos]:
                best_val_pos = j
        bsp[i] = best_val_pos + left
    last_row = n - 2
    if not m[last_row]:
        raise RuntimeError("attempt to subtract with overflow")
    last_col = len(m[last_row]) - 1
    for q in pred[n - 1]:
        if ampl[q][1] - ampl[q][0] < 1:
            raise RuntimeError("attempt to subtract with overflow")
        tmp = (ampl[q][1] - ampl[q][0]) - 1
        if m[q][tmp] > m[last_row][last_col]:
            last_row, last_col = q, tmp
    out = ""
    # band_ampl_enough (:428-476)
    i, j, ok = last_row, last_col, True
    while _at(path[i], j)[1] != "O":
        left, right = ampl[i]
        if i == 0 or (j == 0 and left == 0):
            break
        if (j == 0 and left != 0) or (j == right - left - 1 and right != L):
            ok = False
            break
        prd, dr = path[i][j]
        jp = jpos(prd, i, j)
        if jp < 0:
            raise RuntimeError("attempt to subtract with overflow")
        if dr in "Dd":
            i, j = prd, jp - 1
        elif dr == "L":
            j -= 1
        elif dr == "U":
            i, j = prd, jp
        else:
            raise RuntimeError("explicit panic")
        if j < 0:
            raise RuntimeError("attempt to subtract with overflow")
    if not ok:
        out += "Band length probably too short, maybe try with larger b and f\n"
    return out + gaf_of_global_abpoa(path, seq, name, ampl, last_row, last_col, hofp), m[last_row][last_col]


def gaf_of_global_abpoa(path, seq, name, ampl, last_row, last_col, hofp):
    """gaf_output.rs:254-382"""
    col, row = last_col, last_row
    hia, cigars = [], []
    cigar = ""
    cm = ci = cd = 0
    curr_handle, last_dir = "", " "
    path_length = residues = 0
    while _at(path[row], col)[1] != "O":
        prd, dr = path[row][col]
        if hofp[row] != curr_handle:
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cigars.insert(0, cigar)
            cigar = ""
            cm = ci = cd = 0
        curr_handle = hofp[row]
        if dr.upper() != last_dir.upper():
            cigar = set_cigar_substring(cm, ci, cd, cigar)
            cm = ci = cd = 0
        last_dir = dr
        p_left = ampl[prd][0]
        if ampl[row][0] < p_left:
            j_pos = col - (p_left - ampl[row][0])
            if j_pos < 0:
                raise RuntimeError("attempt to subtract with overflow")
        else:
            j_pos = col + (ampl[row][0] - p_left)
        if dr in "Dd":
            hia.append(hofp[row])
            row, col = prd, j_pos - 1
            cm += 1
            path_length += 1
            if dr == "D":
                residues += 1
        elif dr == "L":
            col -= 1
            cd += 1
        elif dr == "U":
            hia.append(hofp[row])
            row, col = prd, j_pos
            ci += 1
            path_length += 1
        else:
            raise RuntimeError("impossible value in poa path")
        if col < 0:
            raise RuntimeError("attempt to subtract with overflow")
    cigar = set_cigar_substring(cm, ci, cd, cigar)
    cigars.insert(0, cigar)
    hia = dedup(hia)
    hia.reverse()
    return gaf_string(name, len(seq) - 1, col, last_col + ampl[last_row][0], _STRAND[0], [int(h) for h in hia], path_length,
                      node_start(hofp, row), node_start(hofp, last_row), residues, "*", "*", ",".join(cigars[:len(cigars) - 1])) + "\n"


def rev_and_compl(seq):
    """sequences.rs:64-82"""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    return ["$"] + [comp[c] for c in reversed(seq[1:])]


def reverse_handle_map(nwp, segs):
    """utils::create_handle_pos_in_lnz(.., amb_mode = true) (utils.rs:144-165, graph.rs:128-143): the sorted handles
    reversed (and flipped: the id stays), consumed in lnz order"""
    ids = sorted(segs, reverse=True)
    hofp = {0: "-1"}
    cur = 0
    for i in range(1, len(nwp) - 1):
        if nwp[i]:
            cur += 1
        hofp[i] = str(ids[cur - 1])
    return hofp


def run_poa_amb(mode, fasta_text, gfa_text, match=2, mismatch=4, gap_open=4, gap_ext=2, extra_b=1, extra_f=0.01, matrix=None):
    """main.rs:47-252 with -s true on an AVX2 machine (modes 0-3): the forward call, the reverse-complement retry where the
    mode asks for it, the selection rule of each mode; returns stdout"""
    seqs, names = read_fasta(fasta_text)
    segs, _paths = read_gfa(gfa_text)
    lnz, nwp, pred, hofp = create_graph_struct(segs, read_gfa_links(gfa_text))
    hofp_rev = reverse_handle_map(nwp, segs)
    sm = _scores(match, mismatch, matrix)
    smf = f32_scores(sm)
    r_values = set_r_values(nwp, pred, len(lnz))
    out = ""

    def split(text):      # warnings println!'d during the call, and the record main writes afterwards
        lines = text.splitlines(keepends=True)
        return "".join(lines[:-1]), lines[-1]

    def with_strand(strand, fn):
        _STRAND[0] = strand
        try:
            return fn()
        finally:
            _STRAND[0] = "+"

    for k, seq in enumerate(seqs):
        v = F32(F32(extra_b) + F32(F32(extra_f) * F32(len(seq))))
        bta = 0 if not (v > 0) else int(v)
        rseq = None
        if mode == 0:
            ftext, fs = mode0_exec_simd(seq, names[k], lnz, nwp, pred, smf, bta, hofp, r_values)
            retry = fs < 0
            if retry:
                rtext, rs = with_strand("-", lambda: mode0_exec_scalar(rev_and_compl(seq), names[k], lnz, nwp, pred, sm, bta, hofp_rev))
                take_rev = rs > fs
        elif mode == 1:
            ftext, fs = mode1_exec_simd(seq, names[k], lnz, nwp, pred, smf, hofp)
            retry = True
            rtext, rs = with_strand("-", lambda: mode1_exec_simd(rev_and_compl(seq), names[k], lnz, nwp, pred, smf, hofp_rev))
            take_rev = not (fs < rs)      # main.rs:160-164 writes the FORWARD record when it is the lower one
        elif mode == 2:
            ftext, fs = mode2_exec(seq, names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext, bta, hofp)
            retry = fs < 0
            if retry:
                rtext, rs = with_strand("-", lambda: mode2_exec(rev_and_compl(seq), names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext,
                                                                bta, hofp_rev))
                take_rev = rs > fs
        else:
            ftext, fs = mode3_exec(seq, names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext, hofp)
            retry = True
            rtext, rs = mode3_exec(rev_and_compl(seq), names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext, hofp_rev)   # amb_mode = false, main.rs:242
            take_rev = rs > fs
        fw, fg = split(ftext)
        out += fw
        if retry:
            rw, rgaf = split(rtext)
            out += rw + (rgaf if take_rev else fg)
        else:
            out += fg
    return out


def run_poa(mode, fasta_text, gfa_text, match=2, mismatch=4, gap_open=4, gap_ext=2, extra_b=1, extra_f=0.01, max_reads=None, matrix=None):
    """main.rs:47-101 (mode 0), :103-169 (mode 1), :215-252 (mode 3) on an AVX2 machine, without -s; returns stdout"""
    seqs, names = read_fasta(fasta_text)
    segs, _paths = read_gfa(gfa_text)
    lnz, nwp, pred, hofp = create_graph_struct(segs, read_gfa_links(gfa_text))
    sm = _scores(match, mismatch, matrix)
    smf = f32_scores(sm)
    r_values = set_r_values(nwp, pred, len(lnz)) if mode == 0 else None
    out = ""
    for k, seq in enumerate(seqs):
        if max_reads is not None and k >= max_reads:
            break
        if mode == 0:
            v = F32(F32(extra_b) + F32(F32(extra_f) * F32(len(seq))))
            bta = 0 if not (v > 0) else int(v)
            text, _ = mode0_exec_simd(seq, names[k], lnz, nwp, pred, smf, bta, hofp, r_values)
        elif mode == 1:
            text, _ = mode1_exec_simd(seq, names[k], lnz, nwp, pred, smf, hofp)
        else:
            text, _ = mode3_exec(seq, names[k], lnz, nwp, pred, sm, -gap_open, -gap_ext, hofp)
        out += text
    return out


# ------------------------------------------------------------------------------------------------ modes 6 / 7
def gap_pathwise_exec(seq, g, sm, o, e, semi):
    """pathwise_alignment_gap.rs:4-574 (semi = False) / pathwise_alignment_gap_semi.rs:5-473 (semi = True): the two files
    share the interior cell code word for word (diff: the `(_, 0)` arm and the ending). Returns (cigar line, best path).
    The three tensors hold the reference's delta encoding: the alpha path's entry is absolute, every other entry is
    relative to it."""
    lnz, nwp, alphas, P, pn = g.lnz, g.nwp, g.alphas, g.P, g.pn
    n, L = len(lnz), len(seq)
    dpm = [[[0] * P for _ in range(L)] for _ in range(n)]
    x = [[[0] * P for _ in range(L)] for _ in range(n)]
    y = [[[0] * P for _ in range(L)] for _ in range(n)]

    def fix_multiple_alphas(i, j, alphas_deltas, tensors):
        # "remove multiple alpha": every group but the one of alphas[i] is re-expressed relative to alphas[i]
        ai = alphas[i]
        for a, delta in alphas_deltas.items():
            if a != ai:
                for t in tensors:
                    t[i][j][a] -= t[i][j][ai]
                for path in delta:
                    if path != a:
                        for t in tensors:
                            t[i][j][path] += t[i][j][a]

    for i in range(n - 1):
        ai = alphas[i]
        for j in range(L):
            if i == 0 and j == 0:
                continue
            if semi and j == 0:
                continue                      # `(_, 0) => dpm[0][0] = vec![0; path_number]`: no effect
            if i == 0:
                a0 = alphas[0]
                y[i][j][a0] = o + e * j
                dpm[i][j][a0] = y[i][j][a0]
                for k in range(a0 + 1, P):
                    y[i][j][k] = y[i][j - 1][k]
                    dpm[i][j][k] = y[i][j][k]
                continue
            if j == 0:                        # first column of the global mode (pathwise_alignment_gap.rs:35-149)
                if not nwp[i]:
                    common = sorted(pn[i] & pn[i - 1])
                    if alphas[i - 1] in common:
                        for path in common:
                            if path == ai:
                                x[i][j][path] = o + e if i == 1 else x[i - 1][j][path] + e
                            else:
                                x[i][j][path] = x[i - 1][j][path]
                            dpm[i][j][path] = x[i][j][path]
                    else:
                        x[i][j][ai] = x[i - 1][j][ai] + x[i - 1][j][alphas[i - 1]] + e if i != 1 else o + e
                        dpm[i][j][ai] = x[i][j][ai]
                        for path in common:
                            if path != ai:
                                x[i][j][path] = x[i - 1][j][path] - x[i - 1][j][ai]
                                dpm[i][j][path] = x[i][j][path]
                else:
                    alphas_deltas = {}
                    for q, q_paths in preds_and_paths(g, i):
                        common = sorted(pn[i] & q_paths)
                        aq = alphas[q]
                        if aq in common:
                            alphas_deltas[aq] = common
                            x[i][j][aq] = o + e if q == 0 else x[q][j][aq] + e
                            dpm[i][j][aq] = x[i][j][aq]
                            for path in common:
                                if path != aq:
                                    x[i][j][path] = x[q][j][path]
                                    dpm[i][j][path] = x[i][j][path]
                        else:
                            if not common:
                                raise RuntimeError("position(|is_in| is_in).unwrap() on None")
                            ta = ai if ai in common else common[0]
                            alphas_deltas[ta] = common
                            x[i][j][ta] = o + e if q == 0 else x[q][j][ta] + x[q][j][aq] + e
                            dpm[i][j][ta] = x[i][j][ta]
                            for path in common:
                                if path != ta:
                                    x[i][j][path] = x[q][j][path] - x[q][j][ta]
                                    dpm[i][j][path] = x[i][j][path]
                    # here only x is adjusted and dpm re-copied from it (:134-147)
                    for a, delta in alphas_deltas.items():
                        if a != ai:
                            x[i][j][a] -= x[i][j][ai]
                            dpm[i][j][a] = x[i][j][a]
                            for path in delta:
                                if path != a:
                                    x[i][j][path] += x[i][j][a]
                                    dpm[i][j][path] = x[i][j][path]
                continue
            s = sm[(lnz[i], seq[j])]
            if not nwp[i]:
                ap = alphas[i - 1]
                common = sorted(pn[i] & pn[i - 1])
                same = ap in common
                # set y
                if same:
                    u_y = y[i - 1][j][ap] + e
                    u_dpm = dpm[i - 1][j][ap] + o + e
                else:
                    u_y = y[i - 1][j][ap] + y[i - 1][j][ai] + e
                    u_dpm = dpm[i - 1][j][ap] + dpm[i - 1][j][ai] + o + e
                src = dpm if u_dpm >= u_y else y
                for path in common:
                    if path != ai:
                        y[i][j][path] = src[i - 1][j][path] if same else src[i - 1][j][path] - src[i - 1][j][ai]
                y[i][j][ai] = u_dpm if u_dpm >= u_y else u_y
                u = y[i][j][ai]
                # set x
                l_x = x[i][j - 1][ai] + e
                l_dpm = dpm[i][j - 1][ai] + o + e
                src = dpm if l_dpm >= l_x else x
                for path in common:
                    if path != ai:
                        x[i][j][path] = src[i][j - 1][path]
                x[i][j][ai] = l_dpm if l_dpm >= l_x else l_x
                l = x[i][j][ai]
                # set dpm
                d = dpm[i - 1][j - 1][ap] + s if same else dpm[i - 1][j - 1][ap] + dpm[i - 1][j - 1][ai] + s
                best = max(d, u, l)
                dpm[i][j][ai] = best
                for path in common:
                    if path != ai:
                        if best == d:
                            dpm[i][j][path] = dpm[i - 1][j - 1][path] if same else dpm[i - 1][j - 1][path] - dpm[i - 1][j - 1][ai]
                        elif best == u:
                            dpm[i][j][path] = y[i][j][path]
                        else:
                            dpm[i][j][path] = x[i][j][path]
                continue
            # several predecessors: every incoming edge keeps its own alpha first
            alphas_deltas = {}
            for q, q_paths in preds_and_paths(g, i):
                common = sorted(pn[i] & q_paths)
                aq = alphas[q]
                if aq in common:
                    alphas_deltas[aq] = common
                    u_y = y[q][j][aq] + e
                    u_dpm = dpm[q][j][aq] + o + e
                    if u_dpm >= u_y:
                        for path in common:
                            if path != aq:
                                y[i][j][path] = dpm[q][j][path]
                        y[i][j][aq] = u_dpm
                    else:
                        for path in common:
                            if path != ai:            # alphas[i], not alphas[p] (pathwise_alignment_gap.rs:338)
                                y[i][j][path] = y[q][j][path]
                        y[i][j][aq] = u_y
                    u = y[i][j][aq]
                    if aq == ai:
                        l_x = x[i][j - 1][aq] + e
                        l_dpm = dpm[i][j - 1][aq] + o + e
                    else:
                        l_x = x[i][j - 1][aq] + x[i][j - 1][ai] + e
                        l_dpm = dpm[i][j - 1][ai] + dpm[i][j - 1][aq] + o + e
                    src = dpm if l_dpm >= l_x else x
                    for path in common:
                        if path != aq:
                            x[i][j][path] = src[i][j - 1][path] if aq == ai else src[i][j - 1][path] - src[i][j - 1][aq]
                    x[i][j][aq] = l_dpm if l_dpm >= l_x else l_x
                    l = x[i][j][aq]
                    d = dpm[q][j - 1][aq] + s
                    best = max(d, u, l)
                    dpm[i][j][aq] = best
                    for path in common:
                        if path != aq:
                            if best == d:
                                dpm[i][j][path] = dpm[q][j - 1][path]
                            elif best == u:
                                dpm[i][j][path] = y[i][j][path]
                            else:
                                dpm[i][j][path] = x[i][j][path]
                else:
                    if not common:
                        raise RuntimeError("position(|is_in| is_in).unwrap() on None")
                    ta = ai if ai in common else common[0]
                    alphas_deltas[ta] = common
                    u_y = y[q][j][aq] + y[q][j][ta] + e
                    u_dpm = dpm[q][j][aq] + dpm[q][j][ta] + o + e
                    src = dpm if u_dpm >= u_y else y
                    for path in common:
                        if path != ta:
                            y[i][j][path] = src[q][j][path] - src[q][j][ta]
                    y[i][j][ta] = u_dpm if u_dpm >= u_y else u_y
                    u = y[i][j][ta]
                    if ai == ta:
                        l_x = x[i][j - 1][ai] + e
                        l_dpm = dpm[i][j - 1][ai] + o + e
                    else:
                        l_x = x[i][j - 1][ai] + x[i][j - 1][ta] + e
                        l_dpm = dpm[i][j - 1][ai] + dpm[i][j - 1][ta] + o + e
                    src = dpm if l_dpm >= l_x else x
                    for path in common:
                        if path != ta:
                            x[i][j][path] = src[i][j - 1][path] if ta == ai else src[i][j - 1][path] - src[i][j - 1][ta]
                    x[i][j][ta] = l_dpm if l_dpm >= l_x else l_x
                    l = x[i][j][ta]
                    d = dpm[q][j - 1][aq] + dpm[q][j - 1][ta] + s
                    best = max(d, u, l)
                    dpm[i][j][ta] = best
                    for path in common:
                        if path != ta:
                            if best == d:
                                dpm[i][j][path] = dpm[q][j - 1][path] - dpm[q][j - 1][ta]
                            elif best == u:
                                dpm[i][j][path] = y[i][j][path]
                            else:
                                dpm[i][j][path] = x[i][j][path]
            fix_multiple_alphas(i, j, alphas_deltas, (dpm, x, y))

    if not semi:
        results = [0] * P
        for q, paths in preds_and_paths(g, n - 1):
            for path in sorted(paths):
                if path == alphas[q]:
                    results[path] = dpm[q][L - 1][path]
                else:
                    results[path] = dpm[q][L - 1][path] + dpm[q][L - 1][alphas[q]]
        best_path = max((sc, path) for path, sc in enumerate(results))[1]
        end = 0
        for q, paths in preds_and_paths(g, n - 1):
            if best_path in paths:
                end = q
    else:
        # best_ending_node (pathwise_alignment_gap_semi.rs:448-473): every one of the P slots competes, the ones that
        # are not on the node with whatever they hold
        mx = None
        end = best_path = 0
        for i in range(n - 1):
            ab = list(dpm[i][L - 1])
            for path in sorted(pn[i]):
                if path != alphas[i]:
                    ab[path] = ab[path] + ab[alphas[i]]
            sc, path = max((sc, path) for path, sc in enumerate(ab))
            if mx is None or sc > mx:
                mx, end, best_path = sc, i, path
    return _build_alignment_gap(dpm, x, y, g, best_path, end, semi), best_path


def _build_alignment_gap(dpm, x, y, g, bp, ending_node, semi):
    """pathwise_alignment_output.rs:186-316 (build_alignment_gap) / :318-451 (build_alignment_semiglobal_gap): the walks
    are identical, the tails differ. The gap-chain tests compare RAW tensor entries (`dpm[i][j][bp] < y[i][j][bp]`)."""
    nwp, alphas = g.nwp, g.alphas

    def absolute(t, i, j):
        return t[i][j][bp] if alphas[i] == bp else t[i][j][bp] + t[i][j][alphas[i]]

    def pred_on_path(i):
        found = None
        for q, paths in preds_and_paths(g, i):
            if bp in paths:
                found = q
        return found

    cigar = []
    i = ending_node
    j = len(dpm[i]) - 1
    while i != 0 and j != 0:
        curr = absolute(dpm, i, j)
        predecessor = None
        if not nwp[i]:
            d, u, l = absolute(dpm, i - 1, j - 1), absolute(dpm, i - 1, j), absolute(dpm, i, j - 1)
        else:
            d = u = l = 0
            for q, paths in preds_and_paths(g, i):
                if bp in paths:
                    predecessor = q
                    d, u = absolute(dpm, q, j - 1), absolute(dpm, q, j)
                    l = absolute(dpm, i, j - 1)
        mx = max(d, u, l)
        if mx == d:
            cigar.append("d" if curr < d else "D")
            i = i - 1 if predecessor is None else predecessor
            j -= 1
        elif mx == u:
            cigar.append("U")
            i = i - 1 if predecessor is None else predecessor
            while dpm[i][j][bp] < y[i][j][bp]:
                cigar.append("U")
                if nwp[i]:
                    q = pred_on_path(i)
                    if q is None and predecessor is None:
                        raise RuntimeError("called `Option::unwrap()` on a `None` value")
                    if q is None:
                        raise RuntimeError("the reference loops forever here (no predecessor on the best path)")
                    predecessor = q
                else:
                    predecessor = i - 1
                i = predecessor
        else:
            cigar.append("L")
            j -= 1
            while dpm[i][j][bp] < x[i][j][bp]:
                cigar.append("L")
                j -= 1
                if j < 0:
                    raise RuntimeError("attempt to subtract with overflow")
    while j > 0:
        cigar.append("L")
        j -= 1
    if not semi:
        while i > 0:
            cigar.append("U")
            i -= 1
        cigar.reverse()
        if cigar:
            cigar.pop()
        return build_cigar(cigar)
    cigar.reverse()

    def count_back(i):
        steps = 0
        while i > 0:
            if nwp[i]:
                q = pred_on_path(i)
                if q is None:
                    raise RuntimeError("the reference loops forever here (no predecessor on the best path)")
                i = q
            else:
                i -= 1
            steps += 1
        return steps
    starting_node = count_back(i)
    final_node = count_back(ending_node)
    return f"{build_cigar(cigar)}\t({starting_node} {final_node})"


def run_gap_pathwise(mode, fasta_text, gfa_text, match=2, mismatch=4, gap_open=4, gap_ext=2, matrix=None):
    """main.rs:271-288 (modes 6 / 7); returns stdout"""
    seqs, names = read_fasta(fasta_text)
    segs, paths = read_gfa(gfa_text)
    sm = _scores(match, mismatch, matrix)
    g = create_path_graph(segs, paths)
    out = ""
    for k, seq in enumerate(seqs):
        cigar, best_path = gap_pathwise_exec(seq, g, sm, -gap_open, -gap_ext, mode == 7)
        out += f"{cigar}\nBest path sequence {k}: {best_path}\n"
    return out


# ------------------------------------------------------------------------------------------------ driver
def run(mode, fasta_text, gfa_text, match=2, mismatch=4, base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0, max_reads=None, matrix=None):
    """main.rs:253-312 for modes 4, 5, 8, 9 with match / mismatch scoring; returns stdout."""
    seqs, names = read_fasta(fasta_text)
    segs, paths = read_gfa(gfa_text)
    sm = _scores(match, mismatch, matrix)
    g = create_path_graph(segs, paths)
    out = []
    if mode in (8, 9):
        rg = create_reverse_path_graph(g)
        displ = Displacement(g, rg)
    for k, seq in enumerate(seqs):
        if max_reads is not None and k >= max_reads:
            break
        if mode == 4:
            out.append(mode4(seq, g, sm, names[k]))
        elif mode == 5:
            out.append(mode5(seq, g, sm, names[k]))
        else:
            out.append(mode89(mode, seq, g, rg, sm, base_rec_cost, multi_rec_cost, displ, rec_band_width, names[k]))
    return "".join(x + "\n" for x in out)
