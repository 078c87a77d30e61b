"""Synthetic pangenome graphs and reads for parity tests and bench.py (generator is ours).

Backbone of uniform random ACGT cut into segments of 1-64 bp (mean ~16), an SNP bubble roughly every 100 bp
(two single-base alleles) and an indel bubble roughly every 500 bp (a 1-10 bp segment plus a bypass edge),
one source and one sink segment, segment ids increasing in topological order, every segment on >= 1 path,
P haplotype paths choosing alleles i.i.d. Reads are substrings of a path (or of a mosaic of paths) with
substitution / insertion / deletion errors 1:1:1. Everything is seeded.
"""
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


class SynthGraph:
    def __init__(self, segs, links, paths, bubbles):
        self.segs = segs        # list[bytes], id = index + 1
        self.links = links      # list[(from_id, to_id)]
        self.paths = paths      # list[list[seg_id]]
        self.bubbles = bubbles  # list of ("snp", bb, a0, a1) / ("indel", bb, ins) (segment ids)

    @property
    def n_chars(self):
        return sum(len(s) for s in self.segs)

    def gfa(self):
        out = ["H\tVN:Z:1.0"]
        succ = {}
        for a, b in self.links:
            succ.setdefault(a, []).append(b)
        for i, s in enumerate(self.segs, 1):
            out.append(f"S\t{i}\t{s.decode()}")
            for b in succ.get(i, []):
                out.append(f"L\t{i}\t+\t{b}\t+\t0M")
        for k, p in enumerate(self.paths):
            out.append(f"P\tpath{k}\t" + ",".join(f"{x}+" for x in p) + "\t*")
        return "\n".join(out) + "\n"

    def path_sequence(self, p):
        return b"".join(self.segs[s - 1] for s in p)


def make_graph(total_bp, n_paths, seed=1, mean_seg=16, p_snp=0.13, p_indel=0.03):
    rng = np.random.default_rng(seed)
    segs, links, choices = [], [], []  # choices: per bubble, list of alternative id-lists
    structure = []  # sequence of ("bb", id) / ("bubble", k)
    bubbles = []

    def new_seg(seq):
        segs.append(seq)
        return len(segs)

    def rand_seq(k):
        return BASES[rng.integers(0, 4, size=k)].tobytes()

    made = 0
    prev_tails = []  # segment ids whose out-edges go to the next backbone segment
    while True:
        ln = int(min(64, max(1, rng.geometric(1.0 / mean_seg))))
        bb = new_seg(rand_seq(ln))
        made += ln
        for t in prev_tails:
            links.append((t, bb))
        structure.append(("bb", bb))
        if made >= total_bp:
            break
        u = rng.random()
        if u < p_snp:
            b0 = int(rng.integers(0, 4))
            b1 = (b0 + int(rng.integers(1, 4))) % 4
            a0 = new_seg(bytes([BASES[b0]]))
            a1 = new_seg(bytes([BASES[b1]]))
            links.append((bb, a0))
            links.append((bb, a1))
            prev_tails = [a0, a1]
            structure.append(("bubble", len(choices)))
            choices.append([[a0], [a1]])
            bubbles.append(("snp", bb, a0, a1))
            made += 1
        elif u < p_snp + p_indel:
            ins = new_seg(rand_seq(int(rng.integers(1, 11))))
            links.append((bb, ins))
            prev_tails = [ins, bb]  # bypass edge bb -> next backbone
            structure.append(("bubble", len(choices)))
            choices.append([[ins], []])
            bubbles.append(("indel", bb, ins))
        else:
            prev_tails = [bb]
    # links sorted so that L lines of one segment keep creation order (predecessor list order = L-line order)
    nb = len(choices)
    picks = rng.integers(0, 2, size=(max(n_paths, 1), max(nb, 1)))
    if n_paths >= 1:
        picks[0, :] = 0
    if n_paths >= 2:
        picks[1, :] = 1
    paths = []
    for p in range(n_paths):
        paths.append(path_from_picks(structure, choices, picks[p]))
    g = SynthGraph(segs, links, paths, bubbles)
    g.structure, g.choices, g.picks = structure, choices, picks
    return g


def path_from_picks(structure, choices, pick):
    out = []
    for kind, v in structure:
        if kind == "bb":
            out.append(v)
        else:
            out.extend(choices[v][int(pick[v])])
    return out


def mutate(rng, seq, err):
    """Substitution / insertion / deletion at total rate err, 1:1:1 (vectorised)."""
    if err <= 0 or not seq:
        return seq
    a = np.frombuffer(seq, dtype=np.uint8)
    n = len(a)
    mut = rng.random(n) < err
    kinds = rng.integers(0, 3, size=n)
    cnt = np.ones(n, dtype=np.int64)
    is_sub, is_ins, is_del = mut & (kinds == 0), mut & (kinds == 1), mut & (kinds == 2)
    cnt[is_ins] = 2
    cnt[is_del] = 0
    lut = np.zeros(256, dtype=np.uint8)
    lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
    sub = BASES[(lut[a] + rng.integers(1, 4, size=n)) % 4]
    ins = BASES[rng.integers(0, 4, size=n)]
    first = np.where(is_sub, sub, np.where(is_ins, ins, a))
    out = np.repeat(first, cnt)
    pos = np.cumsum(cnt) - 1
    out[pos[is_ins]] = a[is_ins]  # inserted base first, then the original one
    return out.tobytes()


def make_reads(g, n_reads, read_len, err=0.05, seed=3, mosaic_breaks=0, exact_len=True):
    """Reads sampled from paths (mosaic_breaks > 0: from mosaics of mosaic_breaks+1 paths)."""
    rng = np.random.default_rng(seed)
    P = len(g.paths)
    seqs = {}
    reads = []
    for _ in range(n_reads):
        if mosaic_breaks > 0 and len(g.choices) > mosaic_breaks and P > 1:
            ps = rng.choice(P, size=mosaic_breaks + 1, replace=P < mosaic_breaks + 1)
            cuts = np.sort(rng.choice(len(g.choices), size=mosaic_breaks, replace=False))
            pick = np.empty(len(g.choices), dtype=np.int64)
            lo = 0
            for k, pz in enumerate(ps):
                hi = cuts[k] if k < mosaic_breaks else len(g.choices)
                pick[lo:hi] = g.picks[pz, lo:hi]
                lo = hi
            src = b"".join(g.segs[s - 1] for s in path_from_picks(g.structure, g.choices, pick))
        else:
            p = int(rng.integers(0, P)) if P else 0
            if p not in seqs:
                seqs[p] = g.path_sequence(g.paths[p]) if P else b"".join(g.segs)
            src = seqs[p]
        want = min(read_len, len(src))
        span = min(len(src), int(want * 1.2) + 8)
        st = int(rng.integers(0, len(src) - span + 1))
        r = mutate(rng, src[st:st + span], err)
        if exact_len:
            r = r[:want]
        if not r:
            r = src[:1]
        reads.append(r.decode())
    return reads


def fasta(reads, prefix="read"):
    return "".join(f">{prefix}{i}\n{r}\n" for i, r in enumerate(reads))
