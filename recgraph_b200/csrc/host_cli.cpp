// `recgraph` command line (main.rs:25-329 + args_parser.rs:3-202) on top of the public C ABI only.
#include <sys/stat.h>

#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/recgraph_b200.h"

namespace {

struct Args {  // args_parser.rs:3-147: same flags and defaults
    std::string sequence_path, graph_path, out_file = "standard output";
    int alignment_mode = 0, match_score = 2, mismatch_score = 4, gap_open = 4, gap_extension = 2, base_rec_cost = 4;
    std::string matrix = "none", amb_strand = "false";
    float multi_rec_cost = 0.1f, rec_band_width = 1.0f, extra_f = 0.01f;
    int extra_b = 1;
    int gpus = 1;  // not a reference flag: --gpus N shards the reads over N devices of the box
};

struct OptName {
    char s;
    const char* l;
};
const OptName OPTS[] = {{'o', "out_file"}, {'m', "aln-mode"}, {'M', "match"}, {'X', "mismatch"}, {'t', "matrix"},
                        {'O', "gap-open"}, {'E', "gap-ext"}, {'r', "multi-rec-cost"}, {'R', "base-rec-cost"},
                        {'B', "rec-band-width"}, {'s', "amb-strand"}, {'b', "extra-b"}, {'f', "extra-f"}, {'G', "gpus"}};

bool assign(Args& a, char s, const std::string& v) {
    try {
        size_t pos = 0;
        switch (s) {
            case 'o': a.out_file = v; return true;
            case 't': a.matrix = v; return true;
            case 's':
                if (v != "true" && v != "false") return false;
                a.amb_strand = v;
                return true;
            case 'm': a.alignment_mode = std::stoi(v, &pos); break;
            case 'M': a.match_score = std::stoi(v, &pos); break;
            case 'X': a.mismatch_score = std::stoi(v, &pos); break;
            case 'O': a.gap_open = std::stoi(v, &pos); break;
            case 'E': a.gap_extension = std::stoi(v, &pos); break;
            case 'R': a.base_rec_cost = std::stoi(v, &pos); break;
            case 'b': a.extra_b = std::stoi(v, &pos); break;
            case 'G':
                a.gpus = std::stoi(v, &pos);
                if (a.gpus < 1) return false;
                break;
            case 'r': a.multi_rec_cost = std::stof(v, &pos); break;
            case 'B': a.rec_band_width = std::stof(v, &pos); break;
            case 'f': a.extra_f = std::stof(v, &pos); break;
            default: return false;
        }
        return pos == v.size();
    } catch (...) {
        return false;
    }
}

bool parse_args(int argc, const char** argv, Args& a, std::string& err) {
    std::vector<std::string> positional;
    for (int k = 1; k < argc; k++) {
        std::string s = argv[k];
        char sc = 0;
        std::string val;
        bool have_val = false;
        if (s.size() > 2 && s[0] == '-' && s[1] == '-') {
            std::string name = s.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                val = name.substr(eq + 1);
                name.resize(eq);
                have_val = true;
            }
            for (auto& o : OPTS)
                if (name == o.l) sc = o.s;
        } else if (s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9') && s[1] != '.') {
            for (auto& o : OPTS)
                if (s[1] == o.s) sc = o.s;
            if (sc && s.size() > 2) {
                val = s.substr(s[2] == '=' ? 3 : 2);
                have_val = true;
            }
        } else {
            positional.push_back(s);
            continue;
        }
        if (!sc) {
            err = "error: Found argument '" + s + "' which wasn't expected, or isn't valid in this context\n";
            return false;
        }
        if (!have_val) {
            if (k + 1 >= argc) {
                err = "error: The argument '" + s + "' requires a value but none was supplied\n";
                return false;
            }
            val = argv[++k];
        }
        if (!assign(a, sc, val)) {
            err = "error: Invalid value \"" + val + "\" for '" + s + "'\n";
            return false;
        }
    }
    if (positional.size() != 2) {
        err = "error: The following required arguments were not provided:\n    <SEQUENCE_PATH>\n    <GRAPH_PATH>\n";
        return false;
    }
    a.sequence_path = positional[0];
    a.graph_path = positional[1];
    return true;
}

char* dup_text(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

// utils.rs:200-219. `number` is i+1 for POA modes and i for pathwise modes (main.rs:260,268,311), file created
// when it does not exist or number == 1, appended otherwise — kept as is.
bool write_gaf(const Args& a, const std::string& text, size_t number, std::string& out) {
    if (a.out_file == "standard output") {
        out += text;
        return true;
    }
    struct stat st;
    bool exists = stat(a.out_file.c_str(), &st) == 0;
    FILE* f = fopen(a.out_file.c_str(), (exists && number != 1) ? "ab" : "wb");
    if (!f) return false;
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    return true;
}

// Everything one device does for a contiguous range [lo, hi) of the reads: own context, own copy of the graph, alignment,
// `-s` retries, text. Per read: the lines the reference println!s while aligning (warnings; the whole output for modes
// 6 / 7) and the GAF record that goes through write_gaf.
struct ShardOut {
    std::vector<std::string> pre, record;
    int32_t emitted = 0; // leading reads of the shard whose output is valid: the reference prints read by read, so everything
                         // before the first read it panics on has already been written when it dies
    std::string tail;    // what the reference had already printed for the read it dies on (warning lines)
    int code = 0;        // 0 ok, 101 reference panic, 3 device / library error
    std::string err;
};

void run_shard(const Args& a, const rg_scoring& sc, const rg_reads& reads, int device, int32_t lo, int32_t hi, ShardOut& so) {
    const int mode = a.alignment_mode;
    const int32_t n = hi - lo;
    so.pre.assign(n, "");
    so.record.assign(n, "");
    rg_ctx* ctx = nullptr;
    int rc = rg_init(device, &ctx);
    if (rc != RG_OK) {
        so.code = 3;
        so.err = std::string("recgraph_b200: device ") + std::to_string(device) + ": " + rg_strerror(rc) + "\n";
        return;
    }
    struct Guard {
        rg_ctx* c;
        ~Guard() { rg_destroy(c); }
    } guard{ctx};
    auto panic = [&](const std::string& msg) {
        so.code = 101;
        so.err = "thread 'main' panicked at '" + msg + "'\n";
    };
    rc = rg_load_gfa_file(ctx, a.graph_path.c_str());
    if (rc != RG_OK) return panic(rg_last_error(ctx));
    rg_set_scoring(ctx, &sc);
    if (n == 0) return;
    const bool amb_strand = a.amb_strand == "true";
    struct Batch {
        std::vector<std::string> warn, record;
        std::vector<int32_t> score, status;
        std::vector<uint32_t> best_path;
        int32_t panic_at = INT32_MAX;   // shard index of the first read the reference panics on
    };
    // idx: position inside this shard of every read of the batch (empty = identity)
    auto run_batch = [&](int dev_mode, int32_t nb, const uint8_t* codes, const uint64_t* off, const std::vector<int32_t>& idx,
                         int amb_flags, Batch& bt) -> bool {
        rg_batch_result res;
        int r = rg_align_batch(ctx, dev_mode, nb, codes, off, &res);
        if (r == RG_ERR_REF_PANIC) {
            panic(rg_last_error(ctx));
            return false;
        }
        if (r != RG_OK) {
            so.err += std::string("recgraph_b200: ") + rg_strerror(r) + ": " + rg_last_error(ctx) + "\n";
            so.code = 3;
            return false;
        }
        bt.warn.resize(nb);
        bt.record.resize(nb);
        bt.score.resize(nb);
        bt.best_path.resize(nb);
        bt.status.resize(nb);
        for (int32_t k = 0; k < nb; k++) {
            bt.status[k] = res.reads[k].status;
            bt.best_path[k] = res.reads[k].best_path;
            bt.score[k] = res.reads[k].score;
            const int32_t i = lo + (idx.empty() ? k : idx[k]);   // index in the input file
            if (res.reads[k].status & RG_READ_REF_PANIC) {
                bt.panic_at = std::min(bt.panic_at, idx.empty() ? k : idx[k]);
                continue;
            }
            if (res.reads[k].status & RG_READ_TRACE_OVERFLOW) {
                so.err += "recgraph_b200: trace buffers overflowed for read " + std::to_string(i + 1) + "\n";
                so.code = 3;
                return false;
            }
        }
        // text of the records: independent per read, formatted by a few host threads (rg_format_gaf only reads the context)
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::max(1u, std::thread::hardware_concurrency()), (int64_t)16,
                                                                  (int64_t)(nb / 64 + 1)}));
        std::vector<int> bad(T, 0);
        auto work = [&](int t) {
            std::vector<char> buf(1 << 16);
            for (int32_t k = (int32_t)((int64_t)nb * t / T); k < (int32_t)((int64_t)nb * (t + 1) / T); k++) {
                const int32_t i = lo + (idx.empty() ? k : idx[k]);
                if ((idx.empty() ? k : idx[k]) >= bt.panic_at) continue;   // never printed by the reference
                uint32_t len = (uint32_t)(off[k + 1] - off[k]);
                int64_t need = rg_format_gaf(ctx, dev_mode, &res, k, reads.names[i], len, amb_flags, buf.data(), buf.size());
                if (need < 0) {
                    bad[t] = 1;
                    return;
                }
                if ((size_t)need >= buf.size()) {
                    buf.resize((size_t)need + 1);
                    rg_format_gaf(ctx, dev_mode, &res, k, reads.names[i], len, amb_flags, buf.data(), buf.size());
                }
                std::string text(buf.data(), (size_t)need);
                size_t cut = text.size() > 1 ? text.rfind('\n', text.size() - 2) : std::string::npos;
                if (cut != std::string::npos) {
                    bt.warn[k] = text.substr(0, cut + 1);
                    text.erase(0, cut + 1);
                }
                bt.record[k] = std::move(text);
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back(work, t);
            for (auto& x : th) x.join();
        }
        for (int t = 0; t < T; t++)
            if (bad[t]) {
                so.code = 3;
                return false;
            }
        return true;
    };
    Batch fwd;
    if (!run_batch(mode, n, reads.codes, reads.off + lo, {}, 0, fwd)) return;
    int32_t limit = std::min(n, fwd.panic_at);   // reads before the first panic are emitted, then the process dies like the reference
    // ---- -s true: reverse-complement retries (main.rs:82-101 mode 0, 150-164 mode 1, 198-214 mode 2, 233-249 mode 3)
    Batch rev;
    std::vector<int32_t> rev_of(n, -1);
    if (amb_strand && mode <= 3) {
        std::vector<int32_t> idx;
        for (int32_t i = 0; i < limit; i++)
            if (mode == 1 || mode == 3 || fwd.score[i] < 0) {  // modes 0 / 2 retry only when the forward score is negative
                rev_of[i] = (int32_t)idx.size();
                idx.push_back(i);
            }
        if (!idx.empty()) {
            std::vector<uint8_t> rcs;
            std::vector<uint64_t> roff{0};
            for (int32_t i : idx) {
                for (uint64_t k = reads.off[lo + i + 1]; k-- > reads.off[lo + i];) {  // sequences.rs:64-82
                    uint8_t c = reads.codes[k];
                    rcs.push_back(c < 4 ? (uint8_t)(3 - c) : c);
                }
                roff.push_back(rcs.size());
            }
            // mode 0 retries with the scalar routine (global_abpoa::exec); mode 3 keeps amb_mode = false (main.rs:242)
            const int rmode = mode == 0 ? RG_MODE_GLOBAL_SCALAR : mode;
            const int flags = mode == 3 ? RG_AMB_HANDLES : (RG_AMB_HANDLES | RG_AMB_STRAND);
            if (!run_batch(rmode, (int32_t)idx.size(), rcs.data(), roff.data(), idx, flags, rev)) return;
            limit = std::min(limit, rev.panic_at);
        }
    }
    so.emitted = limit;
    if (limit < n) {
        // mode 2 prints its band warning (gap_global_abpoa.rs:226) before the traceback that panics (gaf_of_gap_abpoa); with
        // -s the forward alignment of the read has been through all of that when the retry dies
        const char* bw = "Band length probably too short, maybe try with larger b and f\n";
        if (fwd.panic_at == limit) {
            if (fwd.status[limit] & RG_READ_BAND_WARNING) so.tail = bw;
        } else {
            so.tail = fwd.warn[limit];
            if (rev.status[rev_of[limit]] & RG_READ_BAND_WARNING) so.tail += bw;
        }
    }
    if (limit < n) panic("reference panic while aligning read " + std::to_string(lo + limit + 1) + " (see DESIGN.md, reference quirks)");
    for (int32_t i = 0; i < limit; i++) {
        // warning lines are println!'d to stdout by the reference even with -o; only the record goes to the file
        so.pre[i] = fwd.warn[i];
        if (mode == 6 || mode == 7) {
            // main.rs:271-288: exec println!s the CIGAR line, main the best path; no GAF record, -o is not used
            so.pre[i] += fwd.record[i];
            so.pre[i] += "Best path sequence " + std::to_string(lo + i) + ": " + std::to_string(fwd.best_path[i]) + "\n";
            continue;
        }
        const std::string* rec = &fwd.record[i];
        if (rev_of[i] >= 0) {
            const int32_t k = rev_of[i];
            so.pre[i] += rev.warn[k];
            const bool take_rev = mode == 1 ? !(fwd.score[i] < rev.score[k])  // main.rs:160-164 keeps the LOWER score
                                            : rev.score[k] > fwd.score[i];
            if (take_rev) rec = &rev.record[k];
        }
        so.record[i] = *rec;
    }
}

}  // namespace

extern "C" int rg_cli_main(int argc, const char** argv, char** out_text, char** err_text) {
    auto t0 = std::chrono::steady_clock::now();
    std::string out, err;
    int rc = 0;
    Args a;
    rg_reads reads;
    memset(&reads, 0, sizeof reads);
    auto finish = [&](int code) {
        rg_free_reads(&reads);
        if (out_text) *out_text = dup_text(out);
        if (err_text) *err_text = dup_text(err);
        return code;
    };
    auto panic = [&](const std::string& msg) {
        err += "thread 'main' panicked at '" + msg + "'\n";
        return finish(101);
    };
    if (!parse_args(argc, argv, a, err)) return finish(2);

    char ebuf[512] = {0};
    rc = rg_read_fasta_file(a.sequence_path.c_str(), &reads, ebuf, sizeof ebuf);
    if (rc != RG_OK) return panic(ebuf);

    rg_scoring sc;
    rg_default_scoring(&sc);
    if (a.matrix == "HOXD70.mtx" || a.matrix == "HOXD70")
        rg_make_score_matrix(3, 0, 0, &sc);
    else if (a.matrix == "HOXD55.mtx" || a.matrix == "HOXD55")
        rg_make_score_matrix(2, 0, 0, &sc);
    else if (a.matrix == "none")
        rg_make_score_matrix(0, a.match_score, -a.mismatch_score, &sc);  // args_parser.rs:153-156
    else
        return panic("wrong matrix type");
    sc.gap_open = -a.gap_open;  // args_parser.rs:163-166
    sc.gap_ext = -a.gap_extension;
    sc.base_rec_cost = a.base_rec_cost;
    sc.multi_rec_cost = a.multi_rec_cost;
    sc.rec_band_width = a.rec_band_width;
    sc.extra_b = (float)a.extra_b;
    sc.extra_f = a.extra_f;
    sc.fixed_bta = -1;

    const int mode = a.alignment_mode;
    if (mode < 0 || mode > 9) return panic("Alignment mode must be in [0..9]");

    // ---- reads are independent: contiguous shards balanced by read length (every mode's cost grows with it), one host
    // thread + context + graph replica per device, no exchange between devices; records are emitted in input order
    const int G = a.gpus;
    std::vector<int32_t> bound(G + 1, 0);
    {
        const uint64_t total = reads.off[reads.n_reads] - reads.off[0];
        int32_t i = 0;
        for (int d = 1; d < G; d++) {
            const uint64_t target = reads.off[0] + total * (uint64_t)d / (uint64_t)G;
            while (i < reads.n_reads && reads.off[i + 1] <= target) i++;
            bound[d] = i;
        }
        bound[G] = reads.n_reads;
    }
    std::vector<ShardOut> shards(G);
    if (G == 1) {
        run_shard(a, sc, reads, 0, 0, reads.n_reads, shards[0]);
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++) th.emplace_back([&, d] { run_shard(a, sc, reads, d, bound[d], bound[d + 1], shards[d]); });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < G; d++) {
        for (int32_t k = 0; k < shards[d].emitted; k++) {
            const int32_t i = bound[d] + k;
            const size_t number = mode <= 3 ? (size_t)i + 1 : (size_t)i;
            out += shards[d].pre[k];
            if (mode == 6 || mode == 7) continue;
            if (!write_gaf(a, shards[d].record[k], number, out)) return panic("unable to create file");
        }
        if (shards[d].code) {   // a panic (or a device error) ends the output where the sequential reference would have stopped
            out += shards[d].tail;
            err += shards[d].err;
            return finish(shards[d].code);
        }
    }
    auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
    err += "Done in " + std::to_string(secs) + ".\n";  // main.rs:322
    return finish(0);
}
