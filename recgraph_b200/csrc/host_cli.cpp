// `recgraph` command line (main.rs:25-329 + args_parser.rs:3-202) on top of the public C ABI only.
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/recgraph_b200.h"

namespace {

struct Args {  // args_parser.rs:3-147: same flags and defaults
    std::string sequence_path, graph_path, out_file = "standard output";
    int alignment_mode = 0, match_score = 2, mismatch_score = 4, gap_open = 4, gap_extension = 2, base_rec_cost = 4;
    std::string matrix = "none", amb_strand = "false";
    float multi_rec_cost = 0.1f, rec_band_width = 1.0f, extra_f = 0.01f;
    int extra_b = 1;
};

struct OptName {
    char s;
    const char* l;
};
const OptName OPTS[] = {{'o', "out_file"}, {'m', "aln-mode"}, {'M', "match"}, {'X', "mismatch"}, {'t', "matrix"},
                        {'O', "gap-open"}, {'E', "gap-ext"}, {'r', "multi-rec-cost"}, {'R', "base-rec-cost"},
                        {'B', "rec-band-width"}, {'s', "amb-strand"}, {'b', "extra-b"}, {'f', "extra-f"}};

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

}  // namespace

extern "C" int rg_cli_main(int argc, const char** argv, char** out_text, char** err_text) {
    auto t0 = std::chrono::steady_clock::now();
    std::string out, err;
    int rc = 0;
    Args a;
    rg_ctx* ctx = nullptr;
    rg_reads reads;
    memset(&reads, 0, sizeof reads);
    auto finish = [&](int code) {
        if (ctx) rg_destroy(ctx);
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
    rc = rg_init(0, &ctx);
    if (rc != RG_OK) {
        err += std::string("recgraph_b200: ") + rg_strerror(rc) + "\n";
        return finish(3);
    }
    rc = rg_load_gfa_file(ctx, a.graph_path.c_str());
    if (rc != RG_OK) return panic(rg_last_error(ctx));

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
    rg_set_scoring(ctx, &sc);

    const int mode = a.alignment_mode;
    if (mode < 0 || mode > 9) return panic("Alignment mode must be in [0..9]");
    if (a.amb_strand == "true") {
        err += "recgraph_b200: -s true (ambiguous strand, experimental in the reference) is not implemented yet\n";
        return finish(3);
    }
    rg_batch_result res;
    rc = rg_align_batch(ctx, mode, reads.n_reads, reads.codes, reads.off, &res);
    if (rc == RG_ERR_REF_PANIC) return panic(rg_last_error(ctx));
    if (rc != RG_OK) {
        err += std::string("recgraph_b200: ") + rg_strerror(rc) + ": " + rg_last_error(ctx) + "\n";
        return finish(3);
    }
    std::vector<char> buf(1 << 16);
    for (int32_t i = 0; i < reads.n_reads; i++) {
        if (res.reads[i].status & RG_READ_REF_PANIC)
            return panic("reference panic while aligning read " + std::to_string(i + 1) + " (see DESIGN.md, reference quirks)");
        if (res.reads[i].status & RG_READ_TRACE_OVERFLOW) {
            err += "recgraph_b200: trace buffers overflowed for read " + std::to_string(i + 1) + "\n";
            return finish(3);
        }
        uint32_t len = (uint32_t)(reads.off[i + 1] - reads.off[i]);
        int64_t need = rg_format_gaf(ctx, mode, &res, i, reads.names[i], len, 0, buf.data(), buf.size());
        if (need < 0) return finish(3);
        if ((size_t)need >= buf.size()) {
            buf.resize((size_t)need + 1);
            rg_format_gaf(ctx, mode, &res, i, reads.names[i], len, 0, buf.data(), buf.size());
        }
        size_t number = mode <= 3 ? (size_t)i + 1 : (size_t)i;
        std::string text(buf.data(), (size_t)need);
        // warning lines are println!'d to stdout by the reference even with -o; only the record goes to the file
        size_t cut = text.size() > 1 ? text.rfind('\n', text.size() - 2) : std::string::npos;
        if (a.out_file != "standard output" && cut != std::string::npos) {
            out += text.substr(0, cut + 1);
            text.erase(0, cut + 1);
        }
        if (!write_gaf(a, text, number, out)) return panic("unable to create file");
    }
    auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
    err += "Done in " + std::to_string(secs) + ".\n";  // main.rs:322
    return finish(0);
}
