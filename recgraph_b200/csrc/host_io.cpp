// FASTA ingestion, score tables, small numeric helpers with the reference's exact arithmetic.
#include <charconv>
#include <cstring>

#include "host.h"

namespace rg {

// score_matrix.rs:35-105 on the 6x6 code table (A,C,G,T,N,'-').
int make_score_matrix(int kind, int32_t m, int32_t x, rg_scoring* s) {
    if (!s) return RG_ERR_INVALID;
    static const int32_t h55[5][5] = {{91, -90, -25, -100, 0},
                                      {-90, 100, -100, -25, 0},
                                      {-25, -100, 100, -90, 0},
                                      {-100, -25, -90, 91, 0},
                                      {0, 0, 0, 0, 0}};
    static const int32_t h70[5][5] = {{91, -114, -31, -123, 0},
                                      {-114, 100, -125, -31, 0},
                                      {-31, -125, 100, -114, 0},
                                      {-123, -31, -144, 91, 0},
                                      {0, 0, 0, 0, 0}};
    switch (kind) {
        case 0:
        case 1:
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++) {
                    if (i == j)
                        s->score[i][j] = m;
                    else if ((i == CODE_GAP || j == CODE_GAP) && kind == 0)
                        s->score[i][j] = x * 2;  // score_matrix.rs:42
                    else
                        s->score[i][j] = x;
                }
            s->score[CODE_N][CODE_N] = x;  // score_matrix.rs:48
            break;
        case 2:
        case 3: {
            const int32_t(*t)[5] = kind == 2 ? h55 : h70;
            for (int i = 0; i < 5; i++)
                for (int j = 0; j < 5; j++) s->score[i][j] = t[i][j];
            for (int c = 0; c < 5; c++) {
                s->score[c][CODE_GAP] = -200;
                s->score[CODE_GAP][c] = -200;
            }
            break;
        }
        default: return RG_ERR_INVALID;
    }
    s->score[CODE_GAP][CODE_GAP] = 0;  // key removed in the reference; never looked up
    return RG_OK;
}

uint32_t bases_to_add(float b, float f, uint32_t L) {
    volatile float prod = f * (float)L;
    volatile float v = b + prod;
    float vv = v;
    if (!(vv > 0.0f)) return 0;            // negative and NaN saturate to 0 (`as usize`)
    if (vv >= 536870912.0f) return 1u << 29;  // anything this large means "no band"
    return (uint32_t)vv;
}

std::string f32_display(float v) {
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

// sequences.rs:5-45: multi-line records, '-' -> 'N', ASCII upper-case, name = header without '>'.
bool parse_fasta(const char* text, size_t len, std::vector<std::string>& names, std::vector<uint8_t>& codes,
                 std::vector<uint64_t>& off, std::string& err, int* status) {
    names.clear();
    codes.clear();
    off.assign(1, 0);
    const char* p = text;
    const char* end = text + len;
    size_t n_seqs = 0;
    bool open = false;  // current record has sequence bytes
    *status = RG_OK;
    auto close_rec = [&]() {
        if (open) {
            off.push_back(codes.size());
            n_seqs++;
            open = false;
        }
    };
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* le = eol;
        if (le > p && le[-1] == '\r') le--;
        if (le > p) {
            if (*p == '>') {
                names.emplace_back(p + 1, le);
                close_rec();
            } else {
                for (const char* c = p; c < le; c++) {
                    char ch = *c;
                    if (ch == '-') ch = 'N';
                    if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
                    uint8_t code;
                    switch (ch) {
                        case 'A': code = CODE_A; break;
                        case 'C': code = CODE_C; break;
                        case 'G': code = CODE_G; break;
                        case 'T': code = CODE_T; break;
                        case 'N': code = CODE_N; break;
                        default:
                            err = std::string("read character outside A,C,G,T,N: '") + ch + "'";
                            *status = RG_ERR_BAD_CHAR;
                            return false;
                    }
                    codes.push_back(code);
                }
                open = true;
            }
        }
        p = eol + 1;
    }
    close_rec();
    if (n_seqs != names.size()) {
        err = "wrong fasta file format";  // sequences.rs:41-43
        *status = RG_ERR_IO;
        return false;
    }
    return true;
}

}  // namespace rg
