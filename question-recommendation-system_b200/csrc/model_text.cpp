// csrc/model_text.cpp -- mf_save_model / mf_load_model (mf/mf.cpp:4184-4278) without the stream-per-float loops.
//
// The reference writes every float with `ofstream << float` (default format: "%g", six significant digits, of the value
// widened to double) and flushes after every row; it reads with `ifstream >> float`.  At the Netflix shape (64M floats)
// that is most of a minute.  Here rows are formatted and parsed in parallel (OpenMP) with std::to_chars / std::from_chars
// -- to_chars(general, 6) is specified to produce what "%.6g" produces, from_chars is correctly rounded like strtof -- and
// the file is written and read in large blocks.  The bytes written are the reference's bytes (tests/test_boundary.py
// compares them with the compiled reference's output), so models move freely between the two libraries.
#include "model_text.hpp"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace mfb200 {

namespace {

inline char *put_g6(char *p, float x) {  // "%g" of (double)x
    return std::to_chars(p, p + 32, (double)x, std::chars_format::general, 6).ptr;
}
inline char *put_int(char *p, long long v) { return std::to_chars(p, p + 24, v).ptr; }

// one side of the model, rows [0, rows): formatted in blocks of rows, each block split over the threads
int write_side(std::FILE *f, const float *M, int rows, int k, char prefix) {
    const int kBlock = 8192 > (1 << 21) / (k > 0 ? k : 1) ? 8192 : (1 << 21) / (k > 0 ? k : 1);  // >= 2M values per block
    const size_t row_cap = 16 + (size_t)k * 17;  // "-1.23457e-38 " is 13 bytes
    std::vector<std::string> part;
    for (int r0 = 0; r0 < rows; r0 += kBlock) {
        const int r1 = r0 + kBlock < rows ? r0 + kBlock : rows;
        int nthreads = 1;
        // small models stay on the calling thread: a parallel region costs far more than formatting a few rows
        const bool par = (long long)(r1 - r0) * k >= (1 << 20);
#pragma omp parallel if (par)
        {
#pragma omp single
            {
#ifdef _OPENMP
                nthreads = omp_get_num_threads();
#endif
                part.assign((size_t)nthreads, std::string());
            }
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            const int per = (r1 - r0 + nthreads - 1) / nthreads;
            const int lo = r0 + t * per, hi = lo + per < r1 ? lo + per : r1;
            std::string &out = part[(size_t)t];
            if (hi > lo) out.resize((size_t)(hi - lo) * row_cap);
            char *p = out.empty() ? nullptr : &out[0];
            for (int i = lo; i < hi; i++) {
                const float *row = M + (long long)i * k;
                *p++ = prefix;
                p = put_int(p, i);
                *p++ = ' ';
                if (std::isnan(row[0])) {  // mf/mf.cpp:4203-4208
                    *p++ = 'F';
                    *p++ = ' ';
                    for (int d = 0; d < k; d++) {
                        *p++ = '0';
                        *p++ = ' ';
                    }
                } else {
                    *p++ = 'T';
                    *p++ = ' ';
                    for (int d = 0; d < k; d++) {
                        p = put_g6(p, row[d]);
                        *p++ = ' ';
                    }
                }
                *p++ = '\n';
            }
            if (hi > lo) out.resize((size_t)(p - &out[0]));
        }
        for (const std::string &s : part)
            if (!s.empty() && std::fwrite(s.data(), 1, s.size(), f) != s.size()) return 1;
    }
    return 0;
}

inline const char *skip_ws(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    return p;
}
inline const char *skip_token(const char *p, const char *e) {
    while (p < e && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    return p;
}

}  // namespace

int save_model_text(const char *path, int fun, int m, int n, int k, float b, const float *P, const float *Q) {
    std::FILE *f = std::fopen(path, "wb");
    if (!f) return 1;
    std::vector<char> big(8u << 20);
    std::setvbuf(f, big.data(), _IOFBF, big.size());
    char hdr[160], *p = hdr;
    auto line = [&](char tag, long long v) {
        *p++ = tag;
        *p++ = ' ';
        p = put_int(p, v);
        *p++ = '\n';
    };
    line('f', fun);
    line('m', m);
    line('n', n);
    line('k', k);
    *p++ = 'b';
    *p++ = ' ';
    p = put_g6(p, b);
    *p++ = '\n';
    int rc = std::fwrite(hdr, 1, (size_t)(p - hdr), f) != (size_t)(p - hdr);
    if (!rc) rc = write_side(f, P, m, k, 'p');
    if (!rc) rc = write_side(f, Q, n, k, 'q');
    if (std::fclose(f) != 0) rc = 1;
    return rc ? 1 : 0;
}

int load_model_text(const char *path, ModelTextHeader *hdr, float *(*alloc)(unsigned long long count), float **P,
                    float **Q) {
    *P = *Q = nullptr;
    std::FILE *f = std::fopen(path, "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END);
    const long long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)(size > 0 ? size : 0));
    const bool ok = size <= 0 || std::fread(buf.data(), 1, (size_t)size, f) == (size_t)size;
    std::fclose(f);
    if (!ok) return 1;
    const char *p = buf.data(), *e = buf.data() + buf.size();
    // header: five "tag value" pairs, whitespace separated (mf/mf.cpp:4240-4241)
    long long iv[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        p = skip_token(skip_ws(p, e), e);
        p = skip_ws(p, e);
        p = std::from_chars(p, e, iv[i]).ptr;
    }
    p = skip_token(skip_ws(p, e), e);
    p = skip_ws(p, e);
    float b = 0.f;
    p = std::from_chars(p, e, b).ptr;
    hdr->fun = (int)iv[0];
    hdr->m = (int)iv[1];
    hdr->n = (int)iv[2];
    hdr->k = (int)iv[3];
    hdr->b = b;
    if (hdr->m < 0 || hdr->n < 0 || hdr->k < 0) return 1;
    const long long rows = (long long)hdr->m + hdr->n;
    const int k = hdr->k;
    *P = alloc((unsigned long long)hdr->m * k);
    *Q = alloc((unsigned long long)hdr->n * k);
    // one row per line: find the line starts, then parse the lines in parallel
    p = skip_ws(p, e);
    std::vector<const char *> start;
    start.reserve((size_t)rows + 1);
    for (const char *q = p; q < e && (long long)start.size() < rows;) {
        start.push_back(q);
        const char *nl = (const char *)std::memchr(q, '\n', (size_t)(e - q));
        q = nl ? skip_ws(nl, e) : e;
    }
    const long long have = (long long)start.size();
    const float nan = std::numeric_limits<float>::quiet_NaN();
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (have * k >= (1 << 20))
    for (long long i = 0; i < have; i++) {
        float *row = i < hdr->m ? *P + i * k : *Q + (i - hdr->m) * k;
        const char *q = start[(size_t)i], *le = i + 1 < have ? start[(size_t)i + 1] : e;
        q = skip_ws(skip_token(q, le), le);  // "p<i>"
        const bool unseen = q < le && *q == 'F';  // mf/mf.cpp:4262-4267
        q = skip_token(q, le);
        for (int d = 0; d < k; d++) {
            q = skip_ws(q, le);
            if (unseen) {
                q = skip_token(q, le);
                row[d] = nan;
            } else {
                float v = 0.f;
                auto r = std::from_chars(q, le, v);
                if (r.ec == std::errc::invalid_argument) bad = 1;
                // out-of-range text (a denormal or an overflow written by hand) keeps strtof's answer
                if (r.ec == std::errc::result_out_of_range) v = std::strtof(std::string(q, r.ptr).c_str(), nullptr);
                row[d] = v;
                q = r.ptr > q ? r.ptr : skip_token(q, le);
            }
        }
    }
    (void)bad;  // a malformed file leaves the reference with garbage too (no error path at mf/mf.cpp:4254-4270)
    return 0;
}

namespace {

// Parses triples from [p, e).  Returns false at the first token that is not a number (what is in `out` then is
// everything before it, like the stream loop `for (mf_node N; f >> N.u >> N.v >> N.r;)`); *partial = a triple was
// cut off by the end of the range.
bool parse_triples(const char *p, const char *e, std::vector<TextNode> &out, bool *partial) {
    *partial = false;
    for (;;) {
        p = skip_ws(p, e);
        if (p >= e) return true;
        TextNode N;
        int *iv[2] = {&N.u, &N.v};
        for (int i = 0; i < 2; i++) {
            if (p < e && *p == '+') p++;
            auto r = std::from_chars(p, e, *iv[i]);
            if (r.ec != std::errc()) return false;
            p = skip_ws(r.ptr, e);
            if (p >= e) {
                *partial = true;
                return true;
            }
        }
        if (p < e && *p == '+') p++;
        auto r = std::from_chars(p, e, N.r);
        if (r.ec != std::errc()) return false;
        p = r.ptr;
        out.push_back(N);
    }
}

}  // namespace

int read_problem_text(const char *path, TextNode *(*alloc)(unsigned long long count), TextNode **nodes, long long *nnz,
                      int *m, int *n) {
    *nodes = nullptr;
    *nnz = 0;
    *m = *n = 0;
    std::FILE *f = std::fopen(path, "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END);
    const long long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)(size > 0 ? size : 0));
    const bool ok = size <= 0 || std::fread(buf.data(), 1, (size_t)size, f) == (size_t)size;
    std::fclose(f);
    if (!ok) return 1;
    const char *b = buf.data(), *e = b + buf.size();
    // chunks of ~8 MB that start after a newline
    std::vector<const char *> cut;
    cut.push_back(b);
    const long long kChunk = 8ll << 20;
    for (long long off = kChunk; off < size; off += kChunk) {
        const char *nl = (const char *)std::memchr(b + off, '\n', (size_t)(size - off));
        if (!nl) break;
        if (nl + 1 > cut.back()) cut.push_back(nl + 1);
    }
    cut.push_back(e);
    const int nchunks = (int)cut.size() - 1;
    std::vector<std::vector<TextNode>> part((size_t)nchunks);
    int trouble = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : trouble) if (nchunks > 1)
    for (int c = 0; c < nchunks; c++) {
        bool partial = false;
        part[(size_t)c].reserve((size_t)((cut[(size_t)c + 1] - cut[(size_t)c]) / 8));
        if (!parse_triples(cut[(size_t)c], cut[(size_t)c + 1], part[(size_t)c], &partial) || partial) trouble = 1;
    }
    if (trouble) {  // rare: let one pass over the whole file decide where reading stops
        part.assign(1, std::vector<TextNode>());
        bool partial = false;
        parse_triples(b, e, part[0], &partial);
    }
    long long total = 0;
    std::vector<long long> first(part.size() + 1, 0);
    for (size_t c = 0; c < part.size(); c++) {
        first[c + 1] = first[c] + (long long)part[c].size();
        total = first[c + 1];
    }
    TextNode *out = alloc((unsigned long long)(total > 0 ? total : 1));
    int mm = 0, nn = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(max : mm, nn) if (part.size() > 1)
    for (long long c = 0; c < (long long)part.size(); c++) {
        const std::vector<TextNode> &v = part[(size_t)c];
        if (!v.empty()) std::memcpy(out + first[(size_t)c], v.data(), sizeof(TextNode) * v.size());
        for (const TextNode &N : v) {  // mf/mf.cpp:4170-4173
            if (N.u + 1 > mm) mm = N.u + 1;
            if (N.v + 1 > nn) nn = N.v + 1;
        }
    }
    *nodes = out;
    *nnz = total;
    *m = mm;
    *n = nn;
    return 0;
}

}  // namespace mfb200
