// csrc/model_text.hpp -- the text model format of mf_save_model / mf_load_model (mf/mf.cpp:4184-4278) on a fast path.
#ifndef MFB200_MODEL_TEXT_HPP
#define MFB200_MODEL_TEXT_HPP

namespace mfb200 {

// Writes exactly the bytes the reference writes ("f/m/n/k/b" header, then one line per row: "p<i> T v v ... " or
// "p<i> F 0 0 ... " for a row never seen in training).  0 on success, 1 if the file cannot be written.
int save_model_text(const char *path, int fun, int m, int n, int k, float b, const float *P, const float *Q);

// Reads the header; 0 on success.
struct ModelTextHeader {
    int fun, m, n, k;
    float b;
};
// Parses the whole file: *hdr, then P[m*k] and Q[n*k] into buffers obtained from alloc(count) (so that the caller
// decides where the factors live; they must be free()-able for mf_destroy_model).  0 on success.
int load_model_text(const char *path, ModelTextHeader *hdr, float *(*alloc)(unsigned long long count), float **P,
                    float **Q);

// read_problem (mf/mf.cpp:4143-4182): a text file of "u v r" triples -> nodes (layout of mf_node: int, int, float),
// m = max u + 1, n = max v + 1.  The file is read in one block and parsed by several threads, chunk by chunk at line
// boundaries; anything the chunked pass cannot take (a token that is not a number, a triple split over a chunk
// boundary) falls back to one sequential pass with the stream's semantics: reading stops at the first bad token.
struct TextNode {
    int u, v;
    float r;
};
// nodes come from alloc(count); returns 0, or 1 if the file cannot be read.  *nnz = number of triples read.
int read_problem_text(const char *path, TextNode *(*alloc)(unsigned long long count), TextNode **nodes, long long *nnz,
                      int *m, int *n);

}  // namespace mfb200
#endif
