// Host-only pieces behind the C ABI: MD5 (for sourmash md5sum), input normalisation, FASTA reader.
#pragma once
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "util.cuh"

namespace ks {

// ---- MD5 (RFC 1321), only used to label sketches the way sourmash does --------------------------------
class Md5 {
  public:
    Md5() { a_ = 0x67452301u; b_ = 0xefcdab89u; c_ = 0x98badcfeu; d_ = 0x10325476u; len_ = 0; fill_ = 0; }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        len_ += n;
        while (n) {
            size_t take = 64 - fill_ < n ? 64 - fill_ : n;
            memcpy(buf_ + fill_, p, take);
            fill_ += take; p += take; n -= take;
            if (fill_ == 64) { block(buf_); fill_ = 0; }
        }
    }
    void hex(char out[33]) {
        uint64_t bits = len_ * 8;
        uint8_t pad = 0x80;
        update(&pad, 1);
        uint8_t z = 0;
        while (fill_ != 56) update(&z, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (8 * i));
        update(lb, 8);
        uint32_t w[4] = {a_, b_, c_, d_};
        static const char* hx = "0123456789abcdef";
        for (int i = 0; i < 16; i++) {
            uint8_t v = (uint8_t)(w[i / 4] >> (8 * (i % 4)));
            out[2 * i] = hx[v >> 4];
            out[2 * i + 1] = hx[v & 15];
        }
        out[32] = 0;
    }

  private:
    static uint32_t rl(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t m[16];
        for (int i = 0; i < 16; i++)
            m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
        uint32_t a = a_, b = b_, c = c_, d = d_;
        for (int i = 0; i < 64; i++) {
            uint32_t f; int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            uint32_t t = d; d = c; c = b;
            b = b + rl(a + f + K[i] + m[g], S[i]);
            a = t;
        }
        a_ += a; b_ += b; c_ += c; d_ += d;
    }
    uint32_t a_, b_, c_, d_;
    uint64_t len_;
    uint8_t buf_[64];
    size_t fill_;
};

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

struct InvalidResidue { uint32_t ch; uint64_t pos; uint64_t protein; };

// Normalise one sequence into out (appends).  Mirrors to_uppercase (src/rust/index.rs:1000) followed by
// validate_and_resolve (src/rust/aminoacid.rs:74-105); `raw`: upper-case only (what sourmash add_protein sees on the
// `kmerseek search` path).  Returns false and fills `bad` on an invalid residue.  B/Z/J: resolved from the seed and the
// residue's position in its own sequence (see resolve_ambiguous in ingest.hpp; the rule is restated here).
inline bool normalize_into(const char* s, uint64_t len, uint64_t protein_index, uint64_t ambig_seed, bool raw,
                           std::vector<uint8_t>& out, InvalidResidue* bad) {
    const size_t start = out.size();
    for (uint64_t i = 0; i < len; i++) {
        uint8_t c = (uint8_t)s[i];
        if (c >= 'a' && c <= 'z') c -= 32;
        if (raw) { out.push_back(c); continue; }
        if (c == '*') { out.push_back(c); break; }
        bool ok = false;
        switch (c) {
            case 'A': case 'C': case 'D': case 'E': case 'F': case 'G': case 'H': case 'I': case 'K': case 'L':
            case 'M': case 'N': case 'P': case 'Q': case 'R': case 'S': case 'T': case 'V': case 'W': case 'Y':
            case 'X': case 'U': case 'O':
                ok = true; break;
            case 'B': case 'Z': case 'J': {
                const uint64_t n = out.size() - start;
                const uint64_t r = splitmix64(ambig_seed ^ n) & 1;
                c = c == 'B' ? (r ? 'N' : 'D') : c == 'Z' ? (r ? 'Q' : 'E') : (r ? 'L' : 'I');
                ok = true; break;
            }
            default: break;
        }
        if (!ok) {
            bad->ch = c; bad->pos = out.size() - start + 1; bad->protein = protein_index;
            return false;
        }
        out.push_back(c);
    }
    return true;
}

// ---- decompression of whole files through the system's runtime libraries -------------------------------
// niffler (the reference's reader, src/rust/index.rs:920) sniffs gzip / zstd / bzip2 / xz.  gzip goes through zlib
// (headers present).  The other three only have their runtime .so in this image, no headers, so the few entry
// points needed are declared here and bound with dlopen; if a library is missing the file is reported as ParseError.
struct DlLib {
    void* h = nullptr;
    explicit DlLib(const char* const* names) {
        for (; *names && !h; names++) h = dlopen(*names, RTLD_NOW | RTLD_LOCAL);
    }
    template <class F>
    F sym(const char* n) const { return h ? reinterpret_cast<F>(dlsym(h, n)) : nullptr; }
};

inline std::string slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) fail(KS_ERR_PARSE, std::string("Parse error: cannot open ") + path);
    std::string out;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return out;
}

inline std::string decompress_zstd(const std::string& in) {
    struct InBuf { const void* src; size_t size; size_t pos; };
    struct OutBuf { void* dst; size_t size; size_t pos; };
    static const char* names[] = {"libzstd.so.1", "libzstd.so", nullptr};
    static DlLib lib(names);
    auto create = lib.sym<void* (*)()>("ZSTD_createDStream");
    auto init = lib.sym<size_t (*)(void*)>("ZSTD_initDStream");
    auto step = lib.sym<size_t (*)(void*, OutBuf*, InBuf*)>("ZSTD_decompressStream");
    auto is_err = lib.sym<unsigned (*)(size_t)>("ZSTD_isError");
    auto destroy = lib.sym<size_t (*)(void*)>("ZSTD_freeDStream");
    if (!create || !init || !step || !is_err || !destroy) fail(KS_ERR_PARSE, "Parse error: zstd input but libzstd is not available");
    void* ds = create();
    init(ds);
    std::string out;
    std::vector<char> buf(1 << 20);
    InBuf ib{in.data(), in.size(), 0};
    while (ib.pos < ib.size) {
        OutBuf ob{buf.data(), buf.size(), 0};
        const size_t r = step(ds, &ob, &ib);
        if (is_err(r)) { destroy(ds); fail(KS_ERR_PARSE, "Parse error: corrupt zstd stream"); }
        out.append(buf.data(), ob.pos);
        if (r == 0 && ib.pos >= ib.size) break;
    }
    destroy(ds);
    return out;
}

inline std::string decompress_bz2(const std::string& in) {
    struct BzStream {
        char* next_in; unsigned avail_in, total_in_lo32, total_in_hi32;
        char* next_out; unsigned avail_out, total_out_lo32, total_out_hi32;
        void* state; void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
    };
    static const char* names[] = {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so", nullptr};
    static DlLib lib(names);
    auto init = lib.sym<int (*)(BzStream*, int, int)>("BZ2_bzDecompressInit");
    auto step = lib.sym<int (*)(BzStream*)>("BZ2_bzDecompress");
    auto end = lib.sym<int (*)(BzStream*)>("BZ2_bzDecompressEnd");
    if (!init || !step || !end) fail(KS_ERR_PARSE, "Parse error: bzip2 input but libbz2 is not available");
    BzStream st{};
    if (init(&st, 0, 0) != 0) fail(KS_ERR_PARSE, "Parse error: bzip2 init failed");
    std::string out;
    std::vector<char> buf(1 << 20);
    st.next_in = const_cast<char*>(in.data());
    st.avail_in = (unsigned)in.size();
    for (;;) {
        st.next_out = buf.data();
        st.avail_out = (unsigned)buf.size();
        const int r = step(&st);
        out.append(buf.data(), buf.size() - st.avail_out);
        if (r == 4) break;  // BZ_STREAM_END
        if (r != 0 || (st.avail_in == 0 && st.avail_out != 0)) { end(&st); fail(KS_ERR_PARSE, "Parse error: corrupt bzip2 stream"); }
    }
    end(&st);
    return out;
}

inline std::string decompress_xz(const std::string& in) {
    struct LzmaStream {
        const uint8_t* next_in; size_t avail_in; uint64_t total_in;
        uint8_t* next_out; size_t avail_out; uint64_t total_out;
        const void* allocator; void* internal;
        void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
        uint64_t reserved_int1, reserved_int2; size_t reserved_int3, reserved_int4;
        int reserved_enum1, reserved_enum2;
    };
    static const char* names[] = {"liblzma.so.5", "liblzma.so", nullptr};
    static DlLib lib(names);
    auto init = lib.sym<int (*)(LzmaStream*, uint64_t, uint32_t)>("lzma_stream_decoder");
    auto step = lib.sym<int (*)(LzmaStream*, int)>("lzma_code");
    auto end = lib.sym<void (*)(LzmaStream*)>("lzma_end");
    if (!init || !step || !end) fail(KS_ERR_PARSE, "Parse error: xz input but liblzma is not available");
    LzmaStream st{};
    if (init(&st, UINT64_MAX, 0x08 /* LZMA_CONCATENATED */) != 0) fail(KS_ERR_PARSE, "Parse error: xz init failed");
    std::string out;
    std::vector<uint8_t> buf(1 << 20);
    st.next_in = reinterpret_cast<const uint8_t*>(in.data());
    st.avail_in = in.size();
    for (;;) {
        st.next_out = buf.data();
        st.avail_out = buf.size();
        const int r = step(&st, st.avail_in == 0 ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
        out.append(reinterpret_cast<char*>(buf.data()), buf.size() - st.avail_out);
        if (r == 1) break;  // LZMA_STREAM_END
        if (r != 0) { end(&st); fail(KS_ERR_PARSE, "Parse error: corrupt xz stream"); }
    }
    end(&st);
    return out;
}

inline std::string decompress_gzip_or_plain(const char* path) {
    gzFile f = gzopen(path, "rb");
    if (!f) fail(KS_ERR_PARSE, std::string("Parse error: cannot open ") + path);
    gzbuffer(f, 1 << 20);
    std::string out;
    std::vector<char> buf(1 << 20);
    int n;
    while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), n);
    gzclose(f);
    if (n < 0) fail(KS_ERR_PARSE, "Parse error: read failed (corrupt gzip stream?)");
    return out;
}

// FASTA records the way needletail hands them to process_fasta (src/rust/index.rs:920-935): id = the header
// line without '>', sequence = the following lines joined, CR/LF removed.  Plain, gzip, zstd, bzip2 or xz input,
// sniffed from the magic bytes like niffler does.
inline void read_fasta(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    FILE* probe = fopen(path, "rb");
    if (!probe) fail(KS_ERR_PARSE, std::string("Parse error: cannot open ") + path);
    unsigned char magic[6] = {0};
    size_t got = fread(magic, 1, 6, probe);
    fclose(probe);
    if (got == 0) fail(KS_ERR_PARSE, "Parse error: empty file");
    std::string data;
    if (got >= 4 && magic[0] == 0x28 && magic[1] == 0xb5 && magic[2] == 0x2f && magic[3] == 0xfd) data = decompress_zstd(slurp(path));
    else if (got >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') data = decompress_bz2(slurp(path));
    else if (got >= 6 && magic[0] == 0xfd && magic[1] == '7' && magic[2] == 'z' && magic[3] == 'X' && magic[4] == 'Z') data = decompress_xz(slurp(path));
    else data = decompress_gzip_or_plain(path);
    bool first = true, in_record = false;
    size_t pos = 0;
    while (pos < data.size()) {
        size_t nl = data.find('\n', pos);
        if (nl == std::string::npos) nl = data.size();
        size_t end = nl;
        if (end > pos && data[end - 1] == '\r') end--;
        if (first) {
            if (end == pos) { pos = nl + 1; continue; }
            if (data[pos] != '>') fail(KS_ERR_PARSE, "Parse error: expected '>' at the start of a FASTA record");
            first = false;
        }
        if (end > pos && data[pos] == '>') {
            names.emplace_back(data, pos + 1, end - pos - 1);
            seqs.emplace_back();
            in_record = true;
        } else if (in_record) {
            seqs.back().append(data, pos, end - pos);
        }
        pos = nl + 1;
    }
    if (first) fail(KS_ERR_PARSE, "Parse error: empty file");
}

}  // namespace ks
