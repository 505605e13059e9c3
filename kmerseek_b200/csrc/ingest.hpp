// Multi-threaded FASTA ingest: file bytes -> normalised residues + offsets + names, written straight into their final
// buffers.  Replaces needletail::parse_fastx_file + the per-record Vec copies (src/rust/index.rs:920-935), `to_uppercase`
// (:1000) and AminoAcidAmbiguity::validate_and_resolve (src/rust/aminoacid.rs:74-105) for a whole file at once.
//
// The reference reads records one by one on one thread and hands batches of 1000 to rayon (index.rs:927-941).  Here the
// file (mmap'ed when plain, decompressed into one buffer otherwise) is cut into one byte range per thread at record
// boundaries; pass 1 counts records / residues / header bytes per range (memchr speed), a prefix sum gives every range its
// place in the outputs, pass 2 normalises into place.  The file is never copied into per-record strings.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <thread>

#include "host_util.hpp"

namespace ks {

// How sequences are normalised on the way in.
//   KMERSEEK: what ProteomeIndex::process_fasta does (upper-case, stop at the first '*', B/Z/J resolved, anything outside
//             the 27 accepted letters is an error) -- the index path of the Rust crate.
//   SOURMASH: what `kmerseek search` does to both sides (src/python/kmerseek/sketch.py:28-40 -> branchwater manysketch ->
//             sourmash add_protein): upper-case only; no validation, no truncation, no resolution.
enum NormalizeMode { NORM_KMERSEEK = 0, NORM_SOURMASH = 1 };

struct FastaBytes {  // the whole (decompressed) file
    const char* data = nullptr;
    size_t size = 0;
    std::string owned;
    void* map = nullptr;
    size_t map_size = 0;
    FastaBytes() = default;
    FastaBytes(const FastaBytes&) = delete;
    FastaBytes& operator=(const FastaBytes&) = delete;
    ~FastaBytes() { if (map) munmap(map, map_size); }
};

// Plain files are mapped; gzip / zstd / bzip2 / xz (sniffed from the magic bytes like niffler does) are decompressed.
inline void load_fasta_bytes(const char* path, FastaBytes* fb) {
    FILE* probe = fopen(path, "rb");
    if (!probe) fail(KS_ERR_PARSE, std::string("Parse error: cannot open ") + path);
    unsigned char magic[6] = {0};
    const size_t got = fread(magic, 1, 6, probe);
    fclose(probe);
    if (got == 0) fail(KS_ERR_PARSE, "Parse error: empty file");
    if (got >= 4 && magic[0] == 0x28 && magic[1] == 0xb5 && magic[2] == 0x2f && magic[3] == 0xfd) fb->owned = decompress_zstd(slurp(path));
    else if (got >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') fb->owned = decompress_bz2(slurp(path));
    else if (got >= 6 && magic[0] == 0xfd && magic[1] == '7' && magic[2] == 'z' && magic[3] == 'X' && magic[4] == 'Z') fb->owned = decompress_xz(slurp(path));
    else if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) fb->owned = decompress_gzip_or_plain(path);
    else {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) fail(KS_ERR_PARSE, std::string("Parse error: cannot open ") + path);
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); fb->owned = slurp(path); }
        else {
            // (no MAP_POPULATE: one thread filling the page table of a 200 MB file costs more than the parser threads
            // faulting their own ranges in -- 13 + 19 ms against 25 ms for map + count on 8 cores)
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            close(fd);
            if (m == MAP_FAILED) fb->owned = slurp(path);
            else {
                fb->map = m; fb->map_size = (size_t)st.st_size;
                madvise(m, fb->map_size, MADV_SEQUENTIAL);
                fb->data = (const char*)m; fb->size = fb->map_size;
                return;
            }
        }
    }
    fb->data = fb->owned.data();
    fb->size = fb->owned.size();
}

struct ResidueClass {  // per input byte: the upper-cased byte and what to do with it
    uint8_t up[256];
    uint8_t special[256];  // KMERSEEK mode: 0 plain valid residue, 1 '*', 2 B/Z/J, 3 invalid
    uint16_t both[256];    // up | special << 8: one load per byte on the fast path
    ResidueClass() {
        for (int i = 0; i < 256; i++) {
            uint8_t c = (uint8_t)i;
            if (c >= 'a' && c <= 'z') c -= 32;
            up[i] = c;
            uint8_t s = 3;
            switch (c) {
                case 'A': case 'C': case 'D': case 'E': case 'F': case 'G': case 'H': case 'I': case 'K': case 'L':
                case 'M': case 'N': case 'P': case 'Q': case 'R': case 'S': case 'T': case 'V': case 'W': case 'Y':
                case 'X': case 'U': case 'O': s = 0; break;
                case '*': s = 1; break;
                case 'B': case 'Z': case 'J': s = 2; break;
                default: break;
            }
            special[i] = s;
            both[i] = (uint16_t)(c | (s << 8));
        }
    }
};

// B/Z/J: one of the outcomes the reference can produce (it draws at random per residue, aminoacid.rs:45-54), fixed by
// the seed and the residue's position in ITS sequence -- not by the record's index in the file, so that the same
// sequence resolves the same way wherever it appears (a query that is a copy of a target matches it fully).
inline uint8_t resolve_ambiguous(uint8_t c, uint64_t ambig_seed, uint64_t pos) {
    const uint64_t r = splitmix64(ambig_seed ^ pos) & 1;
    return c == 'B' ? (r ? 'N' : 'D') : c == 'Z' ? (r ? 'Q' : 'E') : (r ? 'L' : 'I');
}

// 5-bit codes, 8 residues per 5 bytes, over [g0, g1) groups (pack_residues of sketch.cu, one group range per thread).
// Returns false when a byte has no code.
struct PackCodes {  // residue byte -> 5-bit code; 0x80 marks a byte that has none
    uint8_t code[256];
    PackCodes() {
        for (int i = 0; i < 256; i++) code[i] = (i >= 'A' && i <= 'Z') ? (uint8_t)(i - 'A' + 1) : i == '*' ? 27 : 0x80;
    }
};

inline bool pack_groups(const uint8_t* res, uint64_t n, uint8_t* out, uint64_t g0, uint64_t g1) {
    static const PackCodes pc;
    uint32_t bad = 0;
    const uint64_t full = std::min<uint64_t>(g1, n / 8);  // groups with all 8 residues
    for (uint64_t g = g0; g < full; g++) {
        const uint8_t* r = res + g * 8;
        const uint32_t c0 = pc.code[r[0]], c1 = pc.code[r[1]], c2 = pc.code[r[2]], c3 = pc.code[r[3]], c4 = pc.code[r[4]],
                       c5 = pc.code[r[5]], c6 = pc.code[r[6]], c7 = pc.code[r[7]];
        bad |= c0 | c1 | c2 | c3 | c4 | c5 | c6 | c7;
        const uint64_t v = (uint64_t)(c0 & 31u) | ((uint64_t)(c1 & 31u) << 5) | ((uint64_t)(c2 & 31u) << 10) |
                           ((uint64_t)(c3 & 31u) << 15) | ((uint64_t)(c4 & 31u) << 20) | ((uint64_t)(c5 & 31u) << 25) |
                           ((uint64_t)(c6 & 31u) << 30) | ((uint64_t)(c7 & 31u) << 35);
        uint8_t* o = out + g * 5;
        const uint32_t lo = (uint32_t)v;
        memcpy(o, &lo, 4);
        o[4] = (uint8_t)(v >> 32);
    }
    for (uint64_t g = std::max(g0, full); g < g1; g++) {  // the last, partial group
        uint64_t v = 0;
        const uint64_t base = g * 8;
        const int cnt = (int)std::min<uint64_t>(8, n - base);
        for (int i = 0; i < cnt; i++) {
            const uint32_t c = pc.code[res[base + i]];
            bad |= c;
            v |= (uint64_t)(c & 31u) << (5 * i);
        }
        uint8_t* o = out + g * 5;
        o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); o[3] = (uint8_t)(v >> 24); o[4] = (uint8_t)(v >> 32);
    }
    return (bad & 0x80u) == 0;
}

struct ParsedFasta {  // outputs of the parse; the buffers are the caller's (sized from the counts)
    uint64_t n_rec = 0, n_res = 0, name_bytes = 0;
};

struct FastaChunk {
    size_t begin = 0, end = 0;                      // byte range; begin sits on a record start
    uint64_t n_rec = 0, n_res = 0, name_bytes = 0;  // pass 1
    uint64_t rec0 = 0, res0 = 0, name0 = 0;         // exclusive prefix over the chunks
    bool bad = false;
    InvalidResidue bad_res{0, 0, 0};
    bool pack_ok = true;                            // pass 2: every byte of the chunk's own groups had a 5-bit code
};

inline int ingest_threads(size_t bytes) {
    static const int forced = [] { const char* e = getenv("KS_INGEST_THREADS"); return e ? atoi(e) : 0; }();  // read once
    int t = forced > 0 ? forced : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    const size_t by_size = bytes / (1u << 20) + 1;  // at least ~1 MB of file per thread
    return (int)std::min<size_t>((size_t)t, by_size);
}

template <class F>
inline void parallel_chunks(int n, F&& f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int i = 1; i < n; i++) th.emplace_back([&f, i] { f(i); });
    f(0);
    for (auto& t : th) t.join();
}

// Lines of [p, end): calls on_header(b, e) for a line that starts with '>', on_seq(b, e) for any other (a trailing CR is
// stripped from both), in file order.
template <class H, class S>
inline void for_each_line(const char* data, size_t p, size_t end, H&& on_header, S&& on_seq) {
    while (p < end) {
        const char* nlp = (const char*)memchr(data + p, '\n', end - p);
        const size_t nl = nlp ? (size_t)(nlp - data) : end;
        size_t e = nl;
        if (e > p && data[e - 1] == '\r') e--;
        if (e > p && data[p] == '>') on_header(p + 1, e);
        else on_seq(p, e);
        p = nl + 1;
    }
}

class FastaParser {
  public:
    FastaParser(const char* data, size_t size, NormalizeMode mode, uint64_t ambig_seed)
        : d_(data), n_(size), mode_(mode), seed_(ambig_seed) {
        // the first non-blank line must be a header (needletail's error otherwise)
        size_t p = 0;
        while (p < n_ && (d_[p] == '\n' || d_[p] == '\r')) p++;
        if (p >= n_) fail(KS_ERR_PARSE, "Parse error: empty file");
        if (d_[p] == '@') fail(KS_ERR_PARSE, "Parse error: FASTQ input is not supported (protein FASTA expected)");
        if (d_[p] != '>') fail(KS_ERR_PARSE, "Parse error: expected '>' at the start of a FASTA record");
        start_ = p;
        const int t = ingest_threads(n_ - p);
        chunks_.resize(t);
        size_t prev = p;
        for (int i = 0; i < t; i++) {
            chunks_[i].begin = prev;
            size_t cut = i + 1 == t ? n_ : record_start_at_or_after(p + (n_ - p) / t * (size_t)(i + 1));
            if (cut < prev) cut = prev;
            chunks_[i].end = cut;
            prev = cut;
        }
    }

    // Pass 1: counts.  Returns the totals the caller sizes its buffers from.
    ParsedFasta count() {
        parallel_chunks((int)chunks_.size(), [&](int i) { count_chunk(chunks_[i]); });
        ParsedFasta t;
        for (auto& c : chunks_) {
            c.rec0 = t.n_rec; c.res0 = t.n_res; c.name0 = t.name_bytes;
            t.n_rec += c.n_rec; t.n_res += c.n_res; t.name_bytes += c.name_bytes;
        }
        return t;
    }

    // Pass 2: normalised residues into `res` (n_res bytes), offsets[n_rec + 1], NUL-terminated names into `names`
    // (name_bytes bytes) with name_off[n_rec].  Returns false (and fills `bad`: lowest protein index, then position) on
    // an invalid residue.
    // `packed` (optional, packed_bytes(n_res) bytes): the 5-bit upload copy is written in the same pass -- every thread packs
    // the 8-residue groups that lie inside its own residue range while they are still in its cache (a separate pass read
    // the 200 MB of a Swiss-Prot-sized proteome back from memory); the few groups that straddle two threads' ranges and the
    // last, partial one are packed here afterwards.  *packed_ok = false when a byte has no code (the caller then uploads
    // the residues themselves).
    bool fill(uint8_t* res, uint64_t* offsets, char* names, uint64_t* name_off, InvalidResidue* bad, uint8_t* packed = nullptr,
              bool* packed_ok = nullptr) {
        uint64_t total = 0;
        for (auto& c : chunks_) total = c.res0 + c.n_res;
        parallel_chunks((int)chunks_.size(), [&](int i) { fill_chunk(chunks_[i], res, offsets, names, name_off, packed, total); });
        uint64_t n_rec = chunks_.empty() ? 0 : chunks_.back().rec0 + chunks_.back().n_rec;
        offsets[n_rec] = total;
        for (auto& c : chunks_)
            if (c.bad) { *bad = c.bad_res; return false; }  // chunks are in file order: the first bad chunk holds the first error
        if (packed) {
            const uint64_t groups = (total + 7) / 8;
            bool ok = true;
            for (auto& c : chunks_) {
                ok = ok && c.pack_ok;
                const uint64_t ga = c.res0 / 8, gb = (c.res0 + c.n_res) / 8;  // the groups that hold this chunk's two ends
                if (ga < groups) ok = pack_groups(res, total, packed, ga, ga + 1) && ok;
                if (gb < groups && gb != ga) ok = pack_groups(res, total, packed, gb, gb + 1) && ok;
            }
            memset(packed + groups * 5, 0, 72);
            if (packed_ok) *packed_ok = ok;
        }
        return true;
    }

  private:
    size_t record_start_at_or_after(size_t p) const {
        if (p >= n_) return n_;
        if (p == 0 || d_[p - 1] == '\n') { if (d_[p] == '>') return p; }
        while (p < n_) {
            const char* q = (const char*)memchr(d_ + p, '>', n_ - p);
            if (!q) return n_;
            const size_t at = (size_t)(q - d_);
            if (at == 0 || d_[at - 1] == '\n') return at;
            p = at + 1;
        }
        return n_;
    }

    void count_chunk(FastaChunk& c) const {
        bool stopped = false;
        for_each_line(d_, c.begin, c.end,
            [&](size_t b, size_t e) { c.n_rec++; c.name_bytes += e - b + 1; stopped = false; },
            [&](size_t b, size_t e) {
                if (stopped || e <= b) return;
                if (mode_ == NORM_KMERSEEK) {
                    const char* star = (const char*)memchr(d_ + b, '*', e - b);
                    if (star) { c.n_res += (size_t)(star - (d_ + b)) + 1; stopped = true; return; }
                }
                c.n_res += e - b;
            });
    }

    void fill_chunk(FastaChunk& c, uint8_t* res, uint64_t* offsets, char* names, uint64_t* name_off, uint8_t* packed,
                    uint64_t total) const {
        static const ResidueClass cls;
        uint64_t rec = c.rec0, out = c.res0, nm = c.name0;
        uint64_t rec_start = out;
        bool stopped = false;
        uint64_t pk = (c.res0 + 7) / 8;  // next group of this chunk's own range to pack (groups whose 8 residues are all ours)
        auto pack_upto = [&](uint64_t g1) {
            if (packed && g1 > pk) { c.pack_ok = pack_groups(res, total, packed, pk, g1) && c.pack_ok; pk = g1; }
        };
        for_each_line(d_, c.begin, c.end,
            [&](size_t b, size_t e) {
                offsets[rec] = out;
                rec_start = out;
                name_off[rec] = nm;
                memcpy(names + nm, d_ + b, e - b);
                names[nm + (e - b)] = 0;
                nm += e - b + 1;
                rec++;
                stopped = false;
            },
            [&](size_t b, size_t e) {
                if (stopped || e <= b || c.bad) return;
                const uint8_t* in = (const uint8_t*)d_ + b;
                const size_t len = e - b;
                uint8_t* o = res + out;
                if (out / 8 >= pk + 512) pack_upto(out / 8);  // the last 4 KB of residues, still in this core's cache
                if (mode_ == NORM_SOURMASH) {
                    for (size_t i = 0; i < len; i++) o[i] = cls.up[in[i]];
                    out += len;
                    return;
                }
                // nothing is written past what pass 1 counted for this line (the next bytes belong to another record,
                // maybe another thread's): the line ends for us right after its first '*'
                const uint8_t* star = (const uint8_t*)memchr(in, '*', len);
                const size_t lim = star ? (size_t)(star - in) + 1 : len;
                uint32_t any = 0;  // fast path: a line of plain valid residues
                for (size_t i = 0; i < lim; i++) { const uint32_t v = cls.both[in[i]]; o[i] = (uint8_t)v; any |= v; }
                if (!(any >> 8)) { out += lim; return; }
                for (size_t i = 0; i < lim; i++) {
                    const uint8_t s = cls.special[in[i]];
                    if (s == 0) continue;
                    if (s == 1) { out += i + 1; stopped = true; return; }  // '*' is kept, the rest of the record dropped
                    if (s == 2) { o[i] = resolve_ambiguous(o[i], seed_, out + i - rec_start); continue; }
                    c.bad = true;  // src/rust/aminoacid.rs:85-87: the character and its 1-based position
                    c.bad_res.ch = o[i]; c.bad_res.pos = out + i - rec_start + 1; c.bad_res.protein = rec - 1;
                    return;
                }
                out += lim;
            });
        if (!c.bad) pack_upto((c.res0 + c.n_res) / 8);
    }

    const char* d_;
    size_t n_;
    NormalizeMode mode_;
    uint64_t seed_;
    size_t start_ = 0;
    std::vector<FastaChunk> chunks_;
};

inline bool pack_residues_parallel(const uint8_t* res, uint64_t n, uint8_t* out) {
    const uint64_t groups = (n + 7) / 8;
    const int t = ingest_threads((size_t)n);
    std::vector<char> ok(t, 1);
    parallel_chunks(t, [&](int i) {
        const uint64_t g0 = groups * (uint64_t)i / t, g1 = groups * (uint64_t)(i + 1) / t;
        ok[i] = pack_groups(res, n, out, g0, g1) ? 1 : 0;
    });
    memset(out + groups * 5, 0, 72);
    for (char c : ok) if (!c) return false;
    return true;
}

}  // namespace ks
