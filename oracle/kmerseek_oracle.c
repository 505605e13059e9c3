/*
 * kmerseek_oracle.c -- CPU restatement of kmerseek's sketch-and-index hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kmerseek_b200/ may link, import or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and there only as the checker or as the timed CPU baseline.
 *
 * Parity status: PINNED.  The arithmetic lives in third-party crates that are not on disk
 * (sourmash 0.20.0 -> murmurhash3 0.0.5; /root/reference/Cargo.lock:3339-3342,2018-2021), so it is
 * restated here from the published algorithm (MurmurHash3_x64_128, Appleby, public domain;
 * sourmash `encodings.rs` Dayhoff/HP tables) and pinned against the reference's own golden vectors
 * (tests/golden/, extracted by tests/golden/make_golden.py): 48 known-answer hashes, six ids,
 * combined sizes 9049/2730/3549/1603, 75 full sketches and 5176 k-mer rows.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KSO_PROTEIN 0
#define KSO_DAYHOFF 1
#define KSO_HP 2

/* ------------------------------------------------------------------------------------------ */
/* MurmurHash3_x64_128, low 64 bits.  sourmash::_hash_murmur (called at src/rust/index.rs:766) */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

uint64_t kso_murmur64(const uint8_t *data, uint64_t len, uint64_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t nblocks = len / 16;
    for (uint64_t i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8); /* little-endian host assumed (x86-64) */
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    uint64_t rem = len & 15;
    for (uint64_t i = rem; i > 8; i--) k2 |= (uint64_t)tail[i - 1] << (8 * (i - 9));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (uint64_t i = (rem > 8 ? 8 : rem); i > 0; i--) k1 |= (uint64_t)tail[i - 1] << (8 * (i - 1));
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* ------------------------------------------------------------------------------------------ */
/* max_hash.  sourmash max_hash_for_scaled, reached through KmerMinHash::new at                */
/* src/rust/signature.rs:124-131.  Golden: 3689348814741910528 for scaled=5 (sig.zip JSON).     */
/* ------------------------------------------------------------------------------------------ */
uint64_t kso_max_hash(uint32_t scaled) {
    if (scaled == 0) return 0;
    if (scaled == 1) return UINT64_MAX;
    double v = 18446744073709551616.0 / (double)scaled;
    return (uint64_t)v;
}

/* ------------------------------------------------------------------------------------------ */
/* Alphabet reduction.  src/rust/encoding.rs:43-53 -> sourmash aa_to_dayhoff / aa_to_hp.        */
/* Goldens: LIVINGALIVE -> eeeecbbeeec / hhhhphhhhhp (src/rust/encoding.rs:195,209).            */
/* ------------------------------------------------------------------------------------------ */
uint8_t kso_translate(uint8_t aa, int moltype) {
    if (moltype == KSO_PROTEIN) return aa;
    if (aa == '*') return '*';
    if (moltype == KSO_DAYHOFF) {
        switch (aa) {
        case 'C': return 'a';
        case 'A': case 'G': case 'P': case 'S': case 'T': return 'b';
        case 'D': case 'E': case 'N': case 'Q': return 'c';
        case 'H': case 'K': case 'R': return 'd';
        case 'I': case 'L': case 'M': case 'V': return 'e';
        case 'F': case 'W': case 'Y': return 'f';
        default: return 'X';
        }
    }
    switch (aa) {
    case 'A': case 'F': case 'G': case 'I': case 'L': case 'M': case 'P': case 'V': case 'W': case 'Y':
        return 'h';
    case 'C': case 'D': case 'E': case 'H': case 'K': case 'N': case 'Q': case 'R': case 'S': case 'T':
        return 'p';
    default: return 'X';
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Input normalisation.  to_uppercase at src/rust/index.rs:1000, then                          */
/* AminoAcidAmbiguity::validate_and_resolve, src/rust/aminoacid.rs:74-105: keep and stop at the */
/* first '*'; reject anything outside 20 standard + XUO* + BZJ with (char, 1-based position in  */
/* the output so far); B/Z/J are resolved -- at random in the reference (aminoacid.rs:45-54),   */
/* here by the deterministic rule the boundary documents (include/kmerseek_b200.h):             */
/* choice = splitmix64(ambig_seed ^ (protein_index << 32) ^ position) & 1.                      */
/* Returns the output length, or -1 with *bad_char / *bad_pos set.  ASCII only.                 */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

int64_t kso_normalize(const uint8_t *seq, uint64_t len, uint64_t protein_index, uint64_t ambig_seed,
                      uint8_t *out, uint8_t *bad_char, uint64_t *bad_pos) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < len; i++) {
        uint8_t c = seq[i];
        if (c >= 'a' && c <= 'z') c -= 32;
        if (c == '*') { out[n++] = c; break; }
        int ok = 0;
        switch (c) {
        case 'A': case 'C': case 'D': case 'E': case 'F': case 'G': case 'H': case 'I': case 'K': case 'L':
        case 'M': case 'N': case 'P': case 'Q': case 'R': case 'S': case 'T': case 'V': case 'W': case 'Y':
        case 'X': case 'U': case 'O':
            ok = 1; break;
        case 'B': case 'Z': case 'J': {
            /* the reference draws at random (aminoacid.rs:45-54); the build fixes the draw by the seed and the residue's
             * position in its own sequence, independent of the record's index (protein_index is kept for the signature) */
            (void)protein_index;
            uint64_t r = splitmix64(ambig_seed ^ n) & 1;
            c = c == 'B' ? (r ? 'N' : 'D') : c == 'Z' ? (r ? 'Q' : 'E') : (r ? 'L' : 'I');
            ok = 1; break;
        }
        default: break;
        }
        if (!ok) { *bad_char = seq[i] >= 'a' && seq[i] <= 'z' ? seq[i] - 32 : seq[i]; *bad_pos = n + 1; return -1; }
        out[n++] = c;
    }
    return (int64_t)n;
}

/* ------------------------------------------------------------------------------------------ */
/* Sketch tuples.  For every window of every protein (src/rust/index.rs:758: i in               */
/* 0..len-(k-1), none when len<k): translate, hash (seed 42, src/rust/signature.rs:12), skip    */
/* h==0 and keep iff h<=max_hash (sourmash KmerMinHash::add_protein via                        */
/* src/rust/signature.rs:273-274).  Emits every kept occurrence (h, pid, pos) in (pid,pos)      */
/* order -- the union of what add_protein keeps (distinct h) and process_kmers records          */
/* (positions, src/rust/index.rs:769-780).  Returns the count; writes at most cap entries.      */
/* ------------------------------------------------------------------------------------------ */
uint64_t kso_sketch_tuples(const uint8_t *residues, const uint64_t *offsets, uint64_t n_prot, uint32_t k,
                           int moltype, uint32_t scaled, uint64_t *out_hash, uint32_t *out_pid,
                           uint32_t *out_pos, uint64_t cap) {
    const uint64_t max_hash = kso_max_hash(scaled);
    uint8_t lut[256];
    for (int i = 0; i < 256; i++) lut[i] = kso_translate((uint8_t)i, moltype);
    uint8_t *buf = (uint8_t *)malloc(k ? k : 1);
    uint64_t n = 0;
    for (uint64_t p = 0; p < n_prot; p++) {
        const uint8_t *s = residues + offsets[p];
        uint64_t len = offsets[p + 1] - offsets[p];
        if (k == 0 || len < k) continue;
        for (uint64_t i = 0; i + k <= len; i++) {
            for (uint32_t j = 0; j < k; j++) buf[j] = lut[s[i + j]];
            uint64_t h = kso_murmur64(buf, k, 42);
            if (h == 0 || h > max_hash) continue;
            if (n < cap) { out_hash[n] = h; out_pid[n] = (uint32_t)p; out_pos[n] = (uint32_t)i; }
            n++;
        }
    }
    free(buf);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Faithful-cost CPU baseline: the reference's own cost structure, for timing only.             */
/*   per protein  (create_protein_signature, src/rust/index.rs:719-747)                        */
/*     pass 1: add_protein -- per window a freshly allocated translated k-mer, murmur,          */
/*             binary search + memmove insert into sorted mins/abunds (sourmash KmerMinHash)    */
/*     pass 2: process_kmers (src/rust/index.rs:749-786) -- re-translate, re-hash, LINEAR       */
/*             `hashvals.contains` scan (:769), then record the position under the hash         */
/*   per batch of 1000 (process_batch_parallel, :984-1016): proteins in parallel (rayon there,  */
/*             pthreads here), then store_signatures (:800-830) under one lock: every           */
/*             (min, abund) is inserted into ONE sorted combined vector (O(U) memmove each).    */
/* `faithful=0` replaces the linear scan by the binary search and skips the combined insert     */
/* (a "fast" CPU variant for the large configs; the reference itself cannot finish them, F8).   */
/* Returns residues processed; *out_unique = combined sketch size (faithful) or 0.              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t *mins; uint64_t *abunds; uint64_t n, cap; } kso_vec;

static void vec_insert(kso_vec *v, uint64_t h, uint64_t abund) {
    uint64_t lo = 0, hi = v->n;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (v->mins[mid] < h) lo = mid + 1; else hi = mid; }
    if (lo < v->n && v->mins[lo] == h) { v->abunds[lo] += abund; return; }
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->mins = (uint64_t *)realloc(v->mins, v->cap * 8);
        v->abunds = (uint64_t *)realloc(v->abunds, v->cap * 8);
    }
    memmove(v->mins + lo + 1, v->mins + lo, (v->n - lo) * 8);
    memmove(v->abunds + lo + 1, v->abunds + lo, (v->n - lo) * 8);
    v->mins[lo] = h; v->abunds[lo] = abund; v->n++;
}

typedef struct {
    const uint8_t *residues; const uint64_t *offsets; uint32_t k; int moltype; uint64_t max_hash;
    int faithful; uint64_t p_begin, p_end; volatile uint64_t *next; kso_vec *sigs; uint64_t *n_positions;
    const uint8_t *lut;
} kso_job;

static void sketch_one(const kso_job *J, uint64_t p, kso_vec *sig, uint64_t *n_pos) {
    const uint8_t *s = J->residues + J->offsets[p];
    uint64_t len = J->offsets[p + 1] - J->offsets[p];
    uint32_t k = J->k;
    sig->n = 0;
    if (len < k) return;
    for (uint64_t i = 0; i + k <= len; i++) { /* pass 1: add_protein */
        uint8_t *kmer = (uint8_t *)malloc(k); /* the reference allocates per window */
        for (uint32_t j = 0; j < k; j++) kmer[j] = J->lut[s[i + j]];
        uint64_t h = kso_murmur64(kmer, k, 42);
        free(kmer);
        if (h == 0 || h > J->max_hash) continue;
        vec_insert(sig, h, 1);
    }
    uint64_t positions = 0; /* pass 2: process_kmers */
    for (uint64_t i = 0; i + k <= len; i++) {
        uint8_t *enc = (uint8_t *)malloc(k), *orig = (uint8_t *)malloc(k);
        for (uint32_t j = 0; j < k; j++) { enc[j] = J->lut[s[i + j]]; orig[j] = s[i + j]; }
        uint64_t h = kso_murmur64(enc, k, 42);
        int found = 0;
        if (J->faithful) {
            for (uint64_t q = 0; q < sig->n; q++) if (sig->mins[q] == h) { found = 1; break; }
        } else {
            uint64_t lo = 0, hi = sig->n;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (sig->mins[mid] < h) lo = mid + 1; else hi = mid; }
            found = lo < sig->n && sig->mins[lo] == h;
        }
        positions += found;
        free(enc); free(orig);
    }
    *n_pos += positions;
}

static void *batch_worker(void *arg) {
    kso_job *J = (kso_job *)arg;
    for (;;) {
        uint64_t p = __sync_fetch_and_add(J->next, 1);
        if (p >= J->p_end) break;
        sketch_one(J, p, &J->sigs[p - J->p_begin], J->n_positions);
    }
    return NULL;
}

uint64_t kso_cpu_baseline(const uint8_t *residues, const uint64_t *offsets, uint64_t n_prot, uint32_t k,
                          int moltype, uint32_t scaled, int faithful, int n_threads, uint64_t batch_size,
                          uint64_t *out_unique, uint64_t *out_kept) {
    uint8_t lut[256];
    for (int i = 0; i < 256; i++) lut[i] = kso_translate((uint8_t)i, moltype);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (batch_size == 0) batch_size = 1000;
    kso_vec combined = {0, 0, 0, 0};
    kso_vec *sigs = (kso_vec *)calloc(batch_size, sizeof(kso_vec));
    uint64_t kept = 0;
    uint64_t *npos = (uint64_t *)calloc(n_threads, sizeof(uint64_t) * 8);
    for (uint64_t b = 0; b < n_prot; b += batch_size) {
        uint64_t e = b + batch_size < n_prot ? b + batch_size : n_prot;
        volatile uint64_t next = b;
        pthread_t th[256];
        kso_job jobs[256];
        for (int t = 0; t < n_threads; t++) {
            kso_job j = {residues, offsets, k, moltype, kso_max_hash(scaled), faithful, b, e, &next, sigs,
                         npos + 8 * t, lut};
            jobs[t] = j;
            if (n_threads > 1) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
        }
        if (n_threads == 1) batch_worker(&jobs[0]);
        else for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        for (uint64_t p = b; p < e; p++) { /* store_signatures: serial, under the reference's one lock */
            kso_vec *s = &sigs[p - b];
            if (faithful) for (uint64_t q = 0; q < s->n; q++) vec_insert(&combined, s->mins[q], s->abunds[q]);
        }
    }
    for (int t = 0; t < n_threads; t++) kept += npos[8 * t];
    for (uint64_t i = 0; i < batch_size; i++) { free(sigs[i].mins); free(sigs[i].abunds); }
    free(sigs); free(npos);
    if (out_unique) *out_unique = combined.n;
    if (out_kept) *out_kept = kept;
    free(combined.mins); free(combined.abunds);
    return offsets[n_prot] - offsets[0];
}

/* ------------------------------------------------------------------------------------------ */
/* All-pairs search baseline: branchwater manysearch's inner loop (called at                    */
/* src/python/kmerseek/search.py:125-141): for every (query, target) pair of sorted distinct    */
/* sketches count the intersection by a sorted-list merge.  Timing baseline + small-case check. */
/* q_ptr/t_ptr are CSR offsets into the concatenated sorted mins.  out_counts is nq*nt.          */
/* ------------------------------------------------------------------------------------------ */
void kso_allpairs_intersect(const uint64_t *q_mins, const uint64_t *q_ptr, uint64_t nq, const uint64_t *t_mins,
                            const uint64_t *t_ptr, uint64_t nt, uint32_t *out_counts) {
    for (uint64_t t = 0; t < nt; t++) {
        for (uint64_t q = 0; q < nq; q++) {
            uint64_t i = q_ptr[q], ie = q_ptr[q + 1], j = t_ptr[t], je = t_ptr[t + 1];
            uint32_t c = 0;
            while (i < ie && j < je) {
                if (q_mins[i] < t_mins[j]) i++;
                else if (q_mins[i] > t_mins[j]) j++;
                else { c++; i++; j++; }
            }
            out_counts[q * nt + t] = c;
        }
    }
}
