// Exercises include/kmerseek_b200.hpp the way the reference's Rust tests use ProteomeIndex
// (src/rust/index.rs:1395-1441, 3021-3036).  `facade_demo builder` needs no GPU; `facade_demo gpu` does.
#include <cstdio>
#include <cstring>
#include <string>

#include "kmerseek_b200.hpp"

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "builder";
    try {
        kmerseek::ProteomeIndex::builder().path("x").scaled(1).moltype("hp").build();
        std::printf("FAIL: builder accepted a missing ksize\n");
        return 1;
    } catch (const kmerseek::IndexError& e) {
        if (std::strcmp(e.what(), "Builder error: K-mer size is required") != 0) { std::printf("FAIL: %s\n", e.what()); return 1; }
    }
    try {
        kmerseek::ProteomeIndex("x", 5, 1, "dna", false);
        std::printf("FAIL: moltype dna accepted\n");
        return 1;
    } catch (const kmerseek::IndexError& e) {
        if (e.status != KS_ERR_INVALID_MOLTYPE) { std::printf("FAIL: %s\n", e.what()); return 1; }
    }
    if (mode == "builder") { std::printf("builder ok\n"); return 0; }
    for (const char* moltype : {"protein", "dayhoff", "hp"}) {
        auto index = kmerseek::ProteomeIndex::builder().path("test.db").ksize(5).scaled(1).moltype(moltype).build();
        auto sig = index.create_protein_signature("PLANTANDANIMALGENQMES", "test_protein");
        const size_t n_infos = sig.kmer_infos().size();
        index.store_signatures({sig});
        std::printf("%s %zu %zu %s\n", moltype, n_infos, index.combined_minhash_size(), sig.md5sum.c_str());
        try {
            index.create_protein_signature("PLANTANDANIMALGEN1MES", "bad");
            return 1;
        } catch (const kmerseek::IndexError& e) {
            if (!std::strstr(e.what(), "Invalid amino acid '1'")) { std::printf("FAIL: %s\n", e.what()); return 1; }
        }
    }
    return 0;
}
