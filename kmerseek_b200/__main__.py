"""`python -m kmerseek_b200 index|search` -- the thin callers of the hot path.

  index   mirrors `kmerseek-rust index` (src/rust/main.rs:16-134): same options, same progress lines.
          The index lives in HBM; `--output` names it and receives the sourmash-compatible `.sig.zip`
          (persistence in the reference's RocksDB layout is row N2, not built).
  search  mirrors `kmerseek search QUERY TARGET` (src/python/kmerseek/search.py:287-384): CSV on stdout
          (22 manysearch columns, or the stitched regions with --extract-kmers, visual block on stderr).
"""
import argparse
import os
import sys

import kmerseek_b200 as K
from kmerseek_b200 import export


def cmd_index(a):
    print(f"Indexing FASTA file: {a.input}")
    if not os.path.exists(a.input):
        print(f"Error: No such file or directory (os error 2): {a.input}", file=sys.stderr)
        return 1
    if a.output:
        out = a.output
        print(f"Output database: {out}")
    else:
        base = os.path.basename(a.input)
        out = os.path.join(os.path.dirname(a.input) or ".", f"{base}.{a.encoding}.k{a.ksize}.scaled{a.scaled}.kmerseek.rocksdb")
        print(f"Auto-generated output database: {out}")
    print(f"\n-------\nK-mer size: {a.ksize}")
    print(f"Scaled: {a.scaled}")
    print(f"Encoding: {a.encoding.capitalize()}")
    print(f"Progress interval: {a.progress_interval}")
    print(f"Store raw sequences: {str(a.store_raw_sequences).lower()}")
    print("-------\n")
    idx = K.ProteomeIndex(out, a.ksize, a.scaled, a.encoding, a.store_raw_sequences, device=a.device)
    print("Processing FASTA file...")
    idx.process_fasta(a.input, a.progress_interval, 1000)
    os.makedirs(out, exist_ok=True)
    sig = os.path.join(out, os.path.basename(export.sig_filename(a.input, a.encoding, a.ksize, a.scaled)))
    export.write_sig_zip(sig, idx.export_sketches(), idx.names(), a.ksize, a.scaled, a.encoding, a.input, K.max_hash(a.scaled))
    print("Indexing completed successfully!")
    print(f"Database saved to: {out}")
    idx.close()
    return 0


def cmd_search(a):
    idx = K.ProteomeIndex(a.target_fasta, a.ksize, a.scaled, a.moltype, device=a.device)
    # the reference's search sketches both files through branchwater manysketch (src/python/kmerseek/sketch.py:28-40):
    # records reach sourmash add_protein upper-cased and otherwise untouched -- no validation, no '*' truncation, B/Z/J
    # kept (they translate to 'X' under dayhoff / hp); the index path's normalisation is not applied here
    t = K.Proteome.from_fasta(a.target_fasta, mode="sourmash")
    idx.add_proteome(t)
    q = K.Proteome.from_fasta(a.query_fasta, mode="sourmash")
    res = K.search(idx, q, hits=a.extract_kmers)
    rows = K.manysearch_rows(res, idx, q.names, targets=t)
    if a.sourmash_search_csv:
        export.write_manysearch_csv(a.sourmash_search_csv, rows)
    if not a.extract_kmers:
        text = export.manysearch_csv(rows)
    else:
        qs = [q.sequence(i) for i in range(q.n_proteins)]
        ts = [t.sequence(i) for i in range(t.n_proteins)]
        st = K.stitch_hits(res, idx, qs, ts, q.names, t.names)
        cols = ["match_name", "query_name", "query_start", "query_end", "query", "match_start", "match_end", "match",
                "encoded", "length"]
        import csv
        import io
        buf = io.StringIO()
        w = csv.writer(buf, lineterminator="\n")
        w.writerow(cols)
        for r in st:
            w.writerow([r[c] for c in cols])
            print(f"\n---\nQuery Name: {r['query_name']}\nMatch Name: {r['match_name']}\n"
                  f"query: {r['query']} ({r['query_start']}-{r['query_end']})\nalpha: {r['encoded']}\n"
                  f"match: {r['match']} ({r['match_start']}-{r['match_end']})", file=sys.stderr)
        text = buf.getvalue()
    if a.output:
        open(a.output, "w").write(text)
    else:
        sys.stdout.write(text)
    idx.close()
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="kmerseek", description="Efficient protein domain annotation search with reduced amino acid k-mers")
    sub = ap.add_subparsers(dest="command", required=True)
    i = sub.add_parser("index", help="Index a FASTA file")
    i.add_argument("-i", "--input", required=True)
    i.add_argument("-o", "--output", default=None)
    i.add_argument("-k", "--ksize", type=int, default=10)
    i.add_argument("-s", "--scaled", type=int, default=1)
    i.add_argument("-e", "--encoding", default="protein", choices=["protein", "dayhoff", "hp"])
    i.add_argument("-p", "--progress-interval", type=int, default=10000)
    i.add_argument("--store-raw-sequences", action="store_true")
    i.add_argument("--device", type=int, default=0)
    i.set_defaults(fn=cmd_index)
    s = sub.add_parser("search", help="Search for k-mers in target sequences")
    s.add_argument("query_fasta")
    s.add_argument("target_fasta")
    s.add_argument("--moltype", default="hp")
    s.add_argument("--ksize", type=int, default=24)
    s.add_argument("--scaled", type=int, default=5)
    s.add_argument("--extract-kmers", action="store_true")
    s.add_argument("--output", default=None)
    s.add_argument("--sourmash-search-csv", default=None)
    s.add_argument("--device", type=int, default=0)
    s.set_defaults(fn=cmd_search)
    a = ap.parse_args(argv)
    try:
        return a.fn(a)
    except K.IndexError_ as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
