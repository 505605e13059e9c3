#!/usr/bin/env python3
"""Per-kernel SASS summary of the built library: resources (cuobjdump -res-usage) and the static instruction mix
(cuobjdump -sass), for the kernels on the hot path.

    python profiles/sass_summary.py kmerseek_b200/libkmerseek_b200.so > profiles/r02_s_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

HOT = ["sketch_quad_kernel<16, true, true, true>", "sketch_quad_kernel<7, false, false, false>", "sketch_dense_kernel<24>",
       "pair_partition_kernel", "dense_partition_kernel", "bucket_sort_bin_kernel<false, true>", "bucket_sort_rep_kernel",
       "dense_bucket_kernel<false>", "query_kernel<512, 1024, 128>", "query_kernel<4096, 4096, 512>", "query_scan_kernel",
       "finalize_pairs_kernel", "expand_hits_kernel", "merge_pairs_kernel", "dense_chunks_kernel", "dense_bucket_offsets_kernel",
       "tile_pid_kernel"]
GROUPS = [("LDG/LD global loads", r"^(LDG|LD)\b"), ("STG/ST global stores", r"^(STG|ST)\b"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("ATOMS/ATOMG/RED", r"^(ATOMS|ATOMG|ATOM|RED)"), ("BAR", r"^BAR"), ("SHFL/VOTE/MATCH", r"^(SHFL|VOTE|MATCH)"),
          ("IMAD/IMAD.WIDE", r"^IMAD"), ("LOP3/SHF/IADD3/LEA", r"^(LOP3|SHF|IADD3|LEA|IADD)"), ("ISETP/PLOP3/SEL", r"^(ISETP|PLOP3|SEL)"),
          ("POPC/FLO/BREV", r"^(POPC|FLO|BREV)"), ("BRA/BSSY/BSYNC/EXIT", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET)"),
          ("LDC/ULDC/S2R", r"^(LDC|ULDC|S2R|S2UR|CS2R)"), ("DADD/DMUL/DFMA/MUFU", r"^(DADD|DMUL|DFMA|MUFU|DSETP)"),
          ("UTMALDG/UTMASTG/UBLKCP (TMA)", r"^(UTMALDG|UTMASTG|UBLKCP|UTMAPF)"), ("UTCMMA/tcgen05", r"^(UTC|TCGEN)"),
          ("LDGSTS/LDGDEPBAR (cp.async)", r"^(LDGSTS|LDGDEPBAR)"), ("SYNCS (mbarrier)", r"^SYNCS")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(lib):
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for l in res.splitlines():
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", l)
        if m and cur:
            st = re.search(r"STACK:(\d+)", l)
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(st.group(1)) if st else 0)
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    mix, cur = {}, None
    for l in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and cur:
            mix[cur][m.group(1)] += 1
    dm = demangle(list(mix))
    short = {k: re.sub(r"\(.*", "", re.sub(r"ks::\(anonymous namespace\)::|ks::|void ", "", v)).replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "") for k, v in dm.items()}
    print("# static SASS summary of the hot-path kernels (sm_100a; profiles/sass_summary.py; the library of the r02_s / r02_w runs)")
    print("# No kernel uses TMA (UTMALDG / UBLKCP), mbarriers or tcgen05: the path is integer / byte work on tiles of a few KB that")
    print("# are read once with plain coalesced LDG.128 / LDG.64 and written once; the tensor cores have nothing to do here.\n")
    for want in HOT:
        hits = [k for k, v in short.items() if v == want]
        if not hits:
            print(f"## {want}: not found\n")
            continue
        k = hits[0]
        c = mix[k]
        n = sum(c.values())
        reg, sh, st = usage.get(k, (0, 0, 0))
        print(f"## {want}: {n} instructions, {reg} registers, {sh} B static shared memory, {st} B stack")
        row = []
        for name, rx in GROUPS:
            cnt = sum(v for op, v in c.items() if re.match(rx, op))
            if cnt or "TMA" in name or "tcgen05" in name:
                row.append(f"{name} {cnt}")
        print("   " + "; ".join(row))
        wide = sum(v for op, v in c.items() if re.match(r"^(LDG|STG|LDS|STS)", op) and ".128" in op)
        w64 = sum(v for op, v in c.items() if re.match(r"^(LDG|STG|LDS|STS)", op) and ".64" in op)
        print(f"   memory instructions of 128 bits: {wide}, of 64 bits: {w64}\n")


if __name__ == "__main__":
    main(sys.argv[1])
