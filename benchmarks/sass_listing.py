#!/usr/bin/env python
"""SASS listing of selected kernels of one object file, with a mnemonic histogram and the Blackwell / async-proxy
mnemonics that prove TMA bulk copies (UBLKCP), mbarriers (SYNCS), tcgen05 (UTCHMMA / LDTM) and packed fp32 (FFMA2).

    python benchmarks/sass_listing.py activezero_b200/build/gwc_volume.o gwc_bwd_systolic gwc_fwd_planes > profiles/r2_sass_gwc.txt
"""
import collections
import re
import subprocess
import sys

SPECIAL = re.compile(r"^(UBLKCP|UTMA|SYNCS|UTC|LDTM|STTM|FFMA2|FADD2|FMUL2|REDUX|UTCBAR)")


def main():
    obj, pats = sys.argv[1], sys.argv[2:]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    print(f"# cuobjdump -sass {obj}  (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo); "
          f"regenerate with: python benchmarks/sass_listing.py {obj} {' '.join(pats)}")
    blocks = re.split(r"\n\s*Function : ", txt)[1:]
    for blk in blocks:
        name = blk.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        if pats and not any(p in dem for p in pats):
            continue
        lines = [ln for ln in blk.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln)]
        hist, special = collections.Counter(), collections.Counter()
        for ln in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                hist[m.group(1).split(".")[0]] += 1
                if SPECIAL.match(m.group(1)):
                    special[m.group(1)] += 1
        print(f"\n## {dem}")
        print(f"# {len(lines)} instructions; mnemonic histogram: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
        if special:
            print("# Blackwell / async-proxy mnemonics present: " + ", ".join(f"{k} x{v}" for k, v in special.items()))
        print("\n".join(lines))


if __name__ == "__main__":
    main()
