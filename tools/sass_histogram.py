"""Opcode histogram of the shipped kernels, read from the built library with cuobjdump (no GPU needed):

    python tools/sass_histogram.py [regex] > profiles/r02_sass_<name>.txt

For every kernel whose demangled name matches `regex` (default: k_tile.*ElasticityQuad4Op|k_hex8_chunk_rows): the
instruction count per SASS opcode and the counters that show which hardware paths the kernel uses -- UBLKCP (TMA
1-D bulk copies, cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier), DFMA/DMUL/DADD (FP64 CUDA cores), and the
absence of UTMALDG/UTMASTG (tensor-map TMA) and UTC*MMA/LDTM/STTM (tcgen05 / TMEM): the element products are 4x4 ..
24x24 per element with data-dependent gather / scatter, not a tileable dense contraction (BASELINE.json north_star).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyfem_gpu_testflight_b200", "libpyfem_b200.so")


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_tile<.*ElasticityQuad4Op|k_hex8_chunk_rows")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    arch = re.search(r"arch = (sm_\w+)", sass)
    print(f"library: {os.path.relpath(LIB, ROOT)}   arch: {arch.group(1) if arch else '?'}")
    for name, hist in kernels.items():
        if not pat.search(name):
            continue
        total = sum(hist.values())
        print(f"\n== {name}\n   {total} SASS instructions")
        keys = ("UBLKCP", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR",
                "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "DMMA")
        print("   " + "  ".join(f"{k}={hist.get(k, 0)}" for k in keys))
        for op, n in hist.most_common(24):
            print(f"   {op:12s} {n:6d}  {100.0 * n / total:5.1f}%")


if __name__ == "__main__":
    main()
