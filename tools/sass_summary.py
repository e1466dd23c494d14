"""Which kernels of libdram_b200.so use the Blackwell tensor cores and TMA: counts of the SASS mnemonics that
B200_PROFILING.md lists (tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS) per kernel, from `cuobjdump -sass` (no GPU needed).
    python tools/sass_summary.py > profiles/sass_mnemonics_TAG.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bodyct-dram-emph-subtype_b200", "libdram_b200.so")
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS"]


_CACHE = {}


def kernel_sass(lib=LIB):
    """{demangled kernel name (no argument list): SASS text} for every kernel in the library."""
    if lib not in _CACHE:
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
        blocks = re.split(r"\n\s*Function : ", sass)[1:]
        names = [b.split("\n", 1)[0].strip() for b in blocks]
        dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.strip().splitlines()
        _CACHE[lib] = {name.split("(")[0].replace("void ", ""): body for name, body in zip(dem, blocks)}
    return _CACHE[lib]


def kernels(lib=LIB):
    """{kernel name: {mnemonic: count}}."""
    return {name: {k: len(re.findall(r"\b" + k, body)) for k in KEYS} for name, body in kernel_sass(lib).items()}


def main():
    table = kernels()
    print("# SASS mnemonics per kernel of libdram_b200.so (`cuobjdump -sass`, sm_100a)\n")
    print("tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk.tensor load / store = UTMALDG / UTMASTG, "
          "tcgen05.commit = UTCBAR, mbarrier = SYNCS (B200_PROFILING.md).\n")
    print("| kernel | " + " | ".join(KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for name in sorted(table):
        c = table[name]
        if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"]:
            print(f"| {name} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
    rest = sorted(n for n, c in table.items() if not (c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"]))
    print(f"\n{len(rest)} memory-bound / glue kernels without tensor-core or TMA instructions: " + ", ".join(rest))


if __name__ == "__main__":
    sys.exit(main())
