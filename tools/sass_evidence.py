#!/usr/bin/env python
"""Count, per kernel of the built objects, the SASS mnemonics that show which Blackwell machinery the code uses
(B200_PROFILING.md): UTCQMMA / UTCOMMA = tcgen05.mma (kind::f8f6f4 / block-scaled kind::mxf4), UTMALDG = TMA tensor
load, UBLKCP = 1-D bulk copy, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
ACQBULK / UCGABAR = cluster / grid-dependency control, HMMA / QMMA = mma.sync.  No GPU needed.
usage: sass_evidence.py [objects...]   (default: mila_b200/csrc/*.o)"""
import collections
import glob
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
MNEMONICS = ("UTCQMMA", "UTCOMMA", "UTCHMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP",
             "SYNCS", "ACQBULK", "UCGABAR", "HMMA", "QMMA", "LDGSTS", "REDUX")


def demangle(name: str) -> str:
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return name
    out = re.sub(r"milab200::\(anonymous namespace\)::|milab200::", "", out)
    return re.sub(r"\(.*", "", out)


def main() -> None:
    objs = sys.argv[1:] or sorted(glob.glob(str(ROOT / "mila_b200" / "csrc" / "*.o")))
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        counts: dict = collections.OrderedDict()
        fn = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = demangle(m.group(1)); counts.setdefault(fn, collections.Counter()); continue
            if fn is None:
                continue
            m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and m.group(1) in MNEMONICS:
                counts[fn][m.group(1)] += 1
        rows = [(fn, c) for fn, c in counts.items() if c]
        if rows:
            print(f"== {Path(obj).name}")
            for fn, c in rows:
                print(f"  {fn}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
