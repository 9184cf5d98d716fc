"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the evidence that the hot kernels
are tcgen05 / TMA code (UTCHMMA, UTMALDG, LDTM, STTM, UBLKCP, UTCBAR ...) and which ones still use the legacy
warp-level tensor path (HMMA).  Writes profiles/r02_sass_opcodes.json and prints a table.

  python scripts/sass_histogram.py [path/to/libflashattention.so]
"""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "HMMA", "LDGSTS", "SYNCS", "LDSM",
       "ELECT", "ACQBULK", "UCGABAR", "MUFU", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    so = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "atoma_infer_b200" / "lib" / "libflashattention.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "")
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("atoma::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    out = {}
    for name, c in kernels.items():
        total = sum(c.values())
        out[name] = {"instructions": total, **{k: c[k] for k in KEY if c[k]}}
    (ROOT / "profiles" / "r02_sass_opcodes.json").write_text(json.dumps(out, indent=1))
    cols = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "HMMA", "LDGSTS"]
    print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in cols))
    for name, rec in out.items():
        print(f"{name[:78]:78s} {rec['instructions']:6d} " + " ".join(f"{rec.get(c, 0):7d}" for c in cols))


if __name__ == "__main__":
    main()
