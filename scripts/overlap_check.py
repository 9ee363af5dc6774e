#!/usr/bin/env python
"""Development check: share of this repo's host-layer code lines that also occur, whitespace removed, in the
same-named reference file (only runs where /root/reference exists; the product never reads it).

    python scripts/overlap_check.py            # prints one line per file, exits 1 above the 15 % bar
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAIRS = [("pypore_b200/core.py", "PyPore/core.py"), ("pypore_b200/DataTypes.py", "PyPore/DataTypes.py"),
         ("pypore_b200/parsers.py", "PyPore/parsers.py")]
BAR = {"pypore_b200/parsers.py": 0.20}


def code_lines(path):
    out = []
    in_doc = False
    for ln in open(path, encoding="utf-8", errors="replace"):
        s = re.sub(r"\s+", "", ln)
        q = s.count('"""') + s.count("'''")
        if in_doc:
            if q % 2 == 1:
                in_doc = False
            continue
        if q % 2 == 1:
            in_doc = True
            continue
        if not s or s.startswith("#") or q == 2:
            continue
        out.append(s)
    return out


def main():
    ref_root = os.environ.get("PYPORE_REFERENCE", "/root/reference")
    bad = False
    for mine, ref in PAIRS:
        a = code_lines(os.path.join(ROOT, mine))
        b = set(code_lines(os.path.join(ref_root, ref)))
        trivial = {"pass", "else:", "try:", "returnd", "delself", "continue", "break", "return"}
        hits = [s for s in a if s in b and s not in trivial]
        # the interface the drop-in must keep (names, signatures, defaults) is shared by construction
        body = [s for s in hits if not s.startswith(("def", "class", "@", "import", "from"))]
        frac = len(body) / max(len(a), 1)
        bar = BAR.get(mine, 0.15)
        print("%-28s %4d code lines, %4d shared with %s (%.1f %%), %d of them not signatures = %.1f %% (bar %.0f %%)" %
              (mine, len(a), len(hits), ref, 100.0 * len(hits) / max(len(a), 1), len(body), 100 * frac, 100 * bar))
        if "-v" in sys.argv:
            for s in hits:
                print("    ", s)
        bad = bad or frac >= bar
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
