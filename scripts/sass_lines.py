#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu source page (SASS view) to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k3_split > /tmp/k.csv
    python scripts/sass_lines.py /tmp/k.csv pypore_b200/libpypore_b200.so k3_split [top]

The i-th instruction of the ncu listing is matched with the i-th instruction of `nvdisasm -gi`
(inline line info, needs -lineinfo), innermost inlined location first.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_map(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # a group of "//## File" lines (innermost inlined location first, outermost last) precedes the
    # instructions it applies to; SASS_DEPTH picks the entry (0 = innermost, -1 = outermost call site)
    depth = int(os.environ.get("SASS_DEPTH", "0"))
    out, active, loc, group, in_group = [], False, None, [], False
    for ln in txt:
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            if not in_group:
                group, in_group = [], True
            group.append((os.path.basename(m.group(1)), int(m.group(2)), m.group(3)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            if in_group:
                own = [g for g in group if g[0].endswith((".cuh", ".cu"))] or group
                loc = own[depth] if -len(own) <= depth < len(own) else own[-1]
                in_group = False
            out.append(loc)
    return out


def main():
    src, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[h], rows[h + 1:]
    for j, r in enumerate(data):  # ncu prints the listing twice; keep the first copy
        if r and r[0] == "Kernel Name":
            data = data[:j]
            break
    ie, si, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    lm = line_map(lib, kernel)
    if len(lm) != len(data):
        print("warning: %d SASS instructions in the report, %d in the library" % (len(data), len(lm)))
    agg = {}
    for i, r in enumerate(data):
        loc = lm[i] if i < len(lm) and lm[i] else ("?", 0, "")
        key = loc[:2]
        a = agg.setdefault(key, [0.0, 0.0, 0])
        a[0] += float(r[ie] or 0)
        a[1] += float(r[si] or 0)
        a[2] += 1
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print("total warp instructions %.0f, samples %.0f" % (ti, ts))
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get('SASS_SORT') == 'samples' else 0])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print("%5.1f%% inst %5.1f%% samp %4d sass  %s:%d  %s" % (100 * a[0] / ti, 100 * a[1] / ts, a[2], f, l, text))


if __name__ == "__main__":
    main()
