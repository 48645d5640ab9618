"""Join an ncu SASS-level source page with nvdisasm line info -> per-CUDA-source-line instruction / stall summary.

    python tools/ncu_lines.py <report.ncu-rep> <library.so> [kernel-substring] [top N]

ncu's CLI prints per-instruction metrics only for the SASS view; nvdisasm -g knows which source line every SASS
instruction came from.  Both list the instructions of a function in address order, so they are joined by position
inside each contiguous address range (function).  The library must be the build the report was taken with.
"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def sass_rows(rep, want=""):
    """rows of the first kernel section whose name contains `want` (a report may hold several kernels)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    sec = next((x for x in sections if want in x["name"]), sections[0])
    print(f"# kernel: {sec['name']}")
    ix = {h: i for i, h in enumerate(sec["hdr"])}
    return ix, sec["data"]


def disasm_functions(lib):
    """{function name: [(source line, sass text), ...]} for every function of every sm_100a cubin in the library"""
    funcs = {}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=td, capture_output=True)
        for cubin in Path(td).glob("*sm_100a*.cubin"):
            txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
            cur, line = None, 0
            for ln in txt.splitlines():
                m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
                if m:
                    cur = m.group(1); funcs[cur] = []; continue
                m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
                if m:
                    line = (m.group(1), int(m.group(2))); continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m and cur:
                    funcs[cur].append((line, m.group(2).strip()))
    return funcs


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ix, data = sass_rows(rep, sys.argv[3] if len(sys.argv) > 3 else "")
    funcs = disasm_functions(lib)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    # split the ncu rows into contiguous address ranges
    ranges, cur, prev = [], [], None
    for r in data:
        a = int(r[ix["Address"]], 16)
        if prev is not None and a != prev + 16:
            ranges.append(cur); cur = []
        cur.append(r); prev = a
    ranges.append(cur)
    by_len = collections.defaultdict(list)
    for name, ins in funcs.items():
        by_len[len(ins)].append(name)
    per_line = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    stalls = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = tot_s = 0.0
    for rg in ranges:
        names = by_len.get(len(rg), [])
        if not names:
            print(f"# unmatched range of {len(rg)} instructions", file=sys.stderr)
            continue
        ins = funcs[names[0]]
        for r, (line, _) in zip(rg, ins):
            e = per_line[(names[0][:40], line)]
            e[0] += f(r, "Instructions Executed"); e[1] += f(r, "# Samples")
            tot_i += f(r, "Instructions Executed"); tot_s += f(r, "# Samples")
            for s in stalls:
                v = f(r, s)
                if v:
                    e[2][s[6:]] += v
    print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
    print(f"{'inst%':>6s} {'samp%':>6s}  function:file:line  top stalls")
    for (fn, line), (ni, ns, st) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        s3 = " ".join(f"{k}={v / max(ns, 1):.0%}" for k, v in st.most_common(3))
        print(f"{ni / tot_i:6.1%} {ns / tot_s:6.1%}  {fn}:{line[0]}:{line[1]}  {s3}")


if __name__ == "__main__":
    main()
