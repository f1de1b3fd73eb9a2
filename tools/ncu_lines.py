"""Executed instructions and stall samples per SOURCE LINE of one kernel of an `ncu --set full --import-source on`
report, with inlined code attributed to the innermost line of the chosen source file.

    python tools/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING SOURCE.cu [--launch N] [--top 40]

The report's source page is SASS only; the line table comes from `nvdisasm -gi` on the cubin inside OBJECT.o (built
with -lineinfo), joined on the instruction offsets.  KERNEL_SUBSTRING is matched against the mangled name of the
.text section (e.g. `sw_full_cs_quad_kernelILb0ELb0`).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def line_table(obj, kernel, src_name):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=d, capture_output=True, text=True).stdout
    table, infn, group, ingroup, cur = {}, False, [], False, None
    for ln in dis.split("\n"):
        if ln.startswith("//----") or ln.startswith(".text."):
            infn = kernel in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if not ingroup:
                group, ingroup = [], True
            group.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            if ingroup:
                ingroup = False
                own = [g for g in group if g[0].endswith(src_name)]
                cur = own[0][1] if own else None
            table[int(m.group(1), 16)] = (cur, m.group(2))
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("source")
    ap.add_argument("--launch", type=int, default=0, help="which captured launch of the kernel (in report order)")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--name", default="", help="substring of the demangled kernel name in the report (default: from KERNEL)")
    a = ap.parse_args()
    table = line_table(a.obj, a.kernel, os.path.basename(a.source))
    page = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    name = a.name or re.sub(r"^\d+", "", a.kernel).split("_kernel")[0] + "_kernel"
    mine = [i for i in heads if re.search(r"\b" + re.escape(name), rows[i][1])]
    start = mine[a.launch]
    end = min([h for h in heads if h > start] + [len(rows)])
    cols = rows[start + 1]
    c_exec, c_samp = cols.index("Instructions Executed"), cols.index("# Samples")
    body = [r for r in rows[start + 2:end] if r and r[0].startswith("0x")]
    base = int(body[0][0], 16)
    by_line, samp, by_op, total = collections.Counter(), collections.Counter(), collections.Counter(), 0
    for r in body:
        line, op = table.get(int(r[0], 16) - base, (None, "?"))
        n = int(r[c_exec])
        total += n
        by_line[line] += n
        samp[line] += int(r[c_samp])
        by_op[op.split(".")[0]] += n
    src = open(a.source).read().split("\n")
    print(rows[start][1])
    print("warp instructions executed:", total)
    print("opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / total) for k, v in by_op.most_common(16)))
    ts = max(1, sum(samp.values()))
    for line, n in by_line.most_common(a.top):
        text = src[line - 1].strip()[:100] if line else "(other files)"
        print("%5s  instr %5.1f%%  samples %5.1f%%  %s" % (line, 100.0 * n / total, 100.0 * samp[line] / ts, text))


if __name__ == "__main__":
    main()
