#!/usr/bin/env python3
"""Writes profiles/<tag>_sass_hot_loop.md: the SASS of the traversal loop of k_frame (one if-if step: the unified fetch,
the node step, the leaf test, the loop tail) with, per instruction, what `ncu --set full --import-source on` sampled:
share of the kernel's warp-stall samples, executions, average active threads, dominant stall reasons.
usage: sass_hot_loop.py <tag> <source.csv.gz from profiles/capture_ncu.sh> [kernel index]"""
import csv, gzip, io, os, sys
tag, path = sys.argv[1:3]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kern.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
k = kern[which]; h = k["hdr"]; ix = {n: i for i, n in enumerate(h)}; R = k["rows"]
S = lambda r: int(r[ix["# Samples"]]); E = lambda r: int(r[ix["Instructions Executed"]])
tot_s, tot_i = sum(map(S, R)), sum(map(E, R))
# the loop: from the first unified LDG.E.128.CONSTANT back to the loop head (the ISETP on the node code) and on to the
# backward branch
first = next(i for i, r in enumerate(R) if "LDG.E.128.CONSTANT" in r[1])
head = first
while head > 0 and "0x7ffffffe" not in R[head][1]: head -= 1
addr = lambda i: int(R[i][0], 16)
end = next(i for i in range(first, len(R)) if "BRA" in R[i][1] and "0x" in R[i][1] and
           int(R[i][1].split("0x")[-1].split()[0].rstrip(";"), 16) <= addr(head) and i > first + 40)
loop = R[head - 1:end + 2]
ls, li = sum(map(S, loop)), sum(map(E, loop))
stall = {}
for r in R:
    for n in h:
        if n.startswith("stall_") and "Not Issued" not in n: stall[n] = stall.get(n, 0) + int(r[ix[n]])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{tag}_sass_hot_loop.md")
with open(out, "w") as f:
    f.write(f"# {tag}: SASS of the traversal loop, `{k['name']}`\n\n")
    f.write("`cuobjdump`-equivalent listing from the `ncu --set full --import-source on` capture of one 4K frame (profiles/capture_ncu.sh); "
            "columns: share of the kernel's warp-stall samples, warp-level executions, average active threads, dominant stall reasons.\n\n")
    f.write(f"* kernel: {tot_i:,} warp instructions, {tot_s:,} samples; this loop: {100 * li / tot_i:.1f} % of the instructions, "
            f"{100 * ls / tot_s:.1f} % of the samples\n")
    f.write("* kernel-wide stall mix: " + ", ".join(f"{n[6:]} {100 * v / tot_s:.1f} %" for n, v in sorted(stall.items(), key=lambda x: -x[1])[:8]) + "\n")
    f.write("* structure: loop head -> address select (node record or the leaf's first triangle) -> ONE group of three LDG.E.128 -> "
            "node lanes: 4th load, 12 FFMA + 20 FMNMX/FMNMX3 slab tests of both children, order, push (STL) / pop (LDL) -> "
            "leaf lanes: the strict Cramer triangle test (reference triangle.cpp:10-24) -> reconverge, loop tail\n\n")
    f.write("| # | SASS | samples | executed | threads | stalls |\n|---|---|---|---|---|---|\n")
    for j, r in enumerate(loop):
        st = {n[6:]: int(r[ix[n]]) for n in h if n.startswith("stall_") and "Not Issued" not in n and int(r[ix[n]]) > 0}
        st = ", ".join(f"{a} {b}" for a, b in sorted(st.items(), key=lambda x: -x[1])[:2])
        f.write(f"| {j} | `{r[1].strip()}` | {100 * S(r) / tot_s:.2f} % | {E(r):,} | {r[ix['Avg. Threads Executed']]} | {st} |\n")
print(out, len(loop), "instructions", f"{100 * li / tot_i:.1f}% of instructions")
