"""Summarise `nvcc -Xptxas=-v` output: registers / spills per kernel (design aid)."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "primus_fhe_b200", "csrc")
def report(src):
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-Xptxas=-v", "-c", os.path.join(CSRC, src), "-o", "/tmp/_rep.o"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    names = subprocess.run(["c++filt"], input=out, stdout=subprocess.PIPE, text=True).stdout
    cur = None; rows = []
    for line in names.splitlines():
        m = re.search(r"Compiling entry function '(.*)' for", line)
        if m:
            cur = re.sub(r"\(.*", "", m.group(1)).replace("void pfhe::", "").replace("unsigned long", "u64").replace("unsigned int", "u32"); continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur: spill = (m.group(1), m.group(2), m.group(3)); continue
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            rows.append((cur, int(m.group(1)), spill)); cur = None
    for r in sorted(rows): print(f"{r[0]:60s} regs={r[1]:4d} stack/spill_st/spill_ld={r[2]}")
    if "error" in out: print(out)
if __name__ == "__main__":
    for s in sys.argv[1:]: report(s)
