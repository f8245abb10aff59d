"""Summarise `nvcc -Xptxas -v` output of csrc/b200mf.cu: registers / spills / smem per kernel."""
import re, subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
csrc = os.path.join(here, "..", "matrix-free-dealii-precice_b200", "csrc")
cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
       "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-shared", "-o", "/tmp/b200mf_info.so"] + sys.argv[1:] + [os.path.join(csrc, "b200mf.cu")]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"entry function '(\S+)'", out)), capture_output=True, text=True).stdout.split("\n")
names = iter(filt)
for line in out.split("\n"):
    if "entry function" in line:
        cur = next(names)
        cur = re.sub(r"b200mf::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
    elif "spill" in line:
        sp = re.findall(r"(\d+) bytes", line)
    elif "Used" in line and cur:
        regs = re.search(r"Used (\d+) registers", line).group(1)
        print(f"{cur:90s} regs={regs:>3s} stack={sp[0]:>4s} spill={sp[1]}/{sp[2]}")
        cur = None
