"""Print registers / shared memory / spills per kernel from build/obj/*.ptxas.log."""
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else ""
log = open("build/obj/ba_kernels.ptxas.log").read()
names = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'", log)
res = re.findall(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log)
spill = re.findall(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log)
for n, r, sp in zip(names, res, spill):
    dn = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    dn = re.sub(r"\(.*", "", dn)[:80]
    if re.search(pat, dn):
        print(f"{dn:82s} regs={r[0]:>3} smem={r[3] or 0:>6} stack={sp[0]:>4} spill={sp[1]}/{sp[2]}")
