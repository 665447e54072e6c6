"""Print registers / shared memory / spills per kernel from build/obj/*.ptxas.log (optional regex filter)."""
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else ""
log = open("build/obj/ba_kernels.ptxas.log").read()
blocks = re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]
rows = []
for b in blocks:
    name = b.split("'", 1)[0]
    used = re.search(r"Used (\d+) registers, used (\d+) barriers", b)
    smem = re.search(r"(\d+) bytes smem", b)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    rows.append((name, used.group(1) if used else "?", smem.group(1) if smem else "0", sp.groups() if sp else ("?", "?", "?")))
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (n, regs, smem, sp), dn in zip(rows, dem):
    dn = re.sub(r"\(.*", "", dn)[:84]
    if re.search(pat, dn):
        print(f"{dn:86s} regs={regs:>3} static_smem={smem:>6} stack={sp[0]:>4} spill={sp[1]}/{sp[2]}")
