"""Registers / spills / shared memory of every kernel, from the -Xptxas -v logs the Makefile leaves next to the objects
(sirius_b200/csrc/*.o.ptxas.log, untracked: they carry compile times).  Writes profiles/r1_ptxas_registers.txt."""
import glob, os, re, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
for log in sorted(glob.glob(os.path.join(root, "sirius_b200", "csrc", "*.o.ptxas.log"))):
    out.append(f"== {os.path.basename(log).replace('.o.ptxas.log', '.cu')}")
    name, spill = None, ""
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            continue
        if name and "spill" in line:
            spill = line.strip()
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and name:
            smem = re.search(r"(\d+) bytes smem", m.group(2))
            out.append(f"   {name:110s} regs={m.group(1):>3s} smem={smem.group(1) if smem else 0:>6}  {spill}")
            name, spill = None, ""
open(os.path.join(root, "profiles", "r1_ptxas_registers.txt"), "w").write("\n".join(out) + "\n")
print(f"{len(out)} lines")
