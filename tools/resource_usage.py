"""Per-kernel resource usage of librecsys_b200.so for profiles/: registers, stack (= spill frame), static shared memory, local memory
(`cuobjdump --dump-resource-usage`; runs without a GPU).  With c++filt-demangled, shortened names.

usage: python tools/resource_usage.py profiles/r02_resource_usage.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "recsys-examples_b200", "lib", "librecsys_b200.so")


def main(out):
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows, src = [], "?"
    for line in txt.splitlines():
        m = re.match(r"identifier = .*/([\w.]+)$", line.strip())
        if m:
            src = m.group(1)
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m:
            rows.append((src, name, *map(int, m.groups())))
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    def shorten(n):
        n = n.replace("void ", "").replace("(anonymous namespace)::", "")
        n = re.sub(r"\((?!anonymous).*", "", n)                      # drop the argument list
        return re.sub(r"\b\w+::", "", n)[:70]
    short = [shorten(n) for n in names]
    lines = [f"# registers / stack bytes (spill frame) / static shared bytes / local bytes per kernel of {os.path.relpath(LIB, ROOT)}",
             "# (cuobjdump --dump-resource-usage; dynamic shared memory is set at launch and not listed)",
             "# source".ljust(20) + "kernel".ljust(72) + "REG".rjust(5) + "STACK".rjust(7) + "SHARED".rjust(8) + "LOCAL".rjust(7)]
    for (s, _, reg, stack, shared, local), n in sorted(zip(rows, short), key=lambda x: (x[0][0], x[1])):
        lines.append(s.ljust(20) + n.ljust(72) + str(reg).rjust(5) + str(stack).rjust(7) + str(shared).rjust(8) + str(local).rjust(7))
    spilling = [(s, n, st) for (s, _, _, st, _, _), n in zip(rows, short) if st > 0]
    lines.append(f"# {len(rows)} kernels; {len(spilling)} with a stack frame: " + ", ".join(f"{n} ({st} B)" for _, n, st in sorted(spilling, key=lambda x: -x[2])[:12]))
    open(out, "w").write("\n".join(lines) + "\n")
    print(lines[-1])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_resource_usage.txt"))
