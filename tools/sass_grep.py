"""SASS evidence for profiles/: per-kernel counts of the Blackwell-specific mnemonics in librecsys_b200.so (cuobjdump -sass; runs without a GPU).

usage: python tools/sass_grep.py profiles/r02_sass_grep.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,
USETMAXREG = setmaxnreg (register rebalancing between specialised warps), HMMA = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "recsys-examples_b200", "lib", "librecsys_b200.so")
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTCBAR": r"\bUTCBAR",
       "SYNCS": r"\bSYNCS", "MATCH": r"\bMATCH", "MUFU.TANH": r"MUFU\.TANH", "USETMAXREG": r"USETMAXREG", "HMMA": r"\bHMMA", "ATOM/RED": r"\b(ATOM|RED|ATOMG)\b"}


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    lines = [f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass)", "# kernel".ljust(74) + " ".join(k.rjust(10) for k in PAT)]
    tot = collections.Counter()
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        c = {k: len(re.findall(v, f)) for k, v in PAT.items()}
        tot.update(c)
        short = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "", name)[:72]
        lines.append(short.ljust(74) + " ".join(str(c[k]).rjust(10) for k in PAT))
    lines.append("TOTAL".ljust(74) + " ".join(str(tot[k]).rjust(10) for k in PAT))
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    lines.append("# embedded cubins: " + ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))))
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"{out}: {len(funcs) - 1} kernels; " + ", ".join(f"{k}={tot[k]}" for k in PAT))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_grep.txt"))
