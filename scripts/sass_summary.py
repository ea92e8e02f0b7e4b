"""Per-kernel SASS evidence (profiles/sass_summary.md): counts of the Blackwell-native mnemonics in every kernel of
loongx_b200/lib/libloongx_b200.so, from `cuobjdump -sass` (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UTMAPF, mma.sync -> HMMA)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "loongx_b200" / "lib" / "libloongx_b200.so"
out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "sass_summary.md"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTC\w*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UTMAREDG", r"\bUTMAREDG"), ("UTMAPF", r"\bUTMAPF"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                                ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("MUFU.EX2", r"\bMUFU\.EX2"),
                                ("FFMA2", r"\bFFMA2"), ("REDG/ATOMG", r"\b(REDG|ATOMG|RED\.)"), ("LDG", r"\bLDG"), ("STG", r"\bSTG")])
cur, counts, arch = None, collections.OrderedDict(), set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("lx::", "")
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
with open(out, "w") as f:
    f.write(f"# SASS summary of `{lib.relative_to(ROOT)}` (`cuobjdump -sass`, arch {', '.join(sorted(arch))})\n\n")
    f.write("Instruction counts per kernel (static occurrences in the SASS listing).  `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = "
            "tcgen05.ld / .st, `UTMALDG` / `UTMASTG` / `UTMAREDG` / `UTMAPF` = TMA load / store / reduce / L2 prefetch, `HMMA` = "
            "mma.sync (legacy tensor path, used only by the small text-encoder attention).\n\n")
    keys = list(pats)
    f.write("| kernel | " + " | ".join(keys) + " |\n|---|" + "---|" * len(keys) + "\n")
    tot = collections.Counter()
    for name, c in counts.items():
        tot.update(c)
        f.write(f"| `{name[:80]}` | " + " | ".join(str(c.get(k, 0)) for k in keys) + " |\n")
    f.write("| **total** | " + " | ".join(str(tot.get(k, 0)) for k in keys) + " |\n")
print(out, len(counts), "kernels")
