"""Writes profiles/sass_summary.txt: counts of the SASS mnemonics that prove which hardware paths libcnnb200.so uses
(tcgen05 MMA, TMEM loads, TMA tensor loads, mbarrier arrive/wait, warp reductions), per kernel for the tensor path.
    python scripts/sass_summary.py            # needs cuobjdump (CUDA toolkit), no GPU
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "commonnnclustering_b200", "libcnnb200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "REDUX", "FMNMX3", "SHF.L.W", "DADD", "DMUL", "DFMA"]


def counts(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    total = {m: len(re.findall(r"\b" + re.escape(m), sass)) for m in MNEMONICS}
    per_kernel = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name:
            for mn in ("UTCHMMA", "LDTM", "UTMALDG"):
                if re.search(r"\b" + mn, line):
                    per_kernel.setdefault(name, {}).setdefault(mn, 0)
                    per_kernel[name][mn] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    return total, per_kernel, arch


if __name__ == "__main__":
    total, per_kernel, arch = counts()
    out = [f"cuobjdump -sass commonnnclustering_b200/libcnnb200.so  (arch: {', '.join(arch)})", "",
           "mnemonic      occurrences   meaning"]
    meaning = {"UTCHMMA": "tcgen05.mma (5th-generation tensor core MMA, accumulators in TMEM)",
               "LDTM": "tcgen05.ld (TMEM -> registers, the epilogue)",
               "UTMALDG": "cp.async.bulk.tensor (TMA tile load)",
               "UTCBAR": "tcgen05.commit (MMA completion -> mbarrier)",
               "SYNCS": "mbarrier arrive / try_wait",
               "REDUX": "warp-wide integer reduction (__reduce_*_sync)",
               "FMNMX3": "three-input min/max (smallest |t| of two candidates per instruction, fp32 filter)",
               "SHF.L.W": "funnel shift (sign bit of the distance test -> hit mask)",
               "DADD": "f64 add (exact recheck: no FMA contraction)", "DMUL": "f64 multiply (exact recheck)",
               "DFMA": "f64 FMA: only inside the division / square-root / floor sequences (distance sums are DMUL + DADD: --fmad=false)"}
    for m in MNEMONICS:
        out.append(f"{m:<13} {total[m]:>11}   {meaning[m]}")
    out += ["", "per kernel (tensor path):"]
    for k, v in sorted(per_kernel.items()):
        out.append(f"  {k[:100]:<100} " + "  ".join(f"{a}={b}" for a, b in sorted(v.items())))
    text = "\n".join(out) + "\n"
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    with open(path, "w") as fp:
        fp.write(text)
    sys.stdout.write(text)
