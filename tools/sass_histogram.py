"""Opcode histogram of the hot kernels of libpvb200.so (cuobjdump -sass): the mnemonics that prove which hardware
features a kernel is built from.  python tools/sass_histogram.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "predict_pv_yield_b200", "libpvb200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FFMA", "HMMA", "F2FP",
         "USETMAXREG", "UCGABAR_ARV", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"^@!?U?P\w+\s+", "", line.split(None, 1)[1]).split()[0].rstrip(";")
            base = ins.split(".")[0]
            kernels[name][base] += 1
            kernels[name]["#total"] += 1
            if base in ("UTCHMMA", "UTCBAR") and "2CTA" in ins:
                kernels[name][base + ".2CTA"] += 1
    print(f"# {os.path.relpath(SO, ROOT)}: SASS opcode counts of the kernels that use the tensor cores, the TMA engine or packed FMA")
    print(f"{'kernel':<64}" + "".join(f"{w[:11]:>12}" for w in WATCH) + f"{'instr':>8}")
    for k, c in kernels.items():
        if not any(c[w] for w in ("UTCHMMA", "UBLKCP", "UTMALDG", "UTMASTG", "UTMAREDG", "FFMA2")):
            continue
        short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().replace("pvb::", "")
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        print(f"{short[:63]:<64}" + "".join(f"{c[w]:>12}" for w in WATCH) + f"{c['#total']:>8}")


if __name__ == "__main__":
    main()
