"""Static evidence for profiles/: per kernel of libpdephi_b200.so, how many tcgen05 MMAs (UTC*MMA), TMEM loads / stores
(LDTM / STTM), TMA transfers (UTMALDG / UTMASTG / UBLKCP), cp.async copies (LDGSTS) and legacy tensor instructions (HMMA)
its SASS holds, with registers and spills from the ptxas logs.  Runs without a GPU (cuobjdump on the built library)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pde_fluid_phi_b200", "lib", "libpdephi_b200.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
            ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("HMMA", r"\bHMMA"),
            ("FFMA", r"\bFFMA\b"), ("FFMA2", r"\bFFMA2"), ("SYNCS", r"\bSYNCS")]


def ptxas_resources():
    """{mangled name: (registers, spill store bytes, spill load bytes)} from the build's ptxas -v logs."""
    res, build = {}, os.path.join(ROOT, "pde_fluid_phi_b200", "build")
    for fn in sorted(os.listdir(build)) if os.path.isdir(build) else []:
        if not fn.endswith(".ptxas.log"):
            continue
        text = open(os.path.join(build, fn)).read()
        for m in re.finditer(r"Function properties for (\S+)\s+\d+ bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s+"
                             r"ptxas info\s+: Used (\d+) registers", text):
            res[m.group(1)] = (int(m.group(4)), int(m.group(2)), int(m.group(3)))
    return res


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    mangled = re.findall(r"Function : (\S+)", sass)
    names = subprocess.run(["cu++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")
    resources, raw_of = ptxas_resources(), {}
    counts, order, cur, total = {}, [], None, collections.Counter()
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it)
            cut = cur.rfind(">(")
            cur = cur[:cut + 1] if cut >= 0 else cur.split("(")[0]
            cur = re.sub(r"\((?:bool|int)\)", "", cur).replace("void ", "").replace("pdephi::", "")
            base, k = cur, 2
            while cur in counts:
                cur = f"{base} #{k}"; k += 1
            counts[cur] = collections.Counter()
            raw_of[cur] = m.group(1)
            order.append(cur)
            continue
        if cur and "/*" in line:
            counts[cur]["instructions"] += 1
            for key, pat in PATTERNS:
                if re.search(pat, line):
                    counts[cur][key] += 1
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "registers", "spill store bytes", "spill load bytes", "instructions"] + [k for k, _ in PATTERNS])
        for k in sorted(order):
            w.writerow([k, *resources.get(raw_of[k], ("", "", "")), counts[k]["instructions"]] + [counts[k][p] for p, _ in PATTERNS])
            total.update(counts[k])
    print(f"{len(order)} kernels; totals:", {p: total[p] for p, _ in PATTERNS})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r1f_sass_census.csv"))
