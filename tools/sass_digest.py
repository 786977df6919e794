"""SASS digest of the built library (no GPU needed): per kernel the instruction count, registers,
shared memory, the commonest mnemonics and the ones that prove the hardware paths the design names
(TMA tile loads, mbarriers, cluster barriers and distributed shared memory, three-input min / max,
128-bit global accesses, L2 cache hints), plus the inner loops of the two kernels the round-1
review asked about.

    python tools/sass_digest.py <tag>        ->  profiles/<tag>_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "katsdpsigproc_b200", "_lib", "libksp_b200.so")
PROOF = ["UTMALDG", "SYNCS", "UCGABAR", "FMNMX3", "LDGSTS",
         "REDUX", "MATCH", "VOTE", "SHFL", "ATOMS", "MUFU", "DFMA", "DADD", "F2F", "PRMT", "MEMBAR", "CCTL",
         "MAPA", "LD.E", "ST.E"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    tag = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(2).strip())
    names = demangle(list(funcs))
    short = lambda n: re.sub(r"\(anonymous namespace\)::", "", names[n]).split("(")[0].replace("void ", "")  # noqa: E731
    lines = [f"# SASS digest of katsdpsigproc_b200/_lib/libksp_b200.so ({tag}; tools/sass_digest.py, cuobjdump -sass; all code sm_100a)",
             f"# {len(funcs)} kernels.  Per kernel: instructions, resource usage, commonest mnemonics, and counts of the",
             "# mnemonics that show TMA (UTMALDG), mbarriers (SYNCS), cluster barriers (UCGABAR) / distributed shared",
             "# memory (MAPA, LD.E / ST.E through the cluster window), three-input min / max (FMNMX3), 128-bit accesses.", ""]
    for f, ins in funcs.items():
        ops = collections.Counter()
        proof = collections.Counter()
        for i in ins:
            body = re.sub(r"^@!?U?P\d+\s+", "", i)
            op = body.split()[0]
            ops[op.split(".")[0]] += 1
            for p in PROOF:
                if op.startswith(p):
                    proof[p] += 1
            base = op.split(".")[0]
            if base in ("LDG", "STG", "LDS", "STS") and ".128" in op:
                proof[base + " 128-bit"] += 1
        lines.append(f"## {short(f)}")
        lines.append(f"   {len(ins)} instructions; {usage.get(f, '')}")
        lines.append("   top: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(12)))
        if proof:
            lines.append("   of note: " + ", ".join(f"{k} {v}" for k, v in sorted(proof.items())))
        lines.append("")

    def excerpt(pattern, anchor, before, after, title):
        for f, ins in funcs.items():
            if re.search(pattern, names[f]):
                idx = [k for k, i in enumerate(ins) if anchor in i]
                if not idx:
                    continue
                # densest window of the anchor mnemonic
                best = max(idx, key=lambda k: sum(1 for j in idx if k <= j < k + after))
                lo, hi = max(0, best - before), min(len(ins), best + after)
                lines.append(f"## inner loop: {title} ({short(f)}, instructions {lo}..{hi} of {len(ins)})")
                lines.extend("   " + i for i in ins[lo:hi])
                lines.append("")
                return

    excerpt(r"bg13_kernel<0, 0, true, 256>", "FMNMX3", 8, 150, "phase 2 of the background tile: shared selection network (FMNMX3 / FMNMX) and the deviations")
    excerpt(r"madnz_stream_kernel", "LDG.E.128", 2, 90, "the pass over the row: three compares, two predicated counters, predicated append per key")
    excerpt(r"threshold_sum_kernel<true, 128, 1>", "UTMALDG", 12, 12, "TMA tile load of a span and its mbarrier")
    excerpt(r"maskedsum_kernel<false, 2>", "UCGABAR", 20, 30, "cluster barrier and the reads of the other blocks' partial strips")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_digest.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(path, len(lines), "lines")


if __name__ == "__main__":
    main()
