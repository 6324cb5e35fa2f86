"""Loop census of a kernel's SASS (no GPU needed): python tools/sass_loops.py <obj|cubin> <mangled-name substring> [-l]

For every backward branch: the body's instruction count and opcode histogram.  Used to keep the instruction
budget of the hot loops in check before spending GPU time (the iteration kernel is issue-bound: instructions per
pixel decide its time).  -l lists the largest loop's instructions.
"""
import collections
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    t = ins.split()
    op = t[1] if t[0].startswith("@") else t[0]
    p = op.split(".")
    if p[0] in ("LDG", "STG", "LDS", "STS", "LDTM", "STTM", "IMAD", "BAR", "BRA"):
        return ".".join(p[:2]) if len(p) > 1 and p[1] in ("WIDE", "x4", "x8", "128", "64", "ARV", "SYNC", "U") else p[0]
    return p[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    listing = "-l" in sys.argv
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        print("==", name[:120], len(ins), "instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, s) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)*(?:!?U?P\d,\s*)?0x([0-9a-f]+)", s)
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr:
                    loops.append((addr[t], i))
        for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:6]:
            c = collections.Counter(opcode(s) for _, s in ins[lo:hi + 1])
            print("  loop %5d..%5d  %4d instr  %s" % (lo, hi, hi - lo + 1, " ".join("%s:%d" % kv for kv in c.most_common(40))))
        if listing and loops:
            lo, hi = sorted(loops, key=lambda x: x[0] - x[1])[int(sys.argv[sys.argv.index("-l") + 1]) if len(sys.argv) > sys.argv.index("-l") + 1 else 0]
            for a, s in ins[lo:hi + 1]:
                print("    %05x  %s" % (a, s))


if __name__ == "__main__":
    main()
