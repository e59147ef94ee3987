"""SASS opcode histogram and integer-pipe clock budget of the kernels in an object / binary.

  python tools/sass_hist.py <file.o|.bin|.so> [name-regex] [--loop]

Per kernel: instruction count, histogram by opcode, and the issue cost on the two integer pipes with the
rates measured by profiles/microbench/int_pipe.cu (IMAD.WIDE / IMAD.HI: 4 clk on the FMA-heavy pipe, other
IMAD: 2 clk there; IADD3 / LOP3 / SHF / SEL / ISETP / PRMT / LEA / VIADD / MOV: 2 clk on the ALU pipe).
--loop restricts the count to the innermost backward-branch loop body (the register-resident microbenchmarks
time exactly that body).
"""
import collections
import re
import subprocess
import sys

FMA_WIDE = ("IMAD.WIDE", "IMAD.HI")
ALU = ("IADD3", "LOP3", "SHF", "SEL", "ISETP", "PRMT", "LEA", "VIADD", "MOV", "IABS", "VIMNMX", "IMNMX", "FLO", "POPC", "BREV", "I2I")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, rows = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            text = m.group(2).strip()
            pred = re.match(r"(@!?U?P\d+)\s+(.*)", text)
            if pred:
                text = pred.group(2)
            rows.append((int(m.group(1), 16), text))
    if name:
        yield name, rows


def loop_body(rows):
    best = None
    for addr, text in rows:
        m = re.match(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    if not best:
        return rows
    return [r for r in rows if best[0] <= r[0] <= best[1]]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    loop = "--loop" in sys.argv
    path = args[0]
    pat = re.compile(args[1]) if len(args) > 1 else None
    for name, rows in kernels(path):
        dn = demangle(name)
        if pat and not pat.search(dn):
            continue
        if loop:
            rows = loop_body(rows)
        ops = collections.Counter()
        fma_clk = alu_clk = 0
        for _, text in rows:
            op = text.split()[0]
            if op == "NOP":
                continue
            ops[op] += 1
            if op.startswith(FMA_WIDE):
                fma_clk += 4
            elif op.startswith("IMAD") or op.startswith("IDP"):
                fma_clk += 2
            elif op.startswith(ALU):
                alu_clk += 2
        total = sum(ops.values())
        print("## %s" % dn)
        print("instructions%s: %d   FMA-heavy pipe clk: %d   ALU pipe clk: %d" % (" (loop body)" if loop else "", total, fma_clk, alu_clk))
        print("  " + "  ".join("%s %d" % kv for kv in ops.most_common()))
        print()


if __name__ == "__main__":
    main()
