"""Cost model of the hot loop of a kernel from its SASS (cuobjdump -sass): instruction mix, FP32-pipe cycles and
register-file operand reads.  On B200 a scheduler reads about two 32-bit register operands per lane and clock
(profiles/microbench/pipes_b200.txt: FFMA with three distinct register operands 2.47 instead of 3.67 warp-inst/clk/SM,
FFMA2 with three distinct pairs 1.31 instead of 1.94), so a loop costs at least max(pipe cycles, operand reads / 2).
Usage: python profiles/tools/sass_loop_cost.py <object or .so> <substring of the mangled kernel name> [unroll]"""
import re
import subprocess
import sys
from collections import Counter


def kernel_sass(obj, key):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if key in name:
            return name, b
    raise SystemExit(f"no kernel matching {key}")


def parse(b):
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def reads_of(text):
    s = text
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op, _, rest = s.partition(" ")
    ops = [o.strip() for o in rest.split(",")]
    base = op.split(".")[0]
    srcs = ops if base in ("ST", "STS", "STG", "BRA", "LDGSTS", "RED", "ATOMS") else ops[1:]
    seen = set()   # a register that appears twice in one instruction is read once
    for o in srcs:
        regs = re.findall(r"(?<![UP])\bR(\d+)", o)
        for r in regs:
            seen.add(int(r))
            if "F32x2" in o or ".64" in o:
                seen.add(int(r) + 1)
    rd = len(seen)
    return base, rd


def main():
    obj, key = sys.argv[1], sys.argv[2]
    unroll = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    name, b = kernel_sass(obj, key)
    ins = parse(b)
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_idx:
                body = ins[addr_idx[tgt]:i + 1]
                packed = sum(1 for _, x in body if re.search(r"\bF(ADD|MUL|FMA)2\b", x))
                if len(body) < 900 and (best is None or packed > best[0]):   # the unrolled loop that holds the packed arithmetic
                    best = (packed, body)
    packed, body = best
    c = Counter()
    reads = pipe = 0
    for _, t in body:
        base, rd = reads_of(t)
        c[base] += 1
        reads += rd
        if base in ("FADD2", "FMUL2", "FFMA2"):
            pipe += 2
        elif base in ("FADD", "FMUL", "FFMA", "IMAD"):
            pipe += 1
    print(name[:150])
    print(f"hot loop: {len(body)} instructions per {unroll} iterations = {len(body) / unroll:.1f} per iteration")
    print(f"  FP32-pipe cycles per iteration {pipe / unroll:.1f}; register operand reads per iteration {reads / unroll:.1f}"
          f" -> {reads / 2 / unroll:.1f} clk at 2 reads/clk; MUFU {c['MUFU'] / unroll:.1f}")
    print("  mix per iteration: " + " ".join(f"{k}={v / unroll:.1f}" for k, v in c.most_common(16)))


if __name__ == "__main__":
    main()
