"""Instruction mix of the innermost FP64 loops of a kernel (developer tool).
usage: python scripts/sass_loop.py <lib.so|binary> <mangled-name-substring> [--dump]"""
import re
import subprocess
import sys
from collections import Counter


def main():
  lib, pat = sys.argv[1], sys.argv[2]
  out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
  blocks = out.split("Function : ")
  for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
      continue
    ins = []
    for l in blk.splitlines():
      m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
      if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
      m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
      if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        nd = sum(1 for x in body if re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL)", x[1]))
        loops.append((len(body), nd, tgt, a))
    # innermost = loops not containing another loop with FP64 work
    inner = [l for l in loops if l[1] >= 16 and not any(o is not l and o[1] >= 16 and l[2] <= o[2] and o[3] <= l[3] for o in loops)]
    print(f"== {name}: {len(ins)} instructions, innermost FP64 loops: {len(inner)}")
    for n, nd, tgt, a in inner:
      body = [x for x in ins if tgt <= x[0] <= a]
      c = Counter(re.sub(r"^(@!?U?P\d+\s+)", "", x[1]).split()[0].split(".")[0] for x in body)
      d3 = 0
      for x in body:
        if re.match(r"(@!?U?P\d+\s+)?DFMA", x[1]):
          ops = x[1].split(None, 1)[1].split(",")
          regs = [o.strip() for o in ops[1:]]
          fresh = [r for r in regs if re.match(r"-?\|?R\d+", r) and ".reuse" not in r]
          if len(set(fresh)) >= 3:
            d3 += 1
      print(f"  loop {tgt:#x}-{a:#x}: {n} instrs, FP64 {nd} (DFMA with 3 non-reuse register operands: {d3}); mix: {dict(c.most_common(12))}")
      if "--dump" in sys.argv:
        for x in body:
          print("     ", x[1])


if __name__ == "__main__":
  main()
