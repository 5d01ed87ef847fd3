"""List the loops (backward branches) of a kernel's SASS with their instruction mix: python scratch/sass_loops.py <lib.so> <mangled-substring>"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in blk.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("==", name, len(ins), "instructions")
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\(\S+\)|0x([0-9a-f]+))", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = [x for _, x in ins[addr_index[tgt]: i + 1]]
                def cnt(p): return sum(1 for x in body if re.search(p, x))
                if cnt(r"SHFL") >= 4:
                    print(f"  loop 0x{tgt:05x}..0x{a:05x}: {len(body):4d} instr  DFMA/DMUL/DADD {cnt(r'\bD(FMA|MUL|ADD)')}  FFMA {cnt(r'\bF(FMA|MUL|ADD)')}  SHFL {cnt('SHFL')}  "
                          f"BAR {cnt(r'BAR')}  LDS {cnt(r'LDS')}  STS {cnt(r'STS')}  LDL {cnt('LDL')}  STL {cnt('STL')}  SEL {cnt(r'SEL')}  LDG {cnt('LDG')}  BRA {cnt('BRA')}")
