#!/usr/bin/env python
"""Static instruction count of a kernel's innermost-largest loop from `cuobjdump -sass` output.

The fused rollout kernel is bound by instruction issue (profiles/), so the number of SASS instructions in its
time-step loop is the quantity to minimise; this tool reports it (total and by pipe class) without a GPU.

    cuobjdump -sass mbt_gym_b200/libmbt_b200.so > /tmp/all.sass
    python tools/sass_loop_count.py /tmp/all.sass rollout Id ILi0ELi1ELi1ELi0ELi0ELi0EELb0ELi1E
"""
import collections
import re
import sys


def functions(path):
    for f in open(path).read().split("Function : ")[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)]
        yield name, ins


def classify(op):
    op = op.split()[0] if not op.startswith("@") else op.split()[1]
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP"):
        return "fp64"
    if base in ("FFMA", "FMUL", "FADD", "FSETP", "FSEL", "FMNMX", "FCHK"):
        return "fp32"
    if base in ("IMAD", "LOP3", "IADD3", "SHF", "PRMT", "ISETP", "SEL", "LEA", "VIADD", "MOV", "IABS", "I2FP", "CS2R"):
        return "int/alu"
    if base in ("MUFU", "I2F", "F2I", "F2F"):
        return "sfu/conv"
    if base in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU", "ATOMG", "RED"):
        return "mem"
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR"):
        return "ctrl"
    return "other:" + base


def main():
    path, keys = sys.argv[1], sys.argv[2:]
    for name, ins in functions(path):
        if not all(k in name for k in keys):
            continue
        loops = []
        for addr, text in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr:
                loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
        if not loops:
            print(name, "no loop")
            continue
        _, lo, hi = max(loops)
        body = [t for a, t in ins if lo <= a <= hi]
        hist = collections.Counter(classify(t) for t in body)
        print(f"{name}\n  total {len(ins)} instr; largest loop 0x{lo:x}..0x{hi:x}: {len(body)} instr  {dict(hist.most_common())}")


if __name__ == "__main__":
    main()
