"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), so that the tcgen05 / TMEM / TMA evidence is in the
tree. Run here, no GPU needed:

    python tools/sass_histogram.py clover_b200/libclover_b200.so > profiles/r02_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

KEY = ["UTCQMMA", "UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "IDP", "FFMA2", "FMUL2", "FADD2", "FMNMX3",
       "F2IP", "I2IP", "PRMT", "HMMA", "IMMA", "LDGSTS", "REDUX", "MATCH", "ATOMS", "MEMBAR", "CCTL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "clover_b200/libclover_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    print(f"# {lib}: SASS opcode histogram per kernel (cuobjdump -sass; tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG / UBLKCP,")
    print("# mbarrier = SYNCS, dp4a = IDP.4A, packed fp32 = FFMA2 / FMUL2 / FADD2). Legacy tensor opcodes (HMMA / IMMA) must be absent.")
    total = collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        ops = collections.Counter()
        n = 0
        for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", f):
            n += 1
            op = m.group(1)
            full = op + m.group(2)
            if op in KEY:
                ops[full if op in ("UTCQMMA", "UTMALDG", "IDP", "LDTM", "SYNCS", "MEMBAR") else op] += 1
                total[op] += 1
        keys = ", ".join(f"{k} {v}" for k, v in sorted(ops.items()))
        print(f"{demangled[:88]:88s} {n:6d} instr | {keys}")
    print("# totals: " + ", ".join(f"{k} {v}" for k, v in sorted(total.items())))


if __name__ == "__main__":
    main()
