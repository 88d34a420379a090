"""Write profiles/r2_sass_excerpt.txt: per selected kernel of the shipped library the opcode histogram and every tensor-core /
tensor-memory / TMA / bulk-copy / mbarrier / system-scope instruction with its address; library-wide counts of the Blackwell-specific
opcodes at the end.  python scripts/sass_excerpt.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "diffusion-llm-rs_b200", "lib", "libdllm_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_excerpt.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
SELECT = [r"quant_d_rows_ring_kernelILi256ELi4ELi[01]E", r"gemv_mma_kernelILi4ELi1ELi3ELi3ELi2ELb1", r"gemv_mma_kernelILi2ELi1ELi3ELi3ELi4ELb1",
          r"umma_qlinear_pair2_kernelILi4ELi3ELb0", r"umma_qlinear_pair2_kernelILi4ELi3ELb1", r"rowquant_i8_warp_kernel",
          r"p2p_allreduce_kernelILb1ELi2ELi4", r"p2p_reduce_gather_kernelILi2ELi4"]
KEEP = re.compile(r"UTC|UTMA|UBLKCP|LDTM|STTM|SYNCS|ELECT|IMMA|HMMA|UCGABAR|ACQBULK|\.SYS|MULTIMEM|REDG|ACTIVEMASK.*SYS|FENCE|CCTL|ERRBAR|MEMBAR")
SPECIAL = re.compile(r"^(ACQBULK|ARRIVES|IMMA|HMMA|LDGSTS|LDTM|STTM|UBLKCP|UCGABAR_ARV|UCGABAR_WAIT|UTCATOMSWS[.A-Z0-9_]*|UTCBAR[.A-Z0-9]*|UTCHMMA[.A-Z0-9]*|"
                     r"UTCIMMA[.A-Z0-9]*|UTMACCTL|UTMACMDFLUSH|UTMALDG|UTMASTG)")
lines_out = ["# cuobjdump -sass diffusion-llm-rs_b200/lib/libdllm_b200.so (sm_100a), round-2 final build: per kernel the opcode histogram and every",
             "# tensor-core / tensor-memory / TMA / bulk-copy / mbarrier / system-scope instruction with its address (scripts/sass_of.py prints the "
             "full listing; this file: scripts/sass_excerpt.py)", ""]
total = collections.Counter()
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    ins = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?;)", l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    for _, t in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        m = SPECIAL.match(op)
        if m:
            key = m.group(1)
            key = re.sub(r"^(IMMA|HMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|LDGSTS|UTMACCTL)\..*", r"\1", key)
            total[key] += 1
    if not any(re.search(p, name) for p in SELECT):
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in ins)
    lines_out.append(f"== {name}")
    lines_out.append(f"   {len(ins)} instructions")
    lines_out.append(f"   histogram: {dict(ops.most_common(28))}")
    for addr, t in ins:
        if KEEP.search(t):
            lines_out.append(f"   {addr}  {t}")
    lines_out.append("")
lines_out.append("== whole library, counts of the Blackwell-specific opcodes")
for k in sorted(total):
    lines_out.append(f"   {k}: {total[k]}")
open(out_path, "w").write("\n".join(lines_out) + "\n")
print("wrote", out_path, len(lines_out), "lines")
