"""Print (or summarise) the SASS of the kernels whose mangled name matches a regex.
Usage: python scripts/sass_of.py <regex> [--full]"""
import re
import subprocess
import sys
import os
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "diffusion-llm-rs_b200", "lib", "libdllm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
pat = re.compile(sys.argv[1])
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if not pat.search(name):
        continue
    lines = [l for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    print(name, len(lines), "instructions")
    if "--full" in sys.argv:
        print("\n".join(l.split("/*")[1].split("*/")[1].strip() + "  " for l in lines))
    else:
        print("  ", dict(ops.most_common(25)))
