#!/usr/bin/env python
"""tools/sass_of.py <object-or-so> <substring-of-mangled-name> — dump the SASS of the first matching kernel and an opcode histogram"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[2] in name:
        lines = [l for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        ops = [re.sub(r"/\*[0-9a-f]+\*/", "", l).strip().rstrip(";").strip() for l in lines]
        if len(sys.argv) > 3 and sys.argv[3] == "-v":
            print("\n".join(ops))
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", o).split(" ")[0].split(".")[0] for o in ops)
        print(name[:120])
        print(len(ops), "instructions;", ", ".join(f"{k} {v}" for k, v in hist.most_common(24)))
        break
