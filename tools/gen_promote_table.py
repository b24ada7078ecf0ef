"""Regenerate hpt_b200/csrc/promote.h from tests/golden/promotion.json (see that header)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "tests/golden/promotion.json")))
D = d["dtypes"]; E = {n: i for i, n in enumerate(D)}
def tab(k):
    return "\n".join("    {" + ", ".join(f"{E[d[k][a][b]]:2d}" for b in D) + "},  // " + a for a in D)
path = os.path.join(ROOT, "hpt_b200/csrc/promote.h")
src = open(path).read()
import re
for name, key in (("kNormalOut", "normal_out"), ("kFloatOutBinary", "float_out_binary")):
    src = re.sub(r"(constexpr signed char %s\[13\]\[13\] = \{\n)(.*?)(\n\};)" % name,
                 lambda m: m.group(1) + tab(key) + m.group(3), src, flags=re.S)
src = re.sub(r"(constexpr signed char kFloatOutUnary\[13\] = \{)(.*?)(\};)",
             lambda m: m.group(1) + ", ".join(str(E[d["float_out_unary"][a]]) for a in D) + m.group(3), src)
open(path, "w").write(src)
print("updated", path)
