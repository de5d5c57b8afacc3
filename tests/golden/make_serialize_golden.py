"""Extracts, from the reference checkout, the ordered `dst.write_*` calls of every in-tree `serialize` body the engine's
stream format restates (generators: gates/mul_nonnative.rs:326-330, 533-536; gadgets/nonnative.rs:647-652, 730-740, 812-817,
874-878; gadgets/biguint.rs:520-526; gadgets/glv.rs:144-150 -- gates: gates/mul_nonnative.rs:181-183, 466-468) and writes
them to tests/golden/serialize_bodies.json.  Run in the build container (needs /root/reference); the JSON travels."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FILES = ["src/gates/mul_nonnative.rs", "src/gadgets/nonnative.rs", "src/gadgets/biguint.rs", "src/gadgets/glv.rs"]


def extract(ref=REF):
    out = {}
    for f in FILES:
        src = open(os.path.join(ref, f)).read()
        # impl blocks: remember the type the serialize fn belongs to
        for m in re.finditer(r"impl<[^{]*?>\s*(?:SimpleGenerator<F, D>|Gate<F, D>)\s*for\s*(\w+)[^{]*\{", src):
            name = m.group(1)
            body = src[m.end():]
            s = re.search(r"fn serialize\(&self, dst: &mut Vec<u8>[^{]*\{", body)
            if not s:
                continue
            depth, i = 1, s.end()
            while depth:
                depth += {"{": 1, "}": -1}.get(body[i], 0)
                i += 1
            calls = re.findall(r"dst\s*\.\s*(write_\w+)\(\s*&?(?:self\.)?([\w.()]+)", body[s.end():i])
            out[name] = [[prim, re.sub(r"\.to_target_vec\(\)|\.0\b|[()]", "", arg)] for prim, arg in calls]
    return out


if __name__ == "__main__":
    data = extract()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "serialize_bodies.json")
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))
