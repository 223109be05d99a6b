"""Dev aid: instruction mix of the innermost MUFU-bearing loop of each gls_strip_kernel variant."""
import re, subprocess, sys, collections
so = sys.argv[1] if len(sys.argv) > 1 else "periodicity_b200/lib/libperiodicity_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "gls_strip_kernel"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+([^;]+);", f)]
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                body = [x for x in ins if tgt <= x[0] <= a]
                ops = [x[1] for x in body]
                if any("MUFU.SIN" in o for o in ops) and not any(o.startswith("BAR") for o in ops):
                    if best is None or len(body) < len(best):
                        best = body
    if not best:
        print(name, "no loop found"); continue
    c = collections.Counter()
    for _, t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        c[t.split()[0].split(".")[0]] += 1
    tot = sum(c.values())
    print(name.split("(")[0][:70], "loop instr:", tot, dict(c.most_common()))
