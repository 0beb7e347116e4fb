"""Per-kernel SASS opcode histogram (whole function and the hottest loop = largest backward-branch span).
    python tools/sass_stats.py [lib.so] [substring of kernel name]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "em_model_manned_bayes_b200/libemb200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", re.M)
for f in funcs:
    name = f.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in dem:
        continue
    ins = [(int(a, 16), b) for a, b in ins_re.findall(f)]
    print("==", dem[:160], "instructions:", len(ins))
    # loops: backward branches
    loops = []
    for addr, s in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s.*0x([0-9a-f]+)", s)
        m2 = re.search(r"0x([0-9a-f]+)", s) if "BRA" in s else None
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < addr:
                loops.append((tgt, addr))
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:3]:
        body = [s for a, s in ins if lo <= a <= hi]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for s in body)
        print("  loop 0x%x..0x%x: %d instr: %s" % (lo, hi, len(body), ", ".join("%s %d" % kv for kv in ops.most_common(18))))
