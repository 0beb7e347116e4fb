"""Opcode histogram of the unchecked (steady-state) four-second group of a k_tracks_fast kernel:
the region from the first uniform branch after the main-loop head to the jump over the checked copy.
    python tools/loop_hist.py "460549ELi4ELb1ELb0"   """
import collections, re, subprocess, sys
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", "em_model_manned_bayes_b200/libemb200.so"], capture_output=True, text=True).stdout
fn = [f for f in re.split(r"\n\s*Function : ", txt)[1:] if pat in f.split("\n", 1)[0]][0]
ins = [(int(a, 16), b.strip()) for a, b in re.findall(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", fn, re.M)]
back = [(int(re.search(r"0x([0-9a-f]+)", s).group(1), 16), a) for a, s in ins if "BRA" in s and re.search(r"0x([0-9a-f]+)", s)
        and int(re.search(r"0x([0-9a-f]+)", s).group(1), 16) < a]
lo, hi = max(back, key=lambda x: x[1] - x[0])
body = [(a, s) for a, s in ins if lo <= a <= hi]
# first forward BRA.U in the loop jumps to the checked copy; the unconditional BRA before that target ends the unchecked copy
first = next((a, int(re.search(r"0x([0-9a-f]+)", s).group(1), 16)) for a, s in body if s.startswith("BRA.U") or "BRA.U" in s)
un = [(a, s) for a, s in body if first[0] < a < first[1]]
tail_start = next(int(re.search(r"0x([0-9a-f]+)", s).group(1), 16) for a, s in reversed(un) if s.startswith("BRA "))
tail = [(a, s) for a, s in body if a >= tail_start]
def hist(rows):
    c = collections.Counter()
    for _, s in rows:
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0]
        op = re.sub(r"\.(U32|LUT|AND|OR|reuse|E|128|64|EF|STRONG|GPU|CONSTANT)", "", op)
        c[op] += 1
    return c
hu, ht = hist(un), hist(tail)
print("loop 0x%x..0x%x; unchecked group %d instr + store tail %d instr = %.1f per second" % (lo, hi, len(un), len(tail), (len(un) + len(tail)) / 4))
print("group:", ", ".join("%s %d" % kv for kv in hu.most_common()))
print("tail :", ", ".join("%s %d" % kv for kv in ht.most_common()))
alu = {"LOP3", "IADD3", "VIADD", "ISETP", "ISETP.GE", "ISETP.NE", "ISETP.GT", "FSEL", "LEA", "SEL", "MOV", "PRMT", "SHF", "PLOP3", "IADD3.X", "LEA.HI", "ISETP.LT", "ISETP.EQ"}
tot = hu + ht
a = sum(v for k, v in tot.items() if k.split(".")[0] in {x.split(".")[0] for x in alu})
f = sum(v for k, v in tot.items() if k.split(".")[0] in {"IMAD", "FFMA", "FADD", "FMUL", "HFMA2"})
print("alu-pipe %d (%.1f/s)  fma-pipe %d (%.1f/s)  other %d" % (a, a / 4, f, f / 4, len(un) + len(tail) - a - f))
