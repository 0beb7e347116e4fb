"""Opcode histogram of the unchecked (steady-state) four-second group of a k_tracks_fast kernel.
    python tools/loop_hist.py [mangled-name substring, default the 7-variable uncor shape] [path to .so/.o]
The main loop is the largest backward branch; its first forward predicated branch jumps to the checked copy of the
group, the unconditional branch just before that target ends the unchecked copy and lands on the store tail."""
import collections
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "460549ELi4ELb1ELb0ELi0"
so = sys.argv[2] if len(sys.argv) > 2 else "em_model_manned_bayes_b200/libemb200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = [f for f in re.split(r"\n\s*Function : ", txt)[1:] if pat in f.split("\n", 1)[0]][0]
ins = [(int(a, 16), b.strip()) for a, b in re.findall(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", fn, re.M)]


def target(s):
    m = re.search(r"BRA.*?0x([0-9a-f]+)", s)
    return int(m.group(1), 16) if m else None


back = [(a - target(s), target(s), a) for a, s in ins if target(s) is not None and target(s) < a]
_, lo, hi = max(back)
body = [(a, s) for a, s in ins if lo <= a <= hi]
fwd = next((a, target(s)) for a, s in body if target(s) is not None and target(s) > a and s.startswith("@"))
un = [(a, s) for a, s in body if fwd[0] < a < fwd[1]]
tail_start = next(target(s) for a, s in reversed(un) if s.startswith("BRA "))
tail = [(a, s) for a, s in body if a >= tail_start]


def hist(rows):
    c = collections.Counter()
    for _, s in rows:
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0]
        c[re.sub(r"\.(U32|LUT|AND|OR|reuse|E|128|64|EF|STRONG|GPU|CONSTANT|W|R|GE|NE|GT|LT|EQ)\b", "", op)] += 1
    return c


hu, ht = hist(un), hist(tail)
n = len(un) + len(tail)
print("loop 0x%x..0x%x; unchecked group %d instr + store tail %d instr = %.1f per second" % (lo, hi, len(un), len(tail), n / 4))
print("group:", ", ".join("%s %d" % kv for kv in hu.most_common()))
print("tail :", ", ".join("%s %d" % kv for kv in ht.most_common()))
tot = hu + ht
alu = sum(v for k, v in tot.items() if k.split(".")[0] in {"LOP3", "IADD3", "VIADD", "ISETP", "FSEL", "LEA", "SEL", "MOV", "PRMT", "SHF", "PLOP3"})
wide = sum(v for k, v in tot.items() if k.startswith("IMAD.WIDE") or k.startswith("UIMAD.WIDE"))
fma = sum(v for k, v in tot.items() if k.split(".")[0] in {"IMAD", "FFMA", "FADD", "FMUL"}) - sum(v for k, v in tot.items() if k.startswith("IMAD.WIDE"))
print("per second: alu-pipe %.1f instr (x2 cycles = %.0f)   IMAD.WIDE %.1f (x4 = %.0f) + other fma-pipe %.1f   other %.1f"
      % (alu / 4, alu / 2, wide / 4, wide, fma / 4, (n - alu - wide - fma) / 4))
