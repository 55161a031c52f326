#!/usr/bin/env python
"""Symbolic dataflow of the floating-point instructions in a SASS listing (cuobjdump -sass).

Why this exists: bit-exact `point_list`/`ranges` require that every float that feeds an integer
decision (depth key bits, tile rectangle, opacity thresholds) is produced by the *same sequence of
rounded operations* as in the reference build.  nvcc contracts a*b+c into FFMA twice (NVVM and
again ptxas), so the ground truth is SASS, not source or PTX.  This tool walks one function
linearly (branches ignored: use it on the straight-line main path of a kernel), tracks a
register -> expression map and prints the expression DAG of every value stored to global memory
(and, with --all-fp, of every FP instruction) in SSA form, so two kernels can be diffed by eye.

    python tools/sass_expr.py listing.sass --fun preprocessCUDAILi3ELb0ELb0 [--stores] [--grep FSETP]
"""
import argparse
import re
import sys

INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")

FP_OPS = ("FADD", "FMUL", "FFMA", "MUFU", "FMNMX", "FSEL", "FSETP", "F2F", "F2I", "I2F", "I2FP", "FRND",
          "DADD", "DMUL", "DFMA", "DSETP", "FCHK", "FSET", "FMNMX3", "F2FP", "FADD2", "FMUL2", "FFMA2")


def parse_function(path, fun_substr, nth=0):
    lines = open(path, errors="replace").read().split("\n")
    starts = [i for i, l in enumerate(lines) if "Function :" in l]
    cands = [i for i in starts if fun_substr in lines[i]]
    if not cands:
        raise SystemExit(f"no function matching {fun_substr!r}")
    s = cands[nth]
    e = next((j for j in starts if j > s), len(lines))
    out = []
    for l in lines[s:e]:
        m = INSTR.match(l)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return lines[s].strip(), out


class Sym:
    def __init__(self):
        self.regs = {}
        self.nodes = []  # (id, text)
        self.memo = {}

    def node(self, text):
        if text in self.memo:
            return self.memo[text]
        nid = f"n{len(self.nodes)}"
        self.nodes.append((nid, text))
        self.memo[text] = nid
        return nid

    def get(self, r):
        if r in ("RZ", "URZ"):
            return "0"
        if r == "PT":
            return "PT"
        return self.regs.get(r, f"{r}@in")

    def operand(self, tok):
        tok = tok.strip()
        neg = ""
        absv = False
        t = tok
        if t.startswith("-"):
            neg, t = "-", t[1:]
        if t.startswith("!"):
            neg, t = "!", t[1:]
        if t.startswith("|") and t.endswith("|"):
            absv, t = True, t[1:-1]
        t = re.sub(r"\.(reuse|H0_H0|H1_H1|64|F32|X4|X8|X16)$", "", t)
        t = re.sub(r"\.reuse", "", t)
        if re.fullmatch(r"U?R\d+|U?RZ|U?P\d+|PT", t):
            v = self.get(t)
        elif t.startswith("c["):
            v = t
        else:
            v = t  # immediate
        if absv:
            v = f"|{v}|"
        return neg + v

    def set(self, r, v):
        if r in ("RZ", "URZ", "PT"):
            return
        self.regs[r] = v


def regnum(r):
    m = re.fullmatch(r"(U?R)(\d+)", r)
    return (m.group(1), int(m.group(2))) if m else None


def split_ops(s):
    # split on commas not inside brackets
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "[(":
            depth += 1
        if ch in "])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def run(instrs, show_all_fp=False, grep=None, int_depth=False):
    S = Sym()
    events = []
    for addr, text in instrs:
        pred = None
        m = re.match(r"^@(!?U?P\d+)\s+(.*)$", text)
        if m:
            pred, text = m.group(1), m.group(2)
        parts = text.split(None, 1)
        op = parts[0]
        ops = split_ops(parts[1]) if len(parts) > 1 else []
        base = op.split(".")[0]
        if base in ("BRA", "EXIT", "BAR", "NOP", "BSSY", "BSYNC", "WARPSYNC", "CALL", "RET", "DEPBAR", "BMOV",
                    "ERRBAR", "MEMBAR", "CCTL", "YIELD", "NANOSLEEP", "BPT", "BREAK"):
            if grep and re.search(grep, op):
                events.append(f"{addr}: {('@'+pred+' ') if pred else ''}{text}")
            continue
        if base in ("STG", "ST", "STS", "STL", "RED", "ATOMG", "ATOMS", "REDG"):
            # STG.E [addr], Rv
            a = ops[0]
            v = ops[-1]
            width = 1
            if ".64" in op:
                width = 2
            if ".128" in op:
                width = 4
            rn = regnum(re.sub(r"\.reuse", "", v))
            vals = []
            if rn:
                for k in range(width):
                    vals.append(S.get(f"{rn[0]}{rn[1]+k}"))
            else:
                vals = [S.operand(v)]
            am = re.findall(r"U?R\d+", a)
            aexpr = a
            for r in am:
                aexpr = aexpr.replace(r, "<" + S.get(r) + ">", 1) if int_depth else aexpr
            events.append(f"{addr}: {('@'+pred+' ') if pred else ''}{op} {aexpr} <= {', '.join(vals)}")
            continue
        if not ops:
            continue
        dst = ops[0]
        srcs = ops[1:]
        # predicate-setting ops: dst are predicates
        if base in ("FSETP", "ISETP", "DSETP", "PLOP3", "UISETP", "FCHK", "VOTE", "VOTEU", "R2P", "UPLOP3"):
            sv = [S.operand(x) for x in srcs[1:]] if base != "FCHK" else [S.operand(x) for x in srcs]
            nid = S.node(f"{op}({', '.join(sv)})")
            S.set(dst, nid)
            if base in ("FSETP", "DSETP", "FCHK") or (grep and re.search(grep, op)):
                events.append(f"{addr}: {('@'+pred+' ') if pred else ''}{dst} = {nid}")
            continue
        if base in ("LDG", "LD", "LDS", "LDL", "LDC", "LDCU", "ULDC"):
            a = srcs[-1]
            regs_in = re.findall(r"U?R\d+", a)
            aexpr = a
            if base in ("LDG", "LD", "LDS", "LDL") or regs_in:
                for r in regs_in:
                    aexpr = aexpr.replace(r, "<" + S.get(r) + ">", 1)
            width = 1
            if ".64" in op:
                width = 2
            if ".128" in op:
                width = 4
            rn = regnum(dst)
            for k in range(width):
                nid = S.node(f"{base}({aexpr})+{4*k}")
                S.set(f"{rn[0]}{rn[1]+k}", nid)
            continue
        sv = [S.operand(x) for x in srcs]
        if base in ("MOV", "UMOV") or op.startswith("IMAD.MOV"):
            val = sv[-1] if base != "IMAD" else sv[1]
            if op.startswith("IMAD.MOV"):
                val = sv[-1]
            if pred:
                val = S.node(f"SEL({S.get(pred.lstrip('!'))}{'!' if pred.startswith('!') else ''}, {val}, {S.get(dst)})")
            S.set(dst, val)
            continue
        text_expr = f"{op}({', '.join(sv)})"
        if pred:
            text_expr = f"[{pred}:{S.get(pred.lstrip('!'))}]{text_expr} else {S.get(dst)}"
        nid = S.node(text_expr)
        is_d = base in ("DADD", "DMUL", "DFMA", "DMNMX") or (base == "F2F" and ".F64" in op.split("F2F")[1][:5])
        wide = ".WIDE" in op or base in ("DADD", "DMUL", "DFMA") or (base == "F2F" and op.startswith("F2F.F64"))
        rn = regnum(dst)
        S.set(dst, nid)
        if wide and rn:
            S.set(f"{rn[0]}{rn[1]+1}", nid + ".hi")
        if (show_all_fp and base in FP_OPS) or (grep and re.search(grep, op)):
            events.append(f"{addr}: {dst} = {nid}")
    return S, events


def reachable(S, events):
    idx = {nid: t for nid, t in S.nodes}
    need = set()
    stack = []
    for e in events:
        stack += re.findall(r"\bn\d+\b", e)
    while stack:
        n = stack.pop()
        if n in need:
            continue
        need.add(n)
        stack += re.findall(r"\bn\d+\b", idx[n])
    return [(nid, t) for nid, t in S.nodes if nid in need]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sass")
    ap.add_argument("--fun", required=True)
    ap.add_argument("--nth", type=int, default=0)
    ap.add_argument("--all-fp", action="store_true")
    ap.add_argument("--grep", default=None)
    ap.add_argument("--int-depth", action="store_true", help="expand integer address registers in stores")
    ap.add_argument("--fp-only", action="store_true", help="print only FP / load nodes")
    a = ap.parse_args()
    name, instrs = parse_function(a.sass, a.fun, a.nth)
    print("#", name[:160], f"({len(instrs)} instrs)")
    S, events = run(instrs, a.all_fp, a.grep, a.int_depth)
    nodes = reachable(S, events)
    for nid, t in nodes:
        if a.fp_only and not any(t.startswith(p) for p in FP_OPS + ("LDG", "LDC", "LDS", "[")):
            continue
        print(f"{nid} = {t}")
    print("# ---- events")
    for e in events:
        print(e)


if __name__ == "__main__":
    main()
