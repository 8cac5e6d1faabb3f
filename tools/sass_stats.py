"""SASS opcode statistics of libbfp_b200.so (cuobjdump -sass), per kernel.
    python tools/sass_stats.py --summary [--out profiles/r02_sass_opcodes.txt]   one line per kernel family + Blackwell opcode counts
    python tools/sass_stats.py --kernel 'quant_stream_kernel<0, 1, 4, 2, 0, true, false>' [--loop]   opcode histogram of one kernel
Pipe classes follow the measured split (B300_MICROARCH.md): fma = FFMA/FMUL/FADD/IMAD/HFMA2..., alu = IADD3/LOP3/SHF/PRMT/
FMNMX/ISETP/SEL..., xu = MUFU/F2I/I2F/FRND, lsu = LDG/STG/LDS/STS/SHFL."""
import argparse, collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantization-sparsity-interplay_b200", "libbfp_b200.so")
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HMUL2", "HADD2", "FFMA2", "FMUL2", "FADD2", "HSET2", "HSETP2", "HMNMX2")
XU = ("MUFU", "F2I", "I2F", "FRND", "F2F", "I2FP", "F2FP", "F2IP", "POPC", "FLO", "BREV")
LSU = ("LDG", "STG", "LDS", "STS", "SHFL", "LDC", "LD", "ST", "ATOM", "RED", "LDSM", "STSM", "ULDC", "LDCU")
BLACKWELL = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR",
             "HMMA", "IMMA", "LDGSTS", "SYNCS")


def load():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)[1:]
    names = [p.split("\n", 1)[0].strip() for p in parts]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return {d: p for d, p in zip(dem, parts)}


def instrs(body):
    out = []
    for line in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            out.append((int(m.group(1), 16), m.group(3), line))
    return out


def pipe(op):
    base = op.split(".")[0]
    if base in FMA: return "fma"
    if base in XU: return "xu"
    if base in LSU: return "lsu"
    if base in ("BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "NOP", "BAR", "WARPSYNC", "YIELD", "DEPBAR", "ACQBULK", "S2R", "S2UR", "CS2R", "ELECT"): return "ctl"
    if base.startswith("U") and base not in ("UTCHMMA", "UTCIMMA"): return "uniform"
    return "alu"


def hist(ins):
    h = collections.Counter(op.split(".")[0] for _, op, _ in ins)
    p = collections.Counter(pipe(op) for _, op, _ in ins)
    return h, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--summary", action="store_true")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--loop", action="store_true", help="restrict to the largest backward-branch loop body")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    K = load()
    lines = []
    if a.kernel:
        for name, body in K.items():
            if a.kernel in name:
                ins = instrs(body)
                if a.loop:
                    best = None
                    for addr, op, line in ins:
                        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", line)
                        if op.startswith("BRA") and m and int(m.group(1), 16) < addr:
                            span = (int(m.group(1), 16), addr)
                            if best is None or span[1] - span[0] > best[1] - best[0]: best = span
                    if best: ins = [i for i in ins if best[0] <= i[0] <= best[1]]
                h, p = hist(ins)
                lines.append(f"{name}: {len(ins)} instructions; pipes {dict(p)}")
                lines.append("  " + ", ".join(f"{k} {v}" for k, v in h.most_common()))
    if a.summary:
        fam = collections.defaultdict(lambda: [0, collections.Counter()])
        for name, body in K.items():
            f = re.sub(r"<.*", "", name.replace("void ", ""))
            ins = instrs(body)
            fam[f][0] += 1
            for _, op, _ in ins:
                b = op.split(".")[0]
                if b in BLACKWELL or op.startswith("LDG.E") or op.startswith("STG.E"):
                    key = b if b in BLACKWELL else op.split(".CONSTANT")[0]
                    if ".2CTA" in op: key += ".2CTA"
                    fam[f][1][key] += 1
        lines.append("kernel family: instantiations; tcgen05 / TMEM / TMA / vector-memory opcodes summed over the family's SASS")
        for f in sorted(fam):
            n, c = fam[f]
            lines.append(f"{f}: {n} instantiation(s); " + (", ".join(f"{k} {v}" for k, v in sorted(c.items())) or "-"))
    text = "\n".join(lines)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
