"""Regenerate profiles/r02_*_sass.txt from the built libantq.so (cuobjdump -sass; no GPU needed).

    python tools/sass_excerpts.py

For each kernel of interest: mnemonic histogram of the whole function and the hottest loop body (the backward branch
whose body holds the most instructions of the marker mnemonic).
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ant-quantization_b200", "csrc", "libantq.so")
OUT = os.path.join(ROOT, "profiles")

INSTR = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")


def functions():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur:
            out[cur].append(line)
    return out


def demangle(n):
    return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n


def mnemonic(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text.strip())
    return t.split()[0] if t else ""


def parse(lines):
    ins = []
    for l in lines:
        m = INSTR.search(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def loops(ins):
    """(start, end) address ranges of backward branches."""
    out = []
    for a, t in ins:
        if mnemonic(t).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a:
                out.append((int(m.group(1), 16), a))
    return out


def hot_loop(ins, marker):
    best, score = None, -1
    for s, e in loops(ins):
        body = [t for a, t in ins if s <= a <= e]
        sc = sum(1 for t in body if mnemonic(t).startswith(marker))
        # prefer the tightest loop that still holds most marker instructions
        if sc > 0 and (sc > score or (sc == score and best and e - s < best[1] - best[0])):
            best, score = (s, e), sc
    return best


def write(name, title, notes, fn_lines, marker, max_lines=220):
    ins = parse(fn_lines)
    hist = collections.Counter(mnemonic(t) for _, t in ins)
    with open(os.path.join(OUT, name), "w") as f:
        f.write("# cuobjdump -sass libantq.so -- %s\n" % title)
        for n in notes:
            f.write("# %s\n" % n)
        f.write("## mnemonic histogram (%d instructions)\n" % len(ins))
        for k, v in hist.most_common(28):
            f.write("%7d %s\n" % (v, k))
        rare = [k for k in hist if re.match(r"(UTC|UTMA|LDTM|UBLKCP|SYNCS|HFMA2|HSET|HMNMX|PRMT|UTCBAR|UTCATOM)", k)]
        f.write("## Blackwell / packed-16-bit mnemonics present: %s\n" % ", ".join("%s x%d" % (k, hist[k]) for k in sorted(rare)))
        hl = hot_loop(ins, marker)
        if hl:
            body = [(a, t) for a, t in ins if hl[0] <= a <= hl[1]]
            bh = collections.Counter(mnemonic(t) for _, t in body)
            f.write("## hot loop 0x%04x..0x%04x: %d instructions; %s\n" % (hl[0], hl[1], len(body),
                    ", ".join("%s x%d" % kv for kv in bh.most_common(14))))
            for a, t in body[:max_lines]:
                f.write("        /*%04x*/  %s ;\n" % (a, t))
            if len(body) > max_lines:
                f.write("        ... (%d more)\n" % (len(body) - max_lines))
    print("wrote", name, len(ins), "instructions")


def pick(fns, *needles):
    for n, l in fns.items():
        d = demangle(n)
        if all(x in d for x in needles):
            return d, l
    raise SystemExit("kernel not found: %r" % (needles,))


def main():
    fns = functions()
    d, l = pick(fns, "antq_stream_kernel<__half, (int)7, (bool)1, (bool)0, (bool)0>")
    write("r02_stream_sass.txt", d + " (headline: signed 4-bit chain)",
          ["cp.async.bulk -> UBLKCP; mbarrier -> SYNCS.*; threshold chain = HSET2 / HFMA2 (two pipes) + LOP3"], l, "HSET2")
    for needles, name, note in (
            (("antq_pu_stream_kernel<__half, (bool)1, (bool)1, (bool)0>",), "r02_pu_sass.txt",
             "closed form, uniform grid (int-k), x-space clamp on packed halves (HMNMX2)"),
            (("antq_pu_stream_kernel<__half, (bool)0, (bool)1, (bool)0>",), "r02_pu_table_sass.txt",
             "closed form, per-octave table (flint / pot / float), x-space clamp")):
        d, l = pick(fns, *needles)
        write(name, d, [note, "per element: cvt (HADD2.F32), FMUL kx, FADD +M, FADD -M, FADD diff, FFMA/FSETP flag, FMUL *c*s ... F2FP pack"], l, "FADD")
    d, l = pick(fns, "antq_linear_p4_kernel<__half, (bool)0>")
    write("r02_gemm_sass.txt", d + " (dequant-fused Linear, 16-bit operands)",
          ["tcgen05.mma kind::f16 -> UTCHMMA; tcgen05.commit -> UTCBAR; cp.async.bulk.tensor.2d -> UTMALDG.2D;",
           "tcgen05.ld -> LDTM.x32; tcgen05.alloc/dealloc -> UTCATOMSWS; weight decoders' byte LUT -> PRMT"], l, "PRMT", 160)
    d, l = pick(fns, "antq_linear_p4_kernel<__half, (bool)1>")
    write("r02_gemm_fp8_sass.txt", d + " (dequant-fused Linear, e4m3 levels on both sides)",
          ["tcgen05.mma kind::f8f6f4 -> UTCQMMA; the rest as the 16-bit variant, one byte plane of PRMT decode"], l, "PRMT", 160)


if __name__ == "__main__":
    main()
