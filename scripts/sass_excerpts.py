"""Writes profiles/sass_r02_{pops,nn,gscan}.txt: excerpts of the SASS of the hot kernels in clustering_b200/libdcb200.so
(cuobjdump -sass; no GPU needed) -- the loops that prove what the kernels are made of: UBLKCP (cp.async.bulk / TMA),
SYNCS (mbarrier), the FFMA blocks, UTCHMMA (tcgen05.mma) and LDTM (tcgen05.ld).

    python scripts/sass_excerpts.py
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "clustering_b200", "libdcb200.so")


def instrs(fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, SO], capture_output=True, text=True).stdout
    out = []
    for ln in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def count(ins, pats):
    return {p: sum(1 for _, i in ins if re.search(p, i)) for p in pats}


def excerpt(ins, lo, hi):
    return "\n".join(f"  /*{a:05x}*/ {i}" for a, i in ins if lo <= a <= hi)


def densest(ins, pat, window):
    """start address of the `window`-instruction stretch with the most matches of pat"""
    hits = [1 if re.search(pat, i) else 0 for _, i in ins]
    best, arg, run = -1, 0, sum(hits[:window])
    for k in range(len(ins) - window):
        if run > best:
            best, arg = run, k
        run += hits[k + window] - hits[k]
    return arg


def mangled(pattern):
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    return sorted(set(re.findall(r"Function : (\S*" + pattern + r"\S*)", txt)))


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    # ---- multi-radius population kernel (C3)
    ins = instrs("_ZN3dcb15pops_bin_kernelILi10EEEvNS_8PopsArgsE")
    f2 = [q for q in range(len(ins)) if ins[q][1].startswith("FFMA2")]
    k = f2[0] - 6
    votes = [q for q in range(f2[-1], len(ins)) if ins[q][1].startswith("VOTE.ANY")]
    v = votes[0]
    ublk = [a for a, i in ins if "UBLKCP" in i]
    with open(os.path.join(ROOT, "profiles", "sass_r02_pops.txt"), "w") as f:
        f.write("pops_bin_kernel<10>: the multi-radius population scan of C3 (cuobjdump -sass clustering_b200/libdcb200.so, sm_100a)\n")
        f.write(f"instructions: {len(ins)}; {count(ins, ['^FFMA2', r'^FFMA ', 'UBLKCP', 'SYNCS', r'BAR\.', r'^LDS', r'^STS', 'FMNMX3', 'LDL|STL'])}\n\n")
        f.write("-- producer: ONE bulk copy (UBLKCP = cp.async.bulk, 1-D TMA) per column tile, completion counted on an mbarrier;\n"
                "   ring stages come back through named barriers (consumers BAR.ARV, producer BAR.SYNC) --\n")
        for a in ublk:
            f.write(excerpt(ins, a - 0x60, a + 0x20) + "\n")
        for a in [a for a, i in ins if i.startswith("BAR.ARV")][:1]:
            f.write(excerpt(ins, a - 0x20, a + 0x10) + "\n")
        f.write("\n-- step: one broadcast LDS.128 per dim + 8 FFMA2 (packed FP32 FMA: rows r, r+1 x one column, the column operand a broadcast scalar),\n"
                "   candidate test (FMNMX / FMNMX3, FADD |x'|^2, FSETP), warp vote --\n")
        f.write(excerpt(ins, ins[k][0], ins[v][0] + 0x30) + "\n")
        f.write("\n-- dense step, first two columns (8 pairs): s = acc + |x'|^2 (FADD), cell (FMUL.SAT, FFMA + 2^21, LOP3), table entry (LDS),\n"
                "   s - e (FADD), band (FSETP |.|), bin from the entry's low bits and the sign (LOP3, SHF, IADD3, LEA/IMAD), histogram LDS.U16 / add / STS.U16 --\n")
        d0 = [q for q in range(v, len(ins)) if ".SAT" in ins[q][1]][0]
        f.write(excerpt(ins, ins[d0][0] - 0x40, ins[d0][0] + 0x5c0) + "\n")
    # ---- neighbour kernel (C3)
    ins = instrs("_ZN3dcb9nn_kernelILi10EEEvNS_6NnArgsE")
    f2 = [q for q in range(len(ins)) if ins[q][1].startswith("FFMA2")]
    k = f2[0] - 6
    sat = [q for q in range(k, len(ins)) if "FADD.SAT" in ins[q][1]]
    ublk = [a for a, i in ins if "UBLKCP" in i]
    with open(os.path.join(ROOT, "profiles", "sass_r02_nn.txt"), "w") as f:
        f.write("nn_kernel<10>: the neighbour scan of C3 (cuobjdump -sass clustering_b200/libdcb200.so, sm_100a)\n")
        f.write(f"instructions: {len(ins)}; {count(ins, ['^FFMA2', r'^FFMA ', 'UBLKCP', 'SYNCS', r'BAR\.', r'^LDS', 'ATOM', 'FADD.SAT', 'LDL|STL'])}\n\n")
        f.write("-- producer: bulk copies of the tile record and of the tile's free-energy ranks --\n")
        for a in ublk:
            f.write(excerpt(ins, a - 0x40, a + 0x20) + "\n")
        f.write("\n-- step: LDS.128 + FFMA2 block, first filter level (min tree FMNMX / FMNMX3 against max(t_nn, t_hd), FSETP, branch); then, for a\n"
                "   block that passed, the free-energy-aware filter per pair: FADD.SAT (rank difference -> {0,1}), FFMA (t_nn + cand * dl), FSETP --\n")
        end = ins[sat[-1]][0] + 0x80 if sat else ins[k][0] + 0xe00
        f.write(excerpt(ins, ins[k][0], min(end, ins[k][0] + 0x1100)) + "\n")
    # ---- GEMM-form scans (C5)
    names = mangled("gscan_pops_kernel") + mangled("gscan_nn_kernel")
    with open(os.path.join(ROOT, "profiles", "sass_r02_gscan.txt"), "w") as f:
        f.write("GEMM-form scans (17 <= n_cols <= 256, C5): tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, bulk TMA = UBLKCP\n\n")
        for fn in names[:1] + names[-1:]:
            ins = instrs(fn)
            f.write(f"{fn}\ninstructions: {len(ins)}; {count(ins, ['UTCHMMA', 'LDTM', 'UBLKCP', 'SYNCS', 'UTCBAR|UTCATOM', '^FFMA'])}\n")
            mm = [q for q in range(len(ins)) if "UTCHMMA" in ins[q][1]]
            if mm:
                f.write("-- MMA issue loop (descriptors in uniform registers, UTCHMMA back to back, commit to an mbarrier) --\n")
                f.write(excerpt(ins, ins[mm[0]][0] - 0x80, ins[mm[0]][0] + 0x180) + "\n")
            ld = [q for q in range(len(ins)) if "LDTM" in ins[q][1]]
            if ld:
                f.write("-- epilogue: accumulator rows out of tensor memory (LDTM), then the count / filter arithmetic --\n")
                f.write(excerpt(ins, ins[ld[0]][0] - 0x20, ins[ld[0]][0] + 0x200) + "\n")
            f.write("\n")
    for n in ("pops", "nn", "gscan"):
        p = os.path.join(ROOT, "profiles", f"sass_r02_{n}.txt")
        print(p, sum(1 for _ in open(p)), "lines")


if __name__ == "__main__":
    main()
