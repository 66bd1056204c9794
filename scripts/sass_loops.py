"""Lists the loops (backward branches) of one kernel's SASS with their instruction mix, to see at a glance whether the
inner loop of a scan kernel holds spill traffic (LDL/STL) and what its FFMA share is.

    python scripts/sass_loops.py clustering_b200/csrc/build/kern_d10.o pops_bin_kernel
"""
import re, subprocess, sys, collections

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name, len(ins), "instructions")
    addr = [a for a, _ in ins]
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.\S+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a:
                loops.append((tgt, a))
    for lo, hi in sorted(set(loops), key=lambda x: x[1] - x[0]):
        body = [t for a, t in ins if lo <= a <= hi]
        def op(t):
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            return t.split()[0].split(".")[0]
        c = collections.Counter(op(t) for t in body)
        if c["FFMA"] < 16:
            continue
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
        print(f"  loop {lo:#06x}..{hi:#06x}: {len(body)} instr, FFMA {c['FFMA']} ({100 * c['FFMA'] / len(body):.0f} %), LDL {c['LDL']}, STL {c['STL']}, LDS {c['LDS']}, STS {c['STS']}, LDG {c['LDG']}\n      {top}")
