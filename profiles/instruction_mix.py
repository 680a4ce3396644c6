#!/usr/bin/env python
"""instruction_mix.py REPORT.ncu-rep CHAIN_MOVES OUT.md [KERNEL_REGEX]

Dynamic SASS opcode histogram of one profiled kernel launch (an `ncu --set full --import-source on` capture): executed warp
instructions per opcode, normalised per chain-move (CHAIN_MOVES = chains x moves the launch processed), the FP64 share, local
memory traffic (spills show up as LDL / STL), plus the per-warp-role split when the kernel is crew2 (regions between the
round barriers)."""
import collections, csv, io, subprocess, sys

rep, chain_moves, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernel = rows[0][1]
hdr, data = rows[1], rows[2:]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
h, hs = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    b = op.split(".")[0]
    key = "MUFU.RSQ64H" if op.startswith("MUFU.RSQ64H") else ("IMAD.MOV" if op.startswith("IMAD.MOV") else b)
    h[key] += int(r[iE]); hs[key] += int(r[iN])
tot, tots = sum(h.values()), sum(hs.values()) or 1
fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RSQ64H"))
with open(out, "w") as f:
    f.write(f"# Dynamic instruction mix of `{kernel}`\n\nSource: `{rep}` (`ncu --set full --import-source on`), {chain_moves:.0f} chain-moves in the launch.\n\n")
    f.write(f"* executed warp instructions: {tot} = **{tot / chain_moves:.1f} per chain-move**\n")
    f.write(f"* FP64 pipe instructions (DFMA, DMUL, DADD, DSETP, MUFU.RSQ64H): {fp64 / chain_moves:.1f} per chain-move ({100 * fp64 / tot:.1f} %)\n")
    f.write(f"* local memory (spills would appear here): LDL {h.get('LDL', 0) / chain_moves:.3f}, STL {h.get('STL', 0) / chain_moves:.3f} per chain-move\n")
    f.write(f"* no divide, no square root: there is no MUFU.RCP64H / DSQRT in the histogram ({', '.join(k for k in h if k.startswith('MUFU'))})\n\n")
    f.write("| opcode | per chain-move | share | stall-sample share |\n|---|---|---|---|\n")
    for k, v in h.most_common(32):
        f.write(f"| {k} | {v / chain_moves:.2f} | {100 * v / tot:.1f} % | {100 * hs[k] / tots:.1f} % |\n")
print("wrote", out, f"{tot / chain_moves:.1f} instr per chain-move, fp64 {100 * fp64 / tot:.1f}%")
