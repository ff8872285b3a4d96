"""SASS evidence from the built library (runs here, no GPU): python profiles/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200", "csrc", "libgsr_b200.so")


def symbols():
    out = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
    return sorted(set(re.findall(r"_ZN3gsr[0-9A-Za-z_]+", out)))


def sass(sym):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    rows = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\*", line)
        if m:
            rows.append((m.group(1), m.group(2).rstrip(" ;")))
    return rows


def mix(rows, top=14):
    c = {}
    for _, ins in rows:
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        c[op] = c.get(op, 0) + 1
    return ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])[:top])


def find(prefix):
    return [s for s in symbols() if prefix in s]


def section(title, sym, pattern=None, window=None):
    rows = sass(sym)
    print(f"\n=== {title}\n    {sym}\n    {len(rows)} instructions; mix: {mix(rows)}")
    if pattern:
        for a, ins in rows:
            if re.search(pattern, ins):
                print(f"    {a}  {ins}")
    if window:
        lo, hi = window
        for a, ins in rows:
            if lo <= int(a, 16) <= hi:
                print(f"    {a}  {ins}")


if __name__ == "__main__":
    print("cuobjdump -sass of csrc/libgsr_b200.so (sm_100a), excerpts.  Mnemonics: FFMA2/FMUL2/FADD2 = packed FP32 (Blackwell);")
    print("LDGSTS = cp.async; UBLKCP = cp.async.bulk (TMA unit); SYNCS.* = mbarrier; LDG.E.128 = 16-byte global loads.")
    b = "_ZN3gsr12blend_kernelILi4ELb0ELb0ELb0ELi%dEEE"
    section("blend_kernel<4,0,0,0,kStageWarp> (default): staging + the first record of the unrolled group",
            find(b % 2)[0], window=(0x0, 0x0))
    rows = sass(find(b % 2)[0])
    first_lds = next(i for i, (_, ins) in enumerate(rows) if ins.startswith("LDS.128"))
    for a, ins in rows[first_lds - 2:first_lds + 92]:
        print(f"    {a}  {ins}")
    print("    ... LDGSTS (cp.async) staging of the per-warp batches:")
    for a, ins in rows:
        if "LDGSTS" in ins or "LDGDEPBAR" in ins or "DEPBAR" in ins:
            print(f"    {a}  {ins}")
    section("blend_kernel<4,0,0,0,kStageBulk>: TMA bulk copies + mbarrier (north_star item 3; measured slower, not the default)",
            find(b % 1)[0], pattern=r"UBLKCP|SYNCS|FENCE|R2UR")
    section("preprocess_prepared_kernel<3>: 15 coalesced 16-byte loads per Gaussian, no FP64",
            find("preprocess_prepared_kernelILi3E")[0], pattern=r"LDG|STG|ATOMS|DFMA|DMUL")
    section("prepare_kernel: the view-independent half (FP64 exp, quaternion normalisation), once per model",
            find("14prepare_kernel")[0], pattern=r"STG|DFMA")
    section("onesweep_kernel<24,false>: ballot ranking, look-back", find("onesweep_kernelILi24ELb0E")[0],
            pattern=r"ATOM|LDG|STG|BAR")
    rows = sass(find("onesweep_kernelILi24ELb0E")[0])
    r2p = [i for i, (_, ins) in enumerate(rows) if ins.startswith("R2P")]
    print("    ... one key of the ranking loop (R2P sets the digit's bits as predicates; per bit VOTE + SEL + LOP3):")
    for a, ins in rows[r2p[8]:r2p[9] + 1]:
        print(f"    {a}  {ins}")
    section("scan_emit_kernel", find("scan_emit_kernel")[0], pattern=r"VOTE|ATOM|LDG|STG|BAR")
