"""SASS instruction histograms of the hot kernels (cuobjdump on the built objects; runs without a GPU).
python profiles/sass_histogram.py > profiles/r2_sass_histograms.txt
Shows what the kernels are made of: packed FP32 (FFMA2 / FADD2 / FMUL2), MUFU, 64/128-bit LDG/STG/LDS/STS, barriers,
L2 prefetches (CCTL / UBLKPF); no tcgen05 / UTMA: these are HBM/L2-streaming butterfly kernels, not GEMMs (DESIGN section 8
records the TMA-ring attempt)."""
import collections
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "qumcmc_b200", "csrc")
WANT = [("sweep.o", r"lo_sweep_kernelILi1ELi64ELb1ELb0"), ("sweep.o", r"hiq_sweep_kernelILb1ELb0"), ("sweep.o", r"lo_sweep_kernelILi0ELi128ELb1ELb0"),
        ("sweep.o", r"lo_sweep_kernelILi2ELi128ELb1ELb0"), ("sweep.o", r"lo_sweep_kernelILi1ELi64ELb1ELb1"), ("sweep.o", r"hiq_sweep_kernelILb1ELb1"),
        ("sweep.o", r"hi8_sweep_kernelILi1ELi8ELi8"), ("sweep128.o", r"sweep128_kernelINS_5LoGeoELi1ELb0"),
        ("sweep128.o", r"sweep128_kernelINS_5HiGeoILi2EEELi1ELb0"), ("general.o", r"gen_pass_kernelIfE"), ("general.o", r"gen_pass_kernelIdE"),
        ("small_n.o", r"warp_chain_kernelIfLi4E"), ("small_n.o", r"warp_chain_kernelIdLi5E")]
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL", "MUFU", "I2FP", "I2F", "IADD3", "LOP3", "LDG", "STG",
        "LDS", "STS", "LDC", "LDCU", "BAR", "CCTL", "UBLKPF", "SHFL", "BRA", "MOV", "LDL", "STL"]
for obj, pat in WANT:
    path = os.path.join(ROOT, obj)
    names = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = [m.group(1) for m in re.finditer(r"Function : (\S+)", names) if re.search(pat, m.group(1))]
    for fn in funcs[:1]:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, path], capture_output=True, text=True).stdout
        ops = collections.Counter()
        wide = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", sass, re.M):
            ops[m.group(1)] += 1
            if m.group(1) in ("LDG", "STG", "LDS", "STS") and (".64" in m.group(2) or ".128" in m.group(2)):
                wide[m.group(1) + (".128" if ".128" in m.group(2) else ".64")] += 1
        total = sum(ops.values())
        demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
        print("%s\n  %d instructions: %s" % (demangled, total, "  ".join("%s %d" % (k, ops[k]) for k in KEYS if ops[k])))
        print("  wide memory ops: %s" % "  ".join("%s %d" % kv for kv in sorted(wide.items())))
