"""Per-kernel SASS instruction histogram of libb200fft.so (no GPU needed): load / store widths, TMA (UBLKCP, UTMA*,
UBLKPF), tensor-memory traffic (LDTM / STTM), packed FP32 (FADD2 / FMUL2 / FFMA2), barriers.
usage: python tools/sass_histogram.py [regex ...] > profiles/sass_r02_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opencl_fft_b200", "lib", "libb200fft.so")
want = [re.compile(a) for a in sys.argv[1:]] or [re.compile(
    r"fft_sm_kernel|pconv_step_kernel<9, false, false>|pconv_step_kernel<1[12], false, true>|pconv_mac_tma|"
    r"large_cols_kernel<7|large_rows_kernel<7, 8, false, true|cfft_kernel<1[02], false|rfft_fwd_reg_kernel<11|"
    r"dconv_fir_kernel<16|pconv_push_ir_kernel<9|fft_thread_kernel<5|fft_thread_kernel<1, 0, false|rfft_fwd_reg_kernel<13")]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["LDG.E.64", "LDG.E.128", "LDG.E (32)", "STG.E.64", "STG.E.128", "STG.E (32)", "LDS.64", "LDS.128", "STS.64", "STS.128",
        "LDGSTS", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FFMA", "FADD",
        "FMUL", "SHFL", "BAR", "total"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    h = hist[cur]
    h["total"] += 1
    base = op.split(".")[0]
    if base in ("LDG", "STG"):
        w = "128" if ".128" in op else ("64" if ".64" in op else None)
        h[f"{base}.E.{w}" if w else f"{base}.E (32)"] += 1
    elif base in ("LDS", "STS"):
        if ".128" in op:
            h[base + ".128"] += 1
        elif ".64" in op:
            h[base + ".64"] += 1
    elif base in ("LDGSTS", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FFMA",
                  "FADD", "FMUL", "SHFL", "BAR"):
        h[base] += 1
print("# SASS instruction counts per kernel (static), libb200fft.so built with -gencode arch=compute_100a,code=sm_100a")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = cp.async.bulk.prefetch.L2, LDTM/STTM = tcgen05.ld/st (tensor memory),")
print("# SYNCS = mbarrier operations, F*2 = packed two-lane FP32")
for name, h in hist.items():
    d = demangle(name)
    if not any(w.search(d) for w in want):
        continue
    print("\n" + d.split("(")[0])
    print("  " + "  ".join(f"{k}={h[k]}" for k in KEYS if h[k]))
