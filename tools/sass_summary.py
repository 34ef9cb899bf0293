"""Per-kernel SASS evidence of the Blackwell-specific instructions in libcqvad.so (cuobjdump -sass; no GPU needed):
tcgen05.mma (UTCHMMA / .2CTA), TMEM ld/st (LDTM / STTM), TMA loads / stores / reduce-stores (UTMALDG / UTMASTG / UTMAREDG),
multicast commits (UTCBAR), mixed-precision FMA (FHFMA), plus the coordinate FFMA of the reference MSDA kernel.
  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "FHFMA", "HMMA", "RED.E", "LDG.E.ENL2.256"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def summarize(path, only=None):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), {}
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); total[cur] = 0
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            continue
        total[cur] += 1
        for p in PAT:
            if re.search(r"\b" + re.escape(p), ln):
                if p == "UTCHMMA" and "UTCHMMA.2CTA" in ln:
                    continue
                counts[cur][p] += 1
    dm = demangle(list(counts))
    for k, c in counts.items():
        name = re.sub(r"cqvad::\(anonymous namespace\)::|\(anonymous namespace\)::", "", dm[k])
        name = name.split("(")[0][:110]
        if only and not re.search(only, name):
            continue
        if not c and not only:
            continue
        print(f"{name:112s} instrs {total[k]:6d}  " + "  ".join(f"{p} {n}" for p, n in c.items()))
    return txt


print("== class_query_vad_b200/libcqvad.so (sm_100a): kernels with tensor-core / TMEM / TMA / mixed-FMA instructions")
summarize(os.path.join(ROOT, "class_query_vad_b200", "libcqvad.so"))
ref = os.path.join(ROOT, "baseline", "_ref", "MultiScaleDeformableAttention.so")
if os.path.exists(ref):
    print("\n== reference MSDA kernel (baseline/_ref, built from ops/src for sm_100a): the sampling coordinate is ONE fused multiply-add")
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z31ms_deformable_im2col_gpu_kernelIfEviPKT_PKlS4_S2_S2_iiiiiiiPS0_", ref],
                         capture_output=True, text=True).stdout
    for ln in txt.split("\n"):
        if "FFMA" in ln and "-0.5" in ln:
            print("   ", ln.strip()[:100])
