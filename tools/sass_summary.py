#!/usr/bin/env python
"""SASS / ptxas evidence for profiles/: per kernel of lib/libpdm_b200.so the count of the Blackwell-specific mnemonics
(tcgen05 MMAs, TMEM loads, TMA loads, tcgen05 barriers ...) plus an excerpt of the headline kernel's MMA issue loop, and
the `nvcc -Xptxas -v` register / spill line of every fused_gemm_kernel variant.

    python tools/sass_summary.py > profiles/r2_sass_fused_gemm.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-of-diffusion-models_b200")
LIB = os.path.join(PKG, "lib", "libpdm_b200.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UTCATOMSWS", "SYNCS", "FADD2", "FMUL2", "FFMA2", "MUFU.EX2",
             "USETMAXREG", "UCGABAR")


def demangle(name):
    m = re.match(r"_ZN3pdm2tc17fused_gemm_kernelILi(\d)ELi(\d)ELi(\d)ELb(\d)ELb(\d)E", name)
    if m:
        cg, terms, epi, aux, f8 = (int(v) for v in m.groups())
        return (f"pdm::tc::fused_gemm_kernel<CG={cg}, TERMS={terms}, EPI={('STATS', 'STORE', 'TOPK')[epi]}, AUX={aux}, F8={f8}>")
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    return out.split("(")[0] if out else name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a): Blackwell mnemonics per kernel")
    print("# UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, .2CTA = cta_group::2, LDTM = tcgen05.ld (TMEM -> registers),")
    print("# UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, FADD2/FMUL2/FFMA2 = packed fp32")
    for name, lines in funcs.items():
        cnt = collections.Counter()
        for ln in lines:
            for mn in MNEMONICS:
                if re.search(r"\b" + re.escape(mn), ln):
                    cnt[mn] += 1
            if "UTCHMMA.2CTA" in ln or "UTCQMMA.2CTA" in ln:
                cnt[".2CTA MMAs"] += 1
        n_instr = sum(1 for ln in lines if re.search(r"/\*[0-9a-f]{4}\*/", ln))
        if cnt.get("UTCHMMA") or cnt.get("UTCQMMA") or "noised_rows" in name or "row_norms" in name or "merge" in name:
            print(f"\n{demangle(name)}\n    {n_instr} instructions; " + ", ".join(f"{k} {v}" for k, v in cnt.items()))
    head = "_ZN3pdm2tc17fused_gemm_kernelILi2ELi3ELi0ELb0ELb0E"
    for name, lines in funcs.items():
        if name.startswith(head):
            idx = [i for i, ln in enumerate(lines) if "UTCHMMA" in ln]
            print(f"\n# excerpt: MMA issue loop of {demangle(name)} (the first three of the 12 UTCHMMA.2CTA of a k-block; one elected thread issues)")
            lo, hi = idx[0] - 6, idx[min(len(idx) - 1, 2)] + 2
            for ln in lines[max(0, lo):hi]:
                code = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip())
                if code.strip() and not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/$", code):
                    print(code)
            ld = [i for i, ln in enumerate(lines) if "LDTM" in ln]
            print("\n# excerpt: accumulator drain (tcgen05.ld) + packed fp32 round-to-nearest adds")
            for ln in lines[ld[0] - 2:ld[0] + 14]:
                code = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip())
                if code.strip() and not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/$", code):
                    print(code)
            tma = [i for i, ln in enumerate(lines) if "UTMALDG" in ln]
            print("\n# excerpt: TMA producer (four tile loads per stage, cta_group::2 form)")
            for ln in lines[tma[0] - 3:tma[0] + 12]:
                code = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip())
                if code.strip() and not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/$", code):
                    print(code)
    # ptxas -v of the tensor kernels
    print("\n# nvcc -Xptxas -v, stats_tcgen05.cu (launch bounds 384 threads x 1 CTA/SM -> 168 registers at launch; the control warps")
    print("# give registers back with setmaxnreg.dec 40 and the 8 epilogue warps take setmaxnreg.inc 232)")
    cmd = ["nvcc", "-c", os.path.join(PKG, "csrc", "stats_tcgen05.cu"), "-o", "/dev/null", "-O3", "-std=c++17", "-lineinfo",
           "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-DPDM_BUILD", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr.splitlines()
    name = None
    for ln in err:
        m = re.search(r"Function properties for (\S+)", ln)
        if m:
            name = m.group(1)
            prop = []
        elif name and ("spill" in ln or "Used" in ln):
            prop.append(ln.replace("ptxas info    :", "").strip())
            if "Used" in ln:
                if "fused_gemm" in name:
                    print(f"{demangle(name)}: " + "; ".join(prop))
                name = None
    return 0


if __name__ == "__main__":
    sys.exit(main())
