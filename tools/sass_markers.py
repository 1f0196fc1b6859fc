"""Static counts, per kernel of the built library, of the opcodes that identify the Blackwell-specific paths.
    python tools/sass_markers.py > profiles/r02_sass_markers.txt"""
import collections, re, subprocess
out = subprocess.run(["cuobjdump", "-sass", "ransac_with_homography_b200/libb200geo.so"], capture_output=True, text=True).stdout
marks = ["UTMALDG", "UBLKCP", "SYNCS", "IDP.2A", "IDP.4A", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "VIMNMX3", "PRMT", "LDS", "STG", "MUFU.RCP", "SHFL", "UTCHMMA",
         "UTCQMMA", "LDTM", "HMMA", "LDGSTS", "POPC", "DFMA", "ATOMS"]
print("# cuobjdump -sass ransac_with_homography_b200/libb200geo.so (sm_100a): static instruction counts per kernel of the opcodes that")
print("# identify the Blackwell-specific paths (UTMALDG = TMA tensor copy, UBLKCP = bulk copy, SYNCS = mbarrier, IDP.2A = 2-way u16 x u8 dot")
print("# product, FFMA2 / FMUL2 / FADD2 = packed FP32 pairs, FMNMX3 / VIMNMX3 = 3-input min / max).  No UTC*MMA / LDTM / HMMA anywhere: the")
print("# path has no tensor-core work by design (K = 3 contraction; north_star).")
for f in re.split(r"\n\s*Function : ", out)[1:]:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(.*?);", f)]
    if not ins or "sample_kernel" in dem or "warp_generic_kernel" in dem:
        continue
    c = collections.Counter()
    for s in ins:
        op = re.sub(r"^@!?U?P\d\s+", "", s).split()[0]
        for m in marks:
            if op.startswith(m):
                c[m] += 1
    print(f"{dem[-64:]:64s} total {len(ins):6d}  " + "  ".join(f"{k} {v}" for k, v in c.items() if v))
