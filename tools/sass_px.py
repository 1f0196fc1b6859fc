"""Instructions per pixel of the full-tile sampling loop, from the SASS of the built library (no GPU needed).
    python tools/sass_px.py [kernel-substring]"""
import re, subprocess, sys, collections
pat = sys.argv[1] if len(sys.argv) > 1 else "warp_pipe_kernelILi1"
out = subprocess.run(["cuobjdump", "-sass", "ransac_with_homography_b200/libb200geo.so"], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
f = next(x for x in funcs if x.startswith("_ZN") and pat in x.split("\n")[0])
ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(.*?);", f)]
# the unrolled loop: stretches between consecutive IDP.2A groups; a pixel = 6 IDP
idp = [i for i, s in enumerate(ins) if "IDP.2A" in s]
groups = [idp[k] for k in range(0, len(idp), 6)]
spans = [b - a for a, b in zip(groups, groups[1:])]
print("kernel", f.split("\n")[0][:80]); print("total SASS", len(ins), "IDP", len(idp), "pixels unrolled", len(groups)); print("instr per pixel (between IDP groups):", spans)
if len(groups) >= 3:
    a, b = groups[1], groups[2]
    c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", s).split()[0].split(".")[0] for s in ins[a:b])
    print(dict(c.most_common()))
    if len(sys.argv) > 2:
        print("\n".join(ins[a:b]))
