"""SASS instruction histogram of the default kernels of the built library (static counts per mnemonic class).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("ALICE_LIB", os.path.join(ROOT, "alice-codec_b200", "lib", "libalice_codec.so"))
KERNELS = ["k_fwd_fused<1, false>", "k_fwd_fused<0, false>", "k_inv_fused<1>", "k_inv_fused<0>", "k_rans_encode<1>", "k_rans_encode<4>",
           "k_rans_decode<1>", "k_rans_decode<4>", "k_fwd_xy<1,", "k_fwd_t_quant<1, 4, 64>", "k_inv_t<1, 4, 64, false>", "k_inv_yx<1,",
           "k_wxy<1, false>", "k_wt<1, false>", "k_build_tables", "k_estimate_stream_bytes"]
CLASSES = [("tensor/bulk copy", r"^(UBLKCP|UTMA|SYNCS)"), ("shuffle/vote", r"^(SHFL|VOTE|MATCH|REDUX)"), ("shared ld/st", r"^(LDS|STS|ATOMS|LDSM)"), ("global ld/st", r"^(LDG|STG|LD\.|ST\.|RED|ATOMG|ATOM)"),
           ("local (spill)", r"^(LDL|STL)"), ("integer pipe", r"^(IADD3|VIADD|LOP3|SHF|SEL|ISETP|LEA|PRMT|IABS|IMNMX|VIMNMX|BMSK|SGXT|POPC|FLO|PLOP3|P2R|R2P|VABSDIFF|I2I|MOV|CS2R|S2R)"),
           ("multiplier pipe", r"^(IMAD|IMUL)"), ("branch/sync", r"^(BRA|BSSY|BSYNC|BAR|WARPSYNC|EXIT|CALL|RET|NANOSLEEP|YIELD|BREAK|BMOV)"),
           ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR)"), ("fp", r"^(F[A-Z]|MUFU|I2F|F2I|DADD|DMUL|DFMA)")]

dem = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input=dem, capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", filt)
print("# SASS instruction histogram of the final round-2 build (static counts; `tools/sass_histogram.py`)\n")
print("| kernel | total | " + " | ".join(c for c, _ in CLASSES) + " | other |")
print("|---|---|" + "---|" * (len(CLASSES) + 1))
for want in KERNELS:
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if "alice::" + want in name:
            ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", b)
            cnt = collections.Counter()
            for op in ops:
                for cname, rx in CLASSES:
                    if re.match(rx, op):
                        cnt[cname] += 1
                        break
                else:
                    cnt["other"] += 1
            short = re.sub(r"\(.*", "", name.replace("void alice::", "").replace("alice::", ""))
            print(f"| `{short}` | {len(ops)} | " + " | ".join(str(cnt[c]) for c, _ in CLASSES) + f" | {cnt['other']} |")
            break
    else:
        sys.stderr.write(f"not found: {want}\n")
