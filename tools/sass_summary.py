"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05 MMA, tensor memory, TMA-engine bulk
copies, mbarrier), taken from the BUILT libraries with cuobjdump.  Runs without a GPU.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = ["torch-nerf_b200/lib/libnerf_b200.so", "torch-nerf_b200/lib/libnerf_b200_selftest.so"]
# mnemonic prefix -> meaning
WATCH = [("UTCHMMA", "tcgen05.mma kind::f16 (5th-gen tensor core MMA, accumulator in TMEM)"),
         ("UTCBAR", "tcgen05.commit (MMA completion -> mbarrier)"),
         ("LDTM", "tcgen05.ld (TMEM -> registers)"),
         ("STTM", "tcgen05.st (registers -> TMEM)"),
         ("UTCATOMSWS", "tcgen05.alloc / dealloc (TMEM allocation)"),
         ("UBLKCP", "cp.async.bulk (TMA engine, 1-D bulk copy global<->shared)"),
         ("UTMALDG", "cp.async.bulk.tensor (TMA tensor-map load)"),
         ("SYNCS", "mbarrier arrive / try_wait"),
         ("HMMA", "legacy mma.sync tensor-core path"),
         ("REDG", "red.global (vector fp32 reductions of the wgrad flush)"),
         ("FADD2", "packed fp32x2 add"),
         ("F2FP", "fp32 -> bf16x2 conversions")]


def main():
    for lib in LIBS:
        path = os.path.join(ROOT, lib)
        if not os.path.exists(path):
            print(f"{lib}: not built")
            continue
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        kernels = collections.OrderedDict()
        cur = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                op = m.group(1)
                kernels[cur]["_total"] += 1
                for key, _ in WATCH:
                    if op.startswith(key):
                        kernels[cur][key] += 1
        arch = set(re.findall(r"arch = (sm_\w+)", out))
        print(f"== {lib}   (cuobjdump -sass; arch {', '.join(sorted(arch))}; {len(kernels)} kernels)")
        names = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
        hdr = f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k, _ in WATCH)
        print(hdr)
        for (mangled, cnt), name in zip(kernels.items(), names):
            short = re.sub(r"\(.*", "", name).replace("nerf::", "")[:58]
            print(f"{short:58s} {cnt['_total']:7d} " + " ".join(f"{cnt[k]:8d}" if cnt[k] else f"{'.':>8s}" for k, _ in WATCH))
        print()
    print("legend:")
    for k, what in WATCH:
        print(f"  {k:11s} {what}")


if __name__ == "__main__":
    main()
