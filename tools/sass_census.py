"""SASS census of libbgp_b200.so: per kernel, how many DMMA (fp64 tensor MMA), DFMA, UBLKCP (bulk TMA copies), SYNCS (mbarrier),
LDS / STS, LDG / STG, ATOM / RED and spill instructions it contains.  CPU only (cuobjdump).
usage: python tools/sass_census.py [lib] > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOM", "RED", "BAR", "NANOSLEEP",
        "UTCMMA", "LDTM"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "branchedgp_b200", "libbgp_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(counts, demangle)) if len(demangle) == len(counts) else {k: k for k in counts}
    print("SASS census of %s (sm_100a); tcgen05 (UTCMMA / LDTM) has no fp64 kind, so the fp64 tensor path is DMMA fed by UBLKCP + SYNCS"
          % os.path.basename(lib))
    print("%-72s %7s " % ("kernel", "total") + " ".join("%6s" % k for k in KEYS))
    tot = collections.Counter()
    for fn, c in counts.items():
        nm = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", names[fn])
        nm = nm.replace("bgp::", "").replace("void ", "")
        print("%-72s %7d " % (nm[:72], c["total"]) + " ".join("%6d" % c[k] for k in KEYS))
        tot.update(c)
    print("%-72s %7d " % ("ALL KERNELS", tot["total"]) + " ".join("%6d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    main()
