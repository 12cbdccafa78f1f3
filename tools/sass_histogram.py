"""Instruction histogram of the hot kernels of libpn2b200.so from `cuobjdump -sass` (no GPU needed): the mnemonics that
prove which hardware path a kernel uses -- UTCHMMA / UTCBAR / LDTM (tcgen05.mma / commit / ld: 5th-generation tensor
cores, TMEM), UBLKCP (cp.async.bulk), LDGSTS (cp.async), LDSM / STSM (ldmatrix / stmatrix), HMMA (legacy mma.sync),
REDG / RED (vector reductions), SYNCS (mbarrier).   python tools/sass_histogram.py [kernel regex] > profiles/rNN_sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "UTCCP", "UBLKCP", "UTMALDG", "LDGSTS", "LDSM", "STSM", "HMMA", "REDG", "RED.", "ATOMG",
        "SYNCS", "BAR.SYNC", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "REDUX", "VOTE", "NANOSLEEP")


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                     r"gemm_tc_kernel|wgrad_tc_kernel|wgrad_kernel|fps_regs|ball_query_kernel|knn_kernel|three_nn_kernel|"
                     r"sa_build_rows|fp_build_rows|pool_fwd|pool_bwd|cm_to_rows|rows_to_cm|sa_rows_bwd|fp_rows_bwd|kabsch")
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "hotrack_b200", "libpn2b200.so")], capture_output=True,
                         text=True, check=True).stdout
    cur, hist, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"pn2::\(anonymous namespace\)::", "", name)
            cur = name if pat.search(name) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            total[cur] += 1
            for k in KEYS:
                if op.startswith(k):
                    hist[cur][k] += 1
    print("# SASS instruction histogram, libpn2b200.so (sm_100a); counts are static instructions per kernel")
    for name, h in hist.items():
        short = re.sub(r"\(.*", "", name)[:100]
        print("%-100s total %5d  " % (short, total[name]) + "  ".join("%s=%d" % (k, v) for k, v in h.items() if v))


if __name__ == "__main__":
    main()
