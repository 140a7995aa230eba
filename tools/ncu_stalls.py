"""Top stall sites of one kernel from an .ncu-rep source page.

    python tools/ncu_stalls.py REPORT KERNEL_REGEX [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                                  text=True, stderr=subprocess.DEVNULL)
    # the page holds one table per matching launch; keep the first
    blocks = raw.split('"Kernel Name"')
    body = '"Kernel Name"' + blocks[1]
    rows = list(csv.reader(io.StringIO(body)))
    print(rows[0][1][:100])
    h = rows[1]
    ci = {n: i for i, n in enumerate(h)}
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: 0 for n in stall_cols}
    recs = []
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        s = int(r[ci["# Samples"]] or 0)
        for n in stall_cols:
            tot[n] += int(r[ci[n]] or 0)
        recs.append((s, r))
    all_s = sum(s for s, _ in recs)
    print("samples %d over %d instructions" % (all_s, len(recs)))
    print("by reason:", ", ".join("%s %d" % (n[6:], v) for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    for idx, (s, r) in enumerate(recs):
        r.append(idx)
    for s, r in sorted(recs, key=lambda x: -x[0])[:top]:
        why = sorted(((int(r[ci[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:2]
        print("%5d  #%4d  exec %7s  %-60s %s" % (s, r[-1], r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:60],
                                                 " ".join("%s:%d" % (n, v) for v, n in why if v)))


if __name__ == "__main__":
    main()
