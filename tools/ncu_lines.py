"""Instructions executed / stall samples per CUDA source line of one kernel (needs -lineinfo + --import-source).

    python tools/ncu_lines.py REPORT KERNEL_REGEX [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                                   "--kernel-name", "regex:" + rx], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    out = []
    fpath = ""
    hdr = None
    first_fn = None
    cur_fn = None
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if first_fn is None:
                first_fn = r[1]
            cur_fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or cur_fn != first_fn or len(r) < len(hdr):
            continue
        if not r[0].strip().isdigit():
            continue  # SASS rows: their counts are already in the CUDA line above them
        try:
            ex = int(r[hdr["Instructions Executed"]] or 0)
            sm = int(r[hdr["# Samples"]] or 0)
        except ValueError:
            continue
        if ex or sm:
            out.append((ex, sm, fpath, r[0], r[1].strip()[:90]))
    tot_e, tot_s = sum(o[0] for o in out), sum(o[1] for o in out)
    print(first_fn[:100])
    print("instructions %d, samples %d" % (tot_e, tot_s))
    for ex, sm, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
        print("%6.2f%% inst %5.2f%% smp  %s:%s  %s" % (100.0 * ex / tot_e, 100.0 * sm / max(tot_s, 1), f, ln, src))


if __name__ == "__main__":
    main()
